"""replay() end to end on the GPU with a tiny VL model: loss and *parameter* gradients against a plain
torch restatement of distillation.py:84-122 on the same model, with and without selective capture."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


class Opts:
    tasks = ["a", "b", "c"]; batch_size = 4; seed = 1; pin_mem = False; accumulate_grad_batches = 2


def _torch_replay(model, teacher, batch, layers, coeffs, n_vis, replay_coeff, distill_coeff):
    """The reference's op chain (balanced weights, mse) in plain torch."""
    batch = dict(batch)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(**batch, compute_loss=True, output_hidden_states=True, return_dict=True)
        loss = replay_coeff * out.loss
        batch.pop("labels")
        with torch.no_grad():
            past = teacher(**batch, output_hidden_states=True, return_dict=True).hidden_states
        am = batch["attention_mask"]
        B, txt = am.shape
        lang = torch.zeros(B, n_vis + txt, dtype=am.dtype, device=am.device); lang[:, n_vis:] = am
        img = torch.zeros_like(lang); img[:, :n_vis] = 1
        total = 0.0
        for l, c in zip(layers, coeffs):
            h, p = out.hidden_states[l], past[l].detach()
            d = h.shape[-1]
            tok = torch.nn.MSELoss(reduction="none")(h.reshape(-1, d), p.reshape(-1, d)).sum(-1) / d
            lt = (tok * lang.reshape(-1)).sum() / lang.sum()
            lv = (tok * img.reshape(-1)).sum() / img.sum()
            total = total + c * distill_coeff * (0.5 * lt + 0.5 * lv)
        loss = loss + total
    return loss


@pytest.mark.parametrize("selective", [False, True], ids=["full-tuple", "selective-capture"])
@pytest.mark.parametrize("strategy,layer", [("discounted", None), ("single", 2)])
def test_replay_matches_torch_restatement(selective, strategy, layer):
    from mafed_b200.methods import CLMethod
    from tiny_vl import TinyVL, make_batch
    torch.manual_seed(0)
    model = TinyVL().cuda()
    teacher = copy.deepcopy(model)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.02 * torch.randn_like(p))                       # the student has moved away from the teacher
    batch = make_batch(device="cuda")
    fd = CLMethod["featdistill"](memory_size=8, opts=Opts(), model_type="vlpythia",
                                 distillation_modality_weighing_strategy="balanced",
                                 distillation_layer_weighing_strategy=strategy, distillation_layer=layer,
                                 distillation_loss="mse", gamma=0.5, num_hidden_layers=3, replay_coeff=1.0,
                                 distillation_coeff=2.0, selective_capture=selective)
    fd.num_vision_tokens = 8
    fd.task_id = 1
    fd._update_model(teacher)
    fd.mem_dataloader = [dict(batch)]
    layers = fd.loss_weights.get_distillation_layers()
    coeffs = [float(fd.loss_weights.get_layer_loss_weight(l)) for l in layers]

    ref_model = copy.deepcopy(model)
    ref_loss = _torch_replay(ref_model, teacher, batch, layers, coeffs, 8, 1.0, 2.0)
    (ref_loss / 2).backward()                                        # Lightning: loss / accumulate_grad_batches

    loss, n_ex = fd.replay(model)
    assert n_ex == 4 and fd.step == 1
    (loss / 2).backward()
    torch.cuda.synchronize()
    assert float(loss) == pytest.approx(float(ref_loss), rel=1e-5)
    checked = 0
    for (name, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
        if q.grad is None:
            assert p.grad is None, name
            continue
        err = float((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-20))
        assert err < 2e-4, (name, err)
        checked += 1
    assert checked > 10
    assert not any(p.grad is not None for p in fd.past_model.parameters())   # teacher untouched
    assert fd.layer_loss_dict().keys() == {f"task_1/distill_loss_{l}" for l in layers}


def test_replay_without_distillation_returns_lm_loss_only():
    from mafed_b200.methods import CLMethod
    from tiny_vl import TinyVL, make_batch
    torch.manual_seed(0)
    model = TinyVL().cuda()
    fd = CLMethod["featdistill"](memory_size=8, opts=Opts(), model_type="vlpythia",
                                 distillation_layer_weighing_strategy="equal", distillation_layer=None,
                                 num_hidden_layers=3, distillation_coeff=0.0, replay_coeff=0.5)
    fd.task_id = 1
    fd.mem_dataloader = [make_batch(device="cuda")]
    loss, n_ex = fd.replay(model)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        want = 0.5 * model(**make_batch(device="cuda"), compute_loss=True).loss
    assert float(loss) == pytest.approx(float(want), rel=1e-6) and n_ex == 4
    fd.task_id = 0                                                   # no replay on the first task (:88)
    fd.distillation_coeff = 0.0
    assert fd.replay(model)[0] is None


def test_training_step_aggregation_without_lightning():
    """vqa_cont_learner.py:213-236 as a plain function: replay every `replay_interval`-th batch once task_id > 0."""
    from mafed_b200.methods import CLMethod
    from mafed_b200.step import backward_step, training_step_loss
    from tiny_vl import TinyVL, make_batch
    torch.manual_seed(0)
    model = TinyVL().cuda()
    fd = CLMethod["featdistill"](memory_size=8, opts=Opts(), model_type="vlpythia",
                                 distillation_modality_weighing_strategy="balanced",
                                 distillation_layer_weighing_strategy="discounted", distillation_layer=None,
                                 num_hidden_layers=3, gamma=0.5)
    fd.num_vision_tokens = 8
    fd._update_model(model)
    class FreshBatches:                                # a DataLoader yields a new dict per iteration
        def __iter__(self):
            yield make_batch(seed=3, device="cuda")

    fd.mem_dataloader = FreshBatches()
    batch = make_batch(seed=4, device="cuda")
    keys = []
    for task_id in (0, 1):
        fd.task_id = task_id
        for idx in range(4):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss, key = training_step_loss(fd, model, dict(batch), idx, task_id, replay_interval=2)
            backward_step(loss, Opts.accumulate_grad_batches)
            keys.append(key.split("/")[1])
    assert keys == ["train_loss"] * 4 + ["train_loss", "replay_train_loss", "train_loss", "replay_train_loss"]
    assert fd.step == 2 and all(p.grad is not None for p in model.gpt_neox.layers[0].parameters())
