"""Host-side mirror of mafed.methods: layer plans, modality tables, API surface (not gpu)."""
import inspect

import pytest
import torch

from golden_util import load_plans
from mafed_b200 import cabi
from mafed_b200.methods import CLMethod, CLStrategy, DistillationWeights, FeatureDistillation, Naive
from mafed_b200.methods.distillation_loss_weights import modality_masks


class Opts:
    tasks = ["a", "b", "c"]
    batch_size = 4
    seed = 42
    pin_mem = False
    accumulate_grad_batches = 4


def test_registry_and_base_api():
    assert CLMethod["featdistill"] is FeatureDistillation and CLMethod["naive"] is Naive
    s = CLStrategy(opts=Opts())
    assert (s.task_id, s.reg_lambda, s.mask, s.scaler, s.update_freq) == (0, 1.0, None, None, 4)
    assert s.replay(model=None) == (None, 0)
    assert [s._is_batch_after_step(i) for i in range(4)] == [False, False, False, True]
    s.update(model=None)
    assert s.task_id == 1
    with pytest.raises(NotImplementedError):
        s.compute_loss(None, 1.0)
    assert Naive().compute_loss(None, 3.0) == 3.0 and CLStrategy().update_freq == 1


def test_constructor_signature_matches_reference():
    # mafed/methods/distillation.py:19-34 and distillation_loss_weights.py:10-22
    sig = inspect.signature(FeatureDistillation.__init__)
    names = list(sig.parameters)
    assert names[:14] == ["self", "memory_size", "opts", "model_type", "distillation_modality_weighing_strategy",
                          "distillation_layer_weighing_strategy", "distillation_coeff", "replay_coeff",
                          "distillation_layer", "cls_distillation", "distillation_loss", "gamma",
                          "num_hidden_layers", "kwargs"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["distillation_modality_weighing_strategy"], d["distillation_layer_weighing_strategy"]) == ("equal", "single")
    assert (d["distillation_coeff"], d["replay_coeff"], d["distillation_layer"], d["gamma"], d["num_hidden_layers"]) == \
        (1.0, 1.0, -1, 0.8, 11)
    wsig = inspect.signature(DistillationWeights.__init__)
    wd = {k: v.default for k, v in wsig.parameters.items()}
    assert (wd["gamma"], wd["num_hidden_layers"], wd["distillation_layer"], wd["num_vision_tokens"]) == (0.9, 11, -1, 256)
    for m in ("update", "compute_loss", "replay", "distill", "feature_distillation", "update_after_new_task",
              "update_after_step", "update_mask", "_get_past_hidden_states", "_compute_mse_distillation_loss",
              "_compute_cosine_distillation_loss", "_compute_cls_distillation_loss", "_update_memory", "_update_model"):
        assert callable(getattr(FeatureDistillation, m))


def test_layer_plans_match_reference():
    for rec in load_plans()["plans"]:
        kw = dict(distillation_modality_weighing_strategy="balanced",
                  distillation_layer_weighing_strategy=rec["strategy"], gamma=rec["gamma"],
                  num_hidden_layers=rec["num_hidden_layers"], distillation_layer=rec["layer"])
        if "error" in rec:
            with pytest.raises(AssertionError):
                DistillationWeights(**kw)
            continue
        dw = DistillationWeights(**kw)
        layers = dw.get_distillation_layers()
        assert layers == rec["layers"] and dw._layer_weighing_strategy == rec["effective"]
        assert [float(dw.get_layer_loss_weight(l)) for l in layers] == rec["coeffs"]  # bit-exact fp32
        coeffs, kind, lang = dw.kernel_tables()
        assert coeffs == rec["coeffs"] and kind == cabi.MODW_TABLE and lang == [0.5] * len(layers)


def test_feature_distillation_constructor_quirks():
    for rec in load_plans()["resolves"]:
        kw = dict(memory_size=8, opts=Opts(), model_type="x", distillation_layer=rec["distillation_layer"],
                  distillation_layer_weighing_strategy="equal", num_hidden_layers=rec["num_hidden_layers"])
        if "error" in rec:
            with pytest.raises(AssertionError):
                FeatureDistillation(**kw)
            continue
        fd = FeatureDistillation(**kw)
        assert fd.loss_weights.get_distillation_layers() == rec["layers"]
        assert fd.memory_per_task == rec["memory_per_task"]
    fd = FeatureDistillation(8, Opts(), "x", distillation_layer_weighing_strategy="discounted", distillation_layer=None,
                             some_unknown_kwarg=1)
    assert fd.update_freq == 4 and fd.num_vision_tokens == 256 and fd.step == 0 and fd.past_model is None
    assert fd.compute_loss(None, 2.5, batch={}) == 2.5
    with pytest.raises(AssertionError):  # CLI default: single + no layer
        FeatureDistillation(8, Opts(), "x", distillation_layer=None)


def test_modality_tables():
    dw = DistillationWeights("equal", "equal", num_hidden_layers=3, distillation_layer=None)
    assert dw.kernel_tables()[1:] == (cabi.MODW_EQUAL, None)
    am = torch.tensor([[0, 1, 1], [1, 1, 1]])
    lang, img = modality_masks(am, 4)
    assert lang.tolist() == [[0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 1, 1, 1]] and lang.dtype == am.dtype
    assert img.tolist() == [[1, 1, 1, 1, 0, 0, 0]] * 2
    lw, vw = dw.get_modality_loss_weights({"lang_masks": lang, "image_masks": img}, 0)
    assert float(lw) == pytest.approx(5 / 13) and float(vw) == pytest.approx(8 / 13)
    ad = DistillationWeights("adaptive", "equal", num_hidden_layers=3, distillation_layer=None)
    ad.lang_coeff = torch.tensor([0.25, 0.5, 0.75])
    assert ad.kernel_tables([0, 2])[2] == [0.25, 0.75]
    assert ad.get_modality_loss_weights({}, 1) == (0.5, 0.5)
    ad.lang_coeff = torch.tensor([0.4])
    assert ad.kernel_tables()[2] == pytest.approx([0.4] * 3)
    with pytest.raises(NotImplementedError):
        DistillationWeights("bogus", "equal", distillation_layer=None).get_modality_loss_weights({}, 0)
    with pytest.raises(ValueError):
        dw._get_dynamic_loss_weights(None)


def test_cpu_tensors_are_refused_loudly():
    from mafed_b200.distill_op import DistillPlan, distill_loss
    plan = DistillPlan(layers=[0], layer_coeffs=[1.0], n_vis=2)
    s = torch.randn(1, 3, 8, requires_grad=True)
    with pytest.raises(cabi.MafedDistillError, match="no CPU fallback"):
        distill_loss([s], [torch.randn(1, 3, 8)], torch.ones(1, 1, dtype=torch.int64), plan, group=False)


def test_constructed_exactly_like_train_py():
    """mafed/train.py:119-134 keyword set, including the ones the strategy swallows."""
    fd = CLMethod["featdistill"](
        opts=Opts(), memory_size=500, model_type="vlpythia", scaler=object(), reg_lambda=1.0, replay_coeff=1.0,
        distillation_coeff=1.0, distillation_modality_weighing_strategy="balanced",
        distillation_layer_weighing_strategy="discounted", distillation_layer=None, cls_distillation=False,
        distillation_loss="mse", gamma=0.5, num_hidden_layers=15)
    assert fd.loss_weights.get_distillation_layers() == list(range(15))
    assert fd.memory_per_task == 250 and fd.assumed_grad_out == 0.25 and fd.single_pass
    assert fd.replay_coeff == 1.0 and fd.distillation_coeff == 1.0 and fd.task_id == 0
    fd.num_training_steps = 10  # attribute write done by vqa_cont_learner.py:211
    fd.update_after_backward(model=None)
    fd.update_after_step(model=None, batch_idx=3)
    fd.update_after_new_task(model=None, dataset=None)
    fd.update_mask()


class _ToyDataset(torch.utils.data.Dataset):
    def __init__(self, n, base=0):
        self.n, self.base = n, base

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return {"input_ids": torch.tensor([self.base + i])}


def _collate(items):
    return {"input_ids": torch.stack([it["input_ids"] for it in items])}


def test_update_memory_and_model_like_reference():
    """distillation.py:75-79,182-213: teacher snapshot, rng-sampled memory, loader rebuilt, task_id += 1."""
    import numpy as np
    fd = FeatureDistillation(8, Opts(), "vlpythia", distillation_modality_weighing_strategy="balanced",
                             distillation_layer_weighing_strategy="equal", distillation_layer=None, num_hidden_layers=2)
    fd.num_workers = 0
    fd.data_hooks = (_collate, lambda loader: loader)           # stands in for mafed.data collate_fn / PrefetchLoader
    model = torch.nn.Linear(2, 2)
    model.train()
    fd.update(dataset=_ToyDataset(50), model=model, dataloader=None)
    assert fd.task_id == 1 and fd.past_model is not model and not fd.past_model.training
    want = np.random.default_rng(Opts.seed).choice(np.arange(50), 4, replace=False)   # memory_per_task = 8 / 2
    assert sorted(fd.datasets[0].indices) == sorted(want)
    batch = fd._next_memory_batch()
    assert set(batch["input_ids"].flatten().tolist()) <= set(int(i) for i in want)
    assert isinstance(fd.mem_sampler, torch.utils.data.RandomSampler)
    fd.update(dataset=_ToyDataset(30, base=1000), model=model, dataloader=None)
    assert fd.task_id == 2 and len(fd.mem_dataloader.dataset) == 8
    # persistent iterator: every memory sample exactly once per epoch, then a new epoch starts
    fd.persistent_memory_iterator = True
    seen = [fd._next_memory_batch()["input_ids"].flatten().tolist() for _ in range(2)]
    assert sorted(sum(seen, [])) == sorted(x for ds in fd.datasets for x in
                                           [int(ds.dataset[i]["input_ids"]) for i in ds.indices])
    assert fd._next_memory_batch()["input_ids"].numel() == 4 and fd._mem_epoch == 1


def test_strategy_state_roundtrip():
    """Checkpoint / resume of the strategy state the reference keeps only in the live object."""
    import numpy as np
    kw = dict(distillation_modality_weighing_strategy="adaptive", distillation_layer_weighing_strategy="equal",
              distillation_layer=None, num_hidden_layers=2)
    fd = FeatureDistillation(8, Opts(), "vlpythia", **kw)
    fd.num_workers = 0
    fd.data_hooks = (_collate, lambda loader: loader)
    fd._update_memory(_ToyDataset(50))
    fd.task_id, fd.step = 1, 17
    fd.loss_weights.lang_coeff = torch.tensor([0.3, 0.7])
    state = fd.state_dict()
    nxt = fd.rng.choice(np.arange(100), 5, replace=False)
    other = FeatureDistillation(8, Opts(), "vlpythia", **kw)
    other.num_workers = 0
    other.data_hooks = (_collate, lambda loader: loader)
    other.load_state_dict(state, datasets=[_ToyDataset(50)])
    assert (other.task_id, other.step) == (1, 17)
    assert sorted(other.datasets[0].indices) == sorted(fd.datasets[0].indices)
    assert other.loss_weights.kernel_tables()[2] == pytest.approx([0.3, 0.7])
    assert (other.rng.choice(np.arange(100), 5, replace=False) == nxt).all()   # sampling continues identically
    # a restored strategy can RUN: the memory loader exists again, the next task's update() works
    batch = other._next_memory_batch()
    assert batch["input_ids"].shape[0] == min(Opts.batch_size, other.memory_per_task)
    assert other.mem_sampler is not None
    other.loss_weights.update_weights = lambda model, dataloader, task_id: None   # (needs a model; not this test)
    other.update(dataset=_ToyDataset(60), model=torch.nn.Linear(2, 2), dataloader=None)
    assert other.task_id == 2 and len(other.datasets) == 2 and other.past_model is not None
    assert len(other.mem_dataloader.dataset) == 2 * other.memory_per_task
    # a fresh strategy restored WITHOUT datasets still updates (no stale `del self.mem_dataloader`)
    bare = FeatureDistillation(8, Opts(), "vlpythia", **kw)
    bare.num_workers, bare.data_hooks = 0, (_collate, lambda loader: loader)
    bare.load_state_dict(state)
    bare._update_memory(_ToyDataset(50))
    assert len(bare.datasets) == 1


def test_assumed_upstream_gradient_follows_the_gate(monkeypatch):
    """The backward gate leaves the upstream gradient it saw in a pinned word; the next distill() re-aims
    `assumed_grad_out` at it (host logic only: the word is faked here)."""
    import numpy as np
    fd = FeatureDistillation(8, Opts(), "vlpythia", distillation_modality_weighing_strategy="balanced",
                             distillation_layer_weighing_strategy="equal", distillation_layer=None, num_hidden_layers=2)
    assert fd.assumed_grad_out == 1.0 / Opts.accumulate_grad_batches and fd.adapt_assumed_grad_out and fd.single_pass
    fd.assumed_grad_out = 1.0
    word = np.zeros(1, dtype=np.float32)
    fd._gout_seen_np = word
    fd._adapt_assumed()
    assert fd.assumed_grad_out == 1.0                       # nothing seen yet (0 = no backward has run)
    for bad in (float("nan"), float("inf")):
        word[0] = bad
        fd._adapt_assumed()
        assert fd.assumed_grad_out == 1.0                   # an overflowed / poisoned step teaches nothing
    word[0] = 0.25
    fd._adapt_assumed()
    assert fd.assumed_grad_out == 0.25 and fd._gout_changes == 1
    fd._adapt_assumed()
    assert fd._gout_changes == 1                            # unchanged value: no churn
    plan = fd._step_plan([0, 1])
    assert plan.assumed_grad_out == 0.25                    # the plan cache follows
    # under a gradient multiplier the gate sees grad_out * multiplier
    fd.grad_multiplier = 4.0
    plan = fd._step_plan([0, 1])
    assert plan.grad_multiplier == 4.0
    word[0] = 2.0
    fd._adapt_assumed()
    assert fd.assumed_grad_out == 0.5
    # an upstream gradient that never settles (dynamic loss scaling) -> two-pass form
    for i in range(12):
        word[0] = float(2 ** (i + 2))
        fd._adapt_assumed()
    assert not fd.single_pass
