"""Golden outputs of the UNMODIFIED reference for the step around the path: ``FeatureDistillation.replay``
(mafed/methods/distillation.py:84-103) and its caller ``VLPythiaVQACLearner.training_step``
(mafed/model/vqa_cont_learner.py:213-236).

    python tests/golden/make_golden_replay.py          (build container: needs /root/reference or oracle/_ref)

The reference classes are imported unmodified through ``oracle/ref_harness.py`` (Lightning, toolz and the
reference's data / model packages stubbed) and driven on the CPU with ``tests/tiny_vl.TinyVL`` (a GPT-NeoX decoder
behind vision tokens, the keyword interface of ``VLCLIPGPTNeoXForCausalLM.forward``) standing in for the model.  The
model runs as an fp32 island (``fp32_island=True``), so the reference's ``torch.autocast("cuda", bfloat16)`` region
-- inert on CPU tensors -- and the GPU run of the mirror see the same fp32 hidden states.

Writes ``tests/golden/replay_cases.npz``: the two models' weights, the batches, and per case the loss, ``n_ex``,
every parameter gradient's norm and 64 sampled entries, the keys ``replay`` left in the batch, the mask sums, the
values sent to W&B, and for ``training_step`` the per-batch losses and Lightning ``log`` calls.
"""
import copy
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import ref_harness as R  # noqa: E402
from tiny_vl import TinyVL, make_batch  # noqa: E402

N_VIS, N_SAMPLES = 8, 64
REPLAY_CASES = [
    dict(name="discounted", layer_strategy="discounted", layer=None, modality="balanced", loss="mse", replay_coeff=1.0,
         coeff=2.0, task_id=1, accumulate=2),
    dict(name="single2", layer_strategy="single", layer=2, modality="balanced", loss="mse", replay_coeff=1.0, coeff=2.0,
         task_id=1, accumulate=2),
    dict(name="equal_cosine", layer_strategy="equal", layer=None, modality="equal", loss="cosine", replay_coeff=0.5,
         coeff=1.0, task_id=2, accumulate=1),
    dict(name="no_distill", layer_strategy="equal", layer=None, modality="balanced", loss="mse", replay_coeff=0.5,
         coeff=0.0, task_id=1, accumulate=1),
    dict(name="first_task", layer_strategy="equal", layer=None, modality="balanced", loss="mse", replay_coeff=1.0,
         coeff=1.0, task_id=0, accumulate=1),     # task_id 0: no LM replay term (distillation.py:88)
]


class Opts(R.Opts):
    pass


def models():
    torch.manual_seed(0)
    student = TinyVL(n_vis=N_VIS, fp32_island=True)
    teacher = copy.deepcopy(student)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in student.parameters():
            p.add_(0.02 * torch.randn(p.shape, generator=g))      # the student has moved away from the teacher
    return student, teacher


def sample_positions(numel, k=N_SAMPLES):
    rng = np.random.default_rng(11)
    return np.sort(rng.choice(numel, size=min(k, numel), replace=False)).astype(np.int64)


def make_method(case):
    opts = Opts()
    opts.accumulate_grad_batches = case["accumulate"]
    fd = R.make_reference_method(modality=case["modality"], layer_strategy=case["layer_strategy"], loss=case["loss"],
                                 gamma=0.5, num_hidden_layers=3, layer=case["layer"], coeff=case["coeff"], n_vis=N_VIS,
                                 opts=opts, replay_coeff=case["replay_coeff"])
    fd.task_id = case["task_id"]
    return fd


def grads_blob(model, prefix, blob):
    names = []
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(name)
        blob[f"{prefix}/gnorm/{name}"] = np.array(float(p.grad.double().norm()))
        pos = sample_positions(p.grad.numel())
        blob[f"{prefix}/gsamp/{name}"] = p.grad.reshape(-1)[pos].double().numpy()
    return names


def main():
    R.load()
    blob, meta = {}, {"replay": [], "n_vis": N_VIS}
    student, teacher = models()
    for k, v in student.state_dict().items():
        blob[f"student/{k}"] = v.numpy()
    for k, v in teacher.state_dict().items():
        blob[f"teacher/{k}"] = v.numpy()
    batch = make_batch(seed=0)
    for k, v in batch.items():
        blob[f"batch/{k}"] = v.numpy()
    for case in REPLAY_CASES:
        fd = make_method(case)
        model = copy.deepcopy(student)
        fd.past_model = copy.deepcopy(teacher).eval()
        b = {k: v.clone() for k, v in batch.items()}
        fd.mem_dataloader = [b]
        R.wandb_log.clear()
        loss, n_ex = fd.replay(model)
        rec = dict(case, n_ex=int(n_ex), step=int(fd.step), loss_is_none=loss is None, batch_keys=sorted(b.keys()))
        if loss is not None:
            (loss / case["accumulate"]).backward()
            blob[f"replay/{case['name']}/loss"] = np.array(float(loss))
            rec["grad_names"] = grads_blob(model, f"replay/{case['name']}", blob)
        logged = {}
        for d in R.wandb_log:
            logged.update(d)
        rec["logged"] = {k: float(v) for k, v in logged.items()}
        if "lang_masks" in b:
            rec["lang_sum"], rec["image_sum"] = int(b["lang_masks"].sum()), int(b["image_masks"].sum())
            rec["mask_shape"] = list(b["lang_masks"].shape)
        meta["replay"].append(rec)
        print(case["name"], None if loss is None else float(loss), n_ex, rec["batch_keys"], flush=True)

    # ---- the caller: VLPythiaVQACLearner.training_step, unmodified, over two tasks x four batches
    Learner = R.learner_class()
    case = dict(layer_strategy="discounted", layer=None, modality="balanced", loss="mse", replay_coeff=1.0, coeff=1.0,
                task_id=0, accumulate=2)
    fd = make_method(case)
    model = copy.deepcopy(student)
    fd.past_model = copy.deepcopy(teacher).eval()
    mem = make_batch(seed=3)
    task_batch = make_batch(seed=4)
    for k, v in mem.items():
        blob[f"ts_mem/{k}"] = v.numpy()
    for k, v in task_batch.items():
        blob[f"ts_task/{k}"] = v.numpy()

    class FreshBatches:                                # a DataLoader yields a new dict per iteration
        def __iter__(self):
            yield {k: v.clone() for k, v in mem.items()}

    fd.mem_dataloader = FreshBatches()
    learner = Learner.__new__(Learner)
    learner.model, learner.cl_method = model, fd
    learner.config = types.SimpleNamespace(replay_interval=2)
    steps = []
    for task_id in (0, 1):
        learner.task_id = fd.task_id = task_id
        for idx in range(4):
            learner.logged = []
            model.zero_grad(set_to_none=True)
            loss = learner.training_step({k: v.clone() for k, v in task_batch.items()}, idx)
            (loss / case["accumulate"]).backward()             # Lightning: loss / accumulate_grad_batches
            name, value, kwargs = learner.logged[0]
            assert len(learner.logged) == 1 and value is loss
            tag = f"ts/{task_id}_{idx}"
            blob[f"{tag}/loss"] = np.array(float(loss))
            steps.append(dict(task_id=task_id, batch_idx=idx, log_name=name, log_kwargs={k: (v if not torch.is_tensor(v) else float(v)) for k, v in kwargs.items()},
                              grad_names=grads_blob(model, tag, blob), fd_step=int(fd.step)))
            print("training_step", task_id, idx, name, float(loss), flush=True)
    meta["training_step"] = dict(case=case, steps=steps)
    blob["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "replay_cases.npz"), **blob)
    print("wrote replay_cases.npz", os.path.getsize(os.path.join(HERE, "replay_cases.npz")))


if __name__ == "__main__":
    main()
