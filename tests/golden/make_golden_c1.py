"""Golden outputs of the UNMODIFIED reference at BASELINE.json's configs[0] -- full size.

    python tests/golden/make_golden_c1.py          (build container only: needs /root/reference)

configs[0] = "VLPythia-base (Pythia-160M, 12 layers, d=768) MAFED distillation loss fwd+bwd, batch 8, 256 visual +
32 text tokens, fp32, on CPU": a 13-entry hidden-state tuple, the reference's own call num_hidden_layers = 12 - 1
(train.py:133), the shipped recipe mse / balanced / discounted gamma 0.5 (scripts/run_seed42.sh:74-93).
The inputs (2 x 13 x 8 x 288 x 768 fp32 = 184 MB) are not committed: they are regenerated from the seed with
oracle.make_inputs (torch CPU generator); this file stores digests of them so that a different random stream is
noticed, the reference's loss and logged per-layer losses, and per layer the gradient's sum, L2 norm and 256 entries
at fixed positions.  Two mask variants: the ragged one of SURVEY 8(d) and all-ones.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from make_golden import _install_stubs, run_reference  # noqa: E402

CASE = dict(modality="balanced", layer_strategy="discounted", loss="mse", coeff=1.0, cls=False, gamma=0.5,
            num_hidden_layers=11, n_tuple=13, layer=None, grad_out=1.0, n_vis=256, txt=32, bsz=8, dim=768,
            dtype="fp32", teacher="close", lang_coeff=None)
SEED = 1234
N_SAMPLES = 256


def sample_positions(numel):
    rng = np.random.default_rng(7)
    return np.sort(rng.choice(numel, size=N_SAMPLES, replace=False)).astype(np.int64)


def main():
    _install_stubs()
    from oracle.distill_oracle import make_inputs
    blob = {}
    for tag, mask_kind in (("ragged", "ragged"), ("ones", "ones")):
        st, te, am = make_inputs(CASE["n_tuple"], CASE["bsz"], CASE["txt"], CASE["dim"], n_vis=CASE["n_vis"],
                                 dtype=torch.float32, seed=SEED, teacher=CASE["teacher"], mask=mask_kind)
        loss, logged, grads = run_reference(CASE, st, te, am)
        pos = sample_positions(st[0].numel())
        blob[f"{tag}_input_digest"] = np.array([[float(s.double().sum()), float(t.double().sum())] for s, t in zip(st, te)])
        blob[f"{tag}_input_samples"] = np.stack([s.reshape(-1)[pos].numpy() for s in st])
        blob[f"{tag}_mask_sum"] = np.array(int(am.sum()))
        blob[f"{tag}_loss"] = loss.float().numpy()
        keys = sorted(logged, key=lambda k: int(k.rsplit("_", 1)[1]))
        blob[f"{tag}_logged_layers"] = np.array([int(k.rsplit("_", 1)[1]) for k in keys], dtype=np.int64)
        blob[f"{tag}_logged"] = np.array([logged[k] for k in keys], dtype=np.float64)
        sel = [j for j, g in enumerate(grads) if g is not None]
        blob[f"{tag}_grad_layers"] = np.array(sel, dtype=np.int64)
        blob[f"{tag}_grad_sum"] = np.array([float(grads[j].double().sum()) for j in sel])
        blob[f"{tag}_grad_norm"] = np.array([float(grads[j].double().norm()) for j in sel])
        blob[f"{tag}_grad_samples"] = np.stack([grads[j].reshape(-1)[pos].numpy() for j in sel])
        print(tag, float(loss), sel, blob[f"{tag}_grad_norm"][:3])
    blob["positions"] = sample_positions(CASE["bsz"] * (CASE["n_vis"] + CASE["txt"]) * CASE["dim"])
    np.savez_compressed(os.path.join(HERE, "c1_reference.npz"), **blob)


if __name__ == "__main__":
    main()
