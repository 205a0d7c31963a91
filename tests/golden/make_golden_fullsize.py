"""Golden outputs of the UNMODIFIED reference at BASELINE.json's configurations -- full size.

    python tests/golden/make_golden_fullsize.py [C1 C2 C3 C4 C2cos]      (build container only: needs /root/reference)

C1 = configs[0] "VLPythia-base (12 layers, d=768) ... batch 8, 256 visual + 32 text tokens, fp32, on CPU";
C2 = configs[1] the same model at batch 128 in bf16; C3 = configs[2] VLPythia-410M (24 layers, d=1024), batch 256,
bf16; C4 = configs[3]'s per-GPU shard, VLPythia-1B (16 layers, d=2048), 64 samples, bf16; C2cos = the C2 shape with
the cosine token loss and the count-weighted ("equal") modality strategy.  Everywhere the reference's own call num_hidden_layers = L - 1 (train.py:133) and the
shipped recipe mse / balanced / discounted gamma 0.5 (scripts/run_seed42.sh:74-93); bf16 runs under
torch.autocast(bfloat16) as distillation.py:90 does (CPU autocast here).
The inputs (up to 2 x 1.1 GB) are not committed: they are regenerated from the seed with oracle.make_inputs (torch
CPU generator); each file stores digests of them so that a different random stream is noticed, the reference's loss
and logged per-layer losses, and per layer the gradient's sum, L2 norm and 256 entries at fixed positions.  Two mask
variants: the ragged one of SURVEY 8(d) and all-ones.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from make_golden import _install_stubs, run_reference  # noqa: E402

BASE = dict(modality="balanced", layer_strategy="discounted", loss="mse", coeff=1.0, cls=False, gamma=0.5, layer=None,
            grad_out=1.0, n_vis=256, txt=32, teacher="close", lang_coeff=None)
CONFIGS = {
    "C1": dict(BASE, num_hidden_layers=11, n_tuple=13, bsz=8, dim=768, dtype="fp32"),
    "C2": dict(BASE, num_hidden_layers=11, n_tuple=13, bsz=128, dim=768, dtype="bf16"),
    "C4": dict(BASE, num_hidden_layers=15, n_tuple=17, bsz=64, dim=2048, dtype="bf16"),
    # configs[2]: VLPythia-410M (24 layers, d=1024), batch 256, bf16 (the +ER part is the LM loss, outside the path)
    "C3": dict(BASE, num_hidden_layers=23, n_tuple=25, bsz=256, dim=1024, dtype="bf16", tags=("ragged",)),
    # the other loss / modality strategy at full size: cosine token loss, count-weighted ("equal") modalities
    "C2cos": dict(BASE, num_hidden_layers=11, n_tuple=13, bsz=128, dim=768, dtype="bf16", loss="cosine", modality="equal",
                  tags=("ragged",)),
}
SEED = 1234
N_SAMPLES = 256


def sample_positions(numel):
    rng = np.random.default_rng(7)
    return np.sort(rng.choice(numel, size=N_SAMPLES, replace=False)).astype(np.int64)


def main():
    _install_stubs()
    from oracle.distill_oracle import make_inputs
    for name in (sys.argv[1:] or list(CONFIGS)):
        case = CONFIGS[name]
        dtype = torch.bfloat16 if case["dtype"] == "bf16" else torch.float32
        blob = {}
        for tag in case.get("tags", ("ragged", "ones")):
            st, te, am = make_inputs(case["n_tuple"], case["bsz"], case["txt"], case["dim"], n_vis=case["n_vis"],
                                     dtype=dtype, seed=SEED, teacher=case["teacher"], mask=tag)
            loss, logged, grads = run_reference(case, st, te, am)
            pos = sample_positions(st[0].numel())
            blob[f"{tag}_input_digest"] = np.array([[float(s.double().sum()), float(t.double().sum())] for s, t in zip(st, te)])
            blob[f"{tag}_input_samples"] = np.stack([s.reshape(-1)[pos].float().numpy() for s in st])
            blob[f"{tag}_mask_sum"] = np.array(int(am.sum()))
            blob[f"{tag}_loss"] = loss.float().numpy()
            keys = sorted(logged, key=lambda k: int(k.rsplit("_", 1)[1]))
            blob[f"{tag}_logged_layers"] = np.array([int(k.rsplit("_", 1)[1]) for k in keys], dtype=np.int64)
            blob[f"{tag}_logged"] = np.array([logged[k] for k in keys], dtype=np.float64)
            sel = [j for j, g in enumerate(grads) if g is not None]
            assert all(grads[j].dtype == dtype for j in sel)
            blob[f"{tag}_grad_layers"] = np.array(sel, dtype=np.int64)
            blob[f"{tag}_grad_sum"] = np.array([float(grads[j].double().sum()) for j in sel])
            blob[f"{tag}_grad_norm"] = np.array([float(grads[j].double().norm()) for j in sel])
            blob[f"{tag}_grad_samples"] = np.stack([grads[j].reshape(-1)[pos].float().numpy() for j in sel])
            blob["positions"] = pos
            print(name, tag, float(loss), sel, blob[f"{tag}_grad_norm"][:3], flush=True)
            del st, te, grads
        np.savez_compressed(os.path.join(HERE, f"fullsize_{name}.npz"), **blob)


if __name__ == "__main__":
    main()
