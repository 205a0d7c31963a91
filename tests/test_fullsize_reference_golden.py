"""BASELINE.json's configurations at FULL size against outputs of the unmodified reference
(tests/golden/fullsize_{C1,C2,C3,C4,C2cos}.npz, made by tests/golden/make_golden_fullsize.py): the oracle on the CPU
at C1 (not gpu) and the CUDA path through the mirrored API at C1 (configs[0], fp32), C2 (configs[1], bf16), C3
(configs[2], VLPythia-410M, batch 256, bf16), the C4 per-GPU shard (configs[3], bf16) and C2 with the cosine loss
and count-weighted modalities (gpu).  Inputs are regenerated from the seed; their digests are checked first."""
import os

import numpy as np
import pytest
import torch

from oracle import distill_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
BASE = dict(modality="balanced", layer_strategy="discounted", loss="mse", coeff=1.0, cls=False, gamma=0.5, layer=None,
            grad_out=1.0, n_vis=256, txt=32, lang_coeff=None)
CONFIGS = {
    "C1": dict(BASE, num_hidden_layers=11, n_tuple=13, bsz=8, dim=768, dtype=torch.float32),
    "C2": dict(BASE, num_hidden_layers=11, n_tuple=13, bsz=128, dim=768, dtype=torch.bfloat16),
    "C4": dict(BASE, num_hidden_layers=15, n_tuple=17, bsz=64, dim=2048, dtype=torch.bfloat16),
    "C3": dict(BASE, num_hidden_layers=23, n_tuple=25, bsz=256, dim=1024, dtype=torch.bfloat16),
    "C2cos": dict(BASE, num_hidden_layers=11, n_tuple=13, bsz=128, dim=768, dtype=torch.bfloat16, loss="cosine",
                  modality="equal"),
}
TAGS = {"C3": ("ragged",), "C2cos": ("ragged",)}     # configurations generated with one mask variant only


def _tol(case):
    """north_star: 1e-5 relative for fp32 inputs, 2e-3 for bf16 inputs."""
    return 1e-5 if case["dtype"] == torch.float32 else 2e-3


def _golden(name):
    return np.load(os.path.join(HERE, "golden", f"fullsize_{name}.npz"))


_TENSORS = {}


def _inputs(case, z, tag):
    # hidden states depend on the seed only (generated once per configuration); the mask on (batch, text length, kind)
    key = (case["n_tuple"], case["bsz"], case["dim"], case["dtype"])
    if key not in _TENSORS:
        _TENSORS.clear()   # one configuration's tensors at a time (up to 2 x 1.1 GB)
        _TENSORS[key] = O.make_inputs(case["n_tuple"], case["bsz"], case["txt"], case["dim"], n_vis=case["n_vis"],
                                      dtype=case["dtype"], seed=1234, teacher="close", mask="ones")[:2]
    st, te = _TENSORS[key]
    am = O.make_inputs(1, case["bsz"], case["txt"], 1, n_vis=case["n_vis"], seed=1234, mask=tag)[2]
    pos = z["positions"]
    got = np.stack([s.reshape(-1)[pos].float().numpy() for s in st])
    assert np.array_equal(got, z[f"{tag}_input_samples"]), \
        "torch's CPU random stream differs from the one the golden file was made with"
    digest = np.array([[float(s.double().sum()), float(t.double().sum())] for s, t in zip(st, te)])
    assert np.allclose(digest, z[f"{tag}_input_digest"], rtol=1e-12)
    assert int(am.sum()) == int(z[f"{tag}_mask_sum"])
    return st, te, am


def _check(case, z, tag, loss, layer_losses, grads):
    tol, pos, nh = _tol(case), z["positions"], case["num_hidden_layers"]
    assert float(loss) == pytest.approx(float(z[f"{tag}_loss"]), rel=tol)
    for l, want in zip(z[f"{tag}_logged_layers"], z[f"{tag}_logged"]):
        assert float(layer_losses[int(l)]) == pytest.approx(float(want), rel=tol)
    sel = [int(x) for x in z[f"{tag}_grad_layers"]]
    assert sel == list(range(nh))
    for j, l in enumerate(sel):
        g = grads[l].detach().float().cpu()
        assert float(g.double().norm()) == pytest.approx(float(z[f"{tag}_grad_norm"][j]), rel=tol)
        want = torch.from_numpy(z[f"{tag}_grad_samples"][j])
        got = g.reshape(-1)[pos]
        assert float((got - want).double().norm() / want.double().norm()) < tol
        if case["dtype"] == torch.float32:
            assert float(g.double().sum()) == pytest.approx(float(z[f"{tag}_grad_sum"][j]), rel=1e-3, abs=1e-9)
    assert all(grads[l] is None for l in range(nh, case["n_tuple"]))   # the reference leaves the tail untouched


@pytest.mark.parametrize("name,tag", [("C1", "ragged"), ("C1", "ones"), ("C2", "ragged"), ("C2cos", "ragged")])
def test_oracle_matches_reference_at_full_size(name, tag):
    """The CPU restatement against the unmodified reference at configs[0] (fp32) and configs[1] (bf16 autocast)."""
    from golden_util import oracle_cfg
    case, z = CONFIGS[name], _golden(name)
    st, te, am = _inputs(case, z, tag)
    ref = O.forward_backward(st, te, am, oracle_cfg(case))
    _check(case, z, tag, ref["loss"], dict(ref["layer_losses"]), ref["grads"])


@pytest.mark.gpu
@pytest.mark.parametrize("single_pass", [True, False], ids=["one-pass", "two-pass"])
@pytest.mark.parametrize("tag", ["ragged", "ones"])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_cuda_path_matches_reference_at_full_size(name, tag, single_pass):
    from gpu_util import run_product
    if tag not in TAGS.get(name, ("ragged", "ones")):
        pytest.skip("this configuration's golden holds the ragged mask only")
    case, z = CONFIGS[name], _golden(name)
    st, te, am = _inputs(case, z, tag)
    out = run_product(dict(case), st, te, am, single_pass=single_pass)
    layer_losses = {int(k.rsplit("_", 1)[1]): v for k, v in out["layer_dict"].items()}
    _check(case, z, tag, out["loss"], layer_losses, out["grads"])
