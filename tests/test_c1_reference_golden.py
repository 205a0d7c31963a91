"""BASELINE.json configs[0] at FULL size against outputs of the unmodified reference
(tests/golden/c1_reference.npz, made by tests/golden/make_golden_c1.py): the oracle on the CPU (not gpu) and the
CUDA path through the mirrored API (gpu).  Inputs are regenerated from the seed; their digests are checked first."""
import os

import numpy as np
import pytest
import torch

from oracle import distill_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = dict(modality="balanced", layer_strategy="discounted", loss="mse", coeff=1.0, cls=False, gamma=0.5,
            num_hidden_layers=11, n_tuple=13, layer=None, grad_out=1.0, n_vis=256, txt=32, bsz=8, dim=768,
            lang_coeff=None)
FP32_TOL = 1e-5   # north_star: 1e-5 relative for fp32 inputs


def _golden():
    return np.load(os.path.join(HERE, "golden", "c1_reference.npz"))


def _inputs(z, tag):
    st, te, am = O.make_inputs(CASE["n_tuple"], CASE["bsz"], CASE["txt"], CASE["dim"], n_vis=CASE["n_vis"],
                               dtype=torch.float32, seed=1234, teacher="close", mask=tag)
    pos = z["positions"]
    got = np.stack([s.reshape(-1)[pos].numpy() for s in st])
    assert np.array_equal(got, z[f"{tag}_input_samples"]), "torch's CPU random stream differs from the one the golden file was made with"
    digest = np.array([[float(s.double().sum()), float(t.double().sum())] for s, t in zip(st, te)])
    assert np.allclose(digest, z[f"{tag}_input_digest"], rtol=1e-12)
    assert int(am.sum()) == int(z[f"{tag}_mask_sum"])
    return st, te, am


def _check(z, tag, loss, layer_losses, grads):
    pos = z["positions"]
    assert float(loss) == pytest.approx(float(z[f"{tag}_loss"]), rel=FP32_TOL)
    for l, want in zip(z[f"{tag}_logged_layers"], z[f"{tag}_logged"]):
        assert float(layer_losses[int(l)]) == pytest.approx(float(want), rel=FP32_TOL)
    sel = [int(x) for x in z[f"{tag}_grad_layers"]]
    assert sel == list(range(11))
    for j, l in enumerate(sel):
        g = grads[l].detach().float().cpu()
        assert float(g.double().norm()) == pytest.approx(float(z[f"{tag}_grad_norm"][j]), rel=FP32_TOL)
        want = torch.from_numpy(z[f"{tag}_grad_samples"][j])
        got = g.reshape(-1)[pos]
        assert float((got - want).double().norm() / want.double().norm()) < FP32_TOL
        assert float(g.double().sum()) == pytest.approx(float(z[f"{tag}_grad_sum"][j]), rel=1e-3, abs=1e-9)
    assert all(grads[l] is None for l in (11, 12))          # the reference leaves the last two entries untouched


@pytest.mark.parametrize("tag", ["ragged", "ones"])
def test_oracle_matches_reference_at_config0(tag):
    from golden_util import oracle_cfg
    z = _golden()
    st, te, am = _inputs(z, tag)
    ref = O.forward_backward(st, te, am, oracle_cfg(CASE))
    _check(z, tag, ref["loss"], {l: v for l, v in ref["layer_losses"].items()}, ref["grads"])


@pytest.mark.gpu
@pytest.mark.parametrize("single_pass", [True, False], ids=["one-pass", "two-pass"])
@pytest.mark.parametrize("tag", ["ragged", "ones"])
def test_cuda_path_matches_reference_at_config0(tag, single_pass):
    from gpu_util import run_product
    z = _golden()
    st, te, am = _inputs(z, tag)
    out = run_product(dict(CASE), st, te, am, single_pass=single_pass)
    layer_losses = {int(k.rsplit("_", 1)[1]): v for k, v in out["layer_dict"].items()}
    _check(z, tag, out["loss"], layer_losses, out["grads"])
