"""Pin the CPU oracle against vectors produced by the unmodified reference (not gpu)."""
import pytest
import torch

from golden_util import case_id, load_cases, load_plans, oracle_cfg
from oracle import distill_oracle as O

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[case_id(c) for c in CASES])
def test_oracle_matches_reference_vectors(case):
    m = case["meta"]
    out = O.forward_backward(case["students"], case["teachers"], case["mask"], oracle_cfg(m), grad_out=m["grad_out"])
    # same torch ops in the same order -> bit-identical on the same torch build; allow 1e-6
    assert float(out["loss"]) == pytest.approx(case["loss"], rel=1e-6)
    sel = [i for i, g in enumerate(out["grads"]) if g is not None]
    assert sel == case["grad_layers"]
    for j, ref in zip(sel, case["grads"]):
        torch.testing.assert_close(out["grads"][j].float(), ref, rtol=1e-6, atol=1e-9)
    if not m["cls"]:
        assert sorted(out["layer_losses"]) == case["logged_layers"]
        for l, ref in zip(case["logged_layers"], case["logged"]):
            assert float(out["layer_losses"][l]) == pytest.approx(ref, rel=1e-6)


@pytest.mark.parametrize("case", [c for c in CASES], ids=[case_id(c) for c in CASES])
def test_closed_form_matches_reference_vectors(case):
    m = case["meta"]
    out = O.closed_form(case["students"], case["teachers"], case["mask"], oracle_cfg(m), grad_out=m["grad_out"])
    bf16 = m["dtype"] == "bf16"
    assert float(out["loss"]) == pytest.approx(case["loss"], rel=2e-5 if not bf16 else 1e-4)
    sel = [i for i, g in enumerate(out["grads"]) if g is not None]
    assert sel == case["grad_layers"]
    for j, ref in zip(sel, case["grads"]):
        g = torch.from_numpy(out["grads"][j]).float()
        err = (g - ref).norm() / ref.norm()
        assert err < (2e-3 if bf16 else 1e-5), float(err)


def test_layer_plans_match_reference():
    for rec in load_plans()["plans"]:
        cfg = O.OracleConfig(layer_strategy=rec["strategy"], gamma=rec["gamma"],
                             num_hidden_layers=rec["num_hidden_layers"], distillation_layer=rec["layer"])
        if "error" in rec:
            with pytest.raises(AssertionError):
                O.layer_plan(cfg)
            continue
        layers, coeffs, eff = O.layer_plan(cfg)
        assert layers == rec["layers"] and eff == rec["effective"]
        got = [1.0 if coeffs is None else float(coeffs[l]) for l in layers]
        assert got == pytest.approx(rec["coeffs"], rel=1e-7)


def test_cls_with_mse_raises_like_reference():
    st, te, am = O.make_inputs(3, 2, 4, 8, n_vis=8)
    cfg = O.OracleConfig(layer_strategy="equal", num_hidden_layers=2, distillation_layer=None, loss="mse",
                         cls_distillation=True, num_vision_tokens=8)
    with pytest.raises(TypeError):
        O.forward_backward(st, te, am, cfg)
