"""the raw ctypes stub a reference maintainer would write, run as written."""
import ctypes

import pytest
import torch

from oracle import distill_oracle as O

pytestmark = pytest.mark.gpu


def test_raw_ctypes_stub_two_pass_and_backward():
    from mafed_b200 import cabi
    lib = ctypes.CDLL(cabi.LIB_PATH)
    lib.mafed_distill_ws_bytes.restype = ctypes.c_size_t
    lib.mafed_distill_error_string.restype = ctypes.c_char_p

    class Shape(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int32) for n in ("n_layers", "B", "T", "n_vis", "D", "dtype", "loss_kind", "cls")] + \
                   [("tuning", ctypes.c_void_p)]

    class Weights(ctypes.Structure):
        _fields_ = [("modality_kind", ctypes.c_int32), ("distill_coeff", ctypes.c_float),
                    ("layer_coeff", ctypes.c_float * 64), ("lang_weight", ctypes.c_float * 64)]

    def ptrs(ts):
        return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])

    st, te, am = O.make_inputs(4, 3, 6, 768, n_vis=256, seed=51)
    cfg = O.OracleConfig(modality_strategy="balanced", layer_strategy="discounted", gamma=0.5, num_hidden_layers=3,
                         distillation_layer=None)
    ref = O.forward_backward(st, te, am, cfg, grad_out=0.25)
    layers, coeffs, _ = O.layer_plan(cfg)
    students = [s.cuda() for s in st[:3]]
    teachers = [t.cuda() for t in te[:3]]
    mask = am.cuda()
    L = 3
    B, T, D = students[0].shape
    sh = Shape(L, B, T, 256, D, 0, 0, 0, None)
    w = Weights(1, 1.0)
    w.layer_coeff[:L] = [float(c) for c in coeffs]
    w.lang_weight[:L] = [0.5] * L
    stream = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(lib.mafed_distill_ws_bytes(L), dtype=torch.uint8, device="cuda")
    out = torch.empty(1 + 3 * L, device="cuda")
    scale = torch.empty(2 * L, device="cuda")
    vp = ctypes.c_void_p
    rc = lib.mafed_distill_fwd(ctypes.byref(sh), ptrs(students), ptrs(teachers), vp(mask.data_ptr()),
                               vp(ws.data_ptr()), vp(stream))
    assert rc == 0, lib.mafed_distill_error_string(rc)
    REDUCE, COUNTS, LOSSES, SCALE = 1, 2, 4, 8
    rc = lib.mafed_distill_scalar_stage(ctypes.byref(sh), ctypes.byref(w), REDUCE | COUNTS | LOSSES | SCALE,
                                        vp(mask.data_ptr()), vp(ws.data_ptr()), None, vp(out.data_ptr()),
                                        vp(scale.data_ptr()), None, 0, vp(stream))
    assert rc == 0
    grads = [torch.empty_like(s) for s in students]
    gout = torch.tensor(0.25, device="cuda")
    rc = lib.mafed_distill_bwd(ctypes.byref(sh), ptrs(students), ptrs(teachers), ptrs(grads), vp(mask.data_ptr()),
                               vp(scale.data_ptr()), vp(gout.data_ptr()), ctypes.c_float(1.0), None, None, vp(stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert float(out[0]) == pytest.approx(float(ref["loss"]), rel=1e-5)
    for l in range(L):
        err = float((grads[l].cpu() - ref["grads"][l]).norm() / ref["grads"][l].norm())
        assert err < 1e-5
