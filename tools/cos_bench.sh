#!/bin/bash
python - <<'PY'
import json, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import torch
from mafed_b200 import cabi
from mafed_b200.distill_op import DistillPlan, distill_backward, distill_forward, distill_fused
lib = cabi.load(); dev = torch.device("cuda", 0)
def time_ms(fn, iters=60, warm=8):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, L, B, D, dtype in [("1B bf16", 15, 64, 2048, torch.bfloat16), ("base bf16", 11, 128, 768, torch.bfloat16), ("1B fp32", 15, 32, 2048, torch.float32)]:
    g = torch.Generator(device=dev).manual_seed(1)
    st = [torch.randn(B, 288, D, generator=g, device=dev).to(dtype) for _ in range(L)]
    te = [(s.float() + 0.1 * torch.randn(B, 288, D, generator=g, device=dev)).to(dtype) for s in st]
    grads = [torch.empty_like(s) for s in st]
    am = torch.ones(B, 32, dtype=torch.int64, device=dev); gout = torch.ones((), device=dev)
    plan = DistillPlan(layers=list(range(L)), layer_coeffs=[1.0 / L] * L, modality_kind=cabi.MODW_EQUAL, loss_kind=cabi.LOSS_COSINE)
    units = B * 288 * L; es = st[0].element_size()
    for warps in (0, 16):
        lib.mafed_distill_set_tuning(cabi.TUNE_TMA_WARPS, warps)
        def one():
            o, s, l = distill_fused(st, te, grads, am, plan, group=False); distill_backward(l, grads, s, gout, skip_if_equals=1.0)
        def two():
            o, s, l = distill_forward(st, te, am, plan, group=False); distill_backward(l, grads, s, gout)
        m1, m2 = time_ms(one), time_ms(two)
        print(name, "cosine warps", warps or "auto(8)", f"one {m1:.4f} ms {3*D*es*units/m1/1e6:.0f} GB/s | two {m2:.4f} ms {5*D*es*units/m2/1e6:.0f} GB/s", flush=True)
    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_WARPS, 0)
    del st, te, grads
PY
