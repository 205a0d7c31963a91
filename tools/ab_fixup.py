#!/usr/bin/env python
"""What the backward fix-up launch costs per step: loop of the fused kernel alone vs fused + fix-up."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mafed_b200.distill_op import distill_backward, distill_fused
dev = torch.device("cuda", 0)
for wl in ("C4", "C2", "C3"):
    desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
    st, te, am = bench.make_device_inputs(wl, 0, dev)
    fd = bench.make_method(n_sel)
    layers = list(range(n_sel))
    coeffs, kind, lang = fd._tables(layers)
    plan = fd._plan(layers, coeffs, 1.0, kind, lang)
    grads = [torch.empty_like(s) for s in st]
    gout = torch.ones((), device=dev)

    def fused_only():
        distill_fused(st, te, grads, am, plan, group=False)

    def with_fixup():
        out, scale, ln = distill_fused(st, te, grads, am, plan, group=False)
        distill_backward(ln, grads, scale, gout, skip_if_equals=1.0)

    for rep in range(4):
        for name, fn in (("fused only      ", fused_only), ("fused + fix-up  ", with_fixup)):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(wl, name, f"{e0.elapsed_time(e1) / 200:.4f} ms/step", flush=True)
    del st, te, grads
    torch.cuda.empty_cache()
