#!/usr/bin/env python
"""A/B: one-pass step with the scale table derived in-kernel vs a separate prologue launch."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mafed_b200 import cabi
from mafed_b200.distill_op import distill_backward, distill_fused

lib = cabi.load()
dev = torch.device("cuda", 0)
for wl in ("C2", "C4", "C3"):
    desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
    st, te, am = bench.make_device_inputs(wl, 0, dev)
    fd = bench.make_method(n_sel)
    layers = list(range(n_sel))
    coeffs, kind, lang = fd._tables(layers)
    plan = fd._plan(layers, coeffs, 1.0, kind, lang)
    grads = [torch.empty_like(s) for s in st]
    gout = torch.ones((), device=dev)

    def one():
        out, scale, ln = distill_fused(st, te, grads, am, plan, group=False)
        distill_backward(ln, grads, scale, gout, skip_if_equals=1.0)

    for rep in range(2):
        for no_inline in (0, 1):
            lib.mafed_distill_set_tuning(16, no_inline)
            for _ in range(10):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(200):
                one()
            e1.record()
            host = (time.perf_counter() - t0) / 200 * 1e6
            torch.cuda.synchronize()
            print(wl, "prologue-launch" if no_inline else "inline-scale   ", f"{e0.elapsed_time(e1) / 200:.4f} ms/step  host {host:.0f} us", flush=True)
    lib.mafed_distill_set_tuning(16, 0)
    del st, te, grads
    torch.cuda.empty_cache()
