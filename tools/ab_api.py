#!/usr/bin/env python
"""API-level (FeatureDistillation.distill + backward) step time per workload, inline scale on/off."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mafed_b200 import cabi
lib = cabi.load()
dev = torch.device("cuda", 0)
for wl in ("C2", "C1", "C4"):
    desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
    st, te, am = bench.make_device_inputs(wl, 0, dev)
    fd = bench.make_method(n_sel)
    fd.past_model = lambda **kw: bench.Out(tuple(te))
    leaves = [s.detach().requires_grad_(True) for s in st]

    def step():
        for s in leaves:
            s.grad = None
        loss = fd.distill(bench.Out(tuple(leaves)), {"attention_mask": am})
        loss.backward()

    for rep in range(2):
        for no_inline in (0, 1):
            lib.mafed_distill_set_tuning(16, no_inline)
            for _ in range(10):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(100):
                step()
            e1.record()
            host = (time.perf_counter() - t0) / 100 * 1e6
            torch.cuda.synchronize()
            print(wl, "prologue-launch" if no_inline else "inline-scale   ", f"{e0.elapsed_time(e1) / 100:.4f} ms/step  host {host:.0f} us", flush=True)
    lib.mafed_distill_set_tuning(16, 0)
    del st, te, leaves, fd
    torch.cuda.empty_cache()
