#!/usr/bin/env python
"""API-level (FeatureDistillation.distill + backward) step time per workload: the one-launch step (loss algebra in
the fused kernel's last CTA, masks written by the same kernel) vs separate mask / epilogue launches."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mafed_b200 import cabi
lib = cabi.load()
dev = torch.device("cuda", 0)
for wl in ("C4", "C2", "C3", "C1"):
    desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
    st, te, am = bench.make_device_inputs(wl, 0, dev)
    fd = bench.make_method(n_sel)
    fd.populate_batch_masks = True
    fd.past_model = lambda **kw: bench.Out(tuple(te))
    leaves = [s.detach().requires_grad_(True) for s in st]

    def step():
        for s in leaves:
            s.grad = None
        loss = fd.distill(bench.Out(tuple(leaves)), {"attention_mask": am})
        loss.backward()

    for rep in range(3):
        for no_tail in (0, 1):
            lib.mafed_distill_set_tuning(cabi.TUNE_NO_TAIL, no_tail)
            for _ in range(20):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(200):
                step()
            e1.record()
            host = (time.perf_counter() - t0) / 200 * 1e6
            torch.cuda.synchronize()
            print(wl, "separate-launches" if no_tail else "one-launch       ", f"{e0.elapsed_time(e1) / 200:.4f} ms/step  host {host:.0f} us", flush=True)
    lib.mafed_distill_set_tuning(cabi.TUNE_NO_TAIL, 0)
    del st, te, leaves, fd
    torch.cuda.empty_cache()
