#!/usr/bin/env python
"""Distribution of GPU-side step-to-step intervals of the API loop (distill + backward) vs the kernel-level loop:
tells a uniformly slower step (GPU-side cause) from occasional bubbles (host-side cause)."""
import os, statistics, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mafed_b200.distill_op import distill_backward, distill_fused
dev = torch.device("cuda", 0)
for wl in (sys.argv[1:] or ["C2", "C4"]):
    desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
    st, te, am = bench.make_device_inputs(wl, 0, dev)
    fd = bench.make_method(n_sel)
    fd.populate_batch_masks = True
    fd.past_model = lambda **kw: bench.Out(tuple(te))
    leaves = [s.detach().requires_grad_(True) for s in st]
    layers = list(range(n_sel))
    coeffs, kind, lang = fd._tables(layers)
    plan = fd._plan(layers, coeffs, 1.0, kind, lang)
    grads = [torch.empty_like(s) for s in st]
    gout = torch.ones((), device=dev)

    def api():
        for s in leaves:
            s.grad = None
        loss = fd.distill(bench.Out(tuple(leaves)), {"attention_mask": am})
        loss.backward()

    def kernel_level():
        out, scale, ln = distill_fused(st, te, grads, am, plan, group=False)
        distill_backward(ln, grads, scale, gout, skip_if_equals=1.0)

    for name, fn in (("kernel-level", kernel_level), ("api", api), ("kernel-level", kernel_level), ("api", api)):
        for _ in range(20):
            fn()
        n = 400
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            evs[i].record()
            fn()
        evs[n].record()
        host = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
        iv = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n))
        med = statistics.median(iv)
        print(f"{wl} {name:12s} mean {sum(iv) / n:.4f} median {med:.4f} p10 {iv[n // 10]:.4f} p90 {iv[9 * n // 10]:.4f} "
              f"max {iv[-1]:.4f} intervals>1.2x median: {sum(1 for x in iv if x > 1.2 * med)}  host {host:.4f} ms/step", flush=True)
    del st, te, leaves, grads
    torch.cuda.empty_cache()
