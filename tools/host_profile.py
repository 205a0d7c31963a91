#!/usr/bin/env python
"""cProfile of the host side of one API step (distill + backward) on the small C1 workload, where the
loop is CPU-bound: shows where the Python time per step goes."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    wl = sys.argv[1] if len(sys.argv) > 1 else "C1"
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

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{wl}: {1e6 * (t1 - t0) / 300:.1f} us of host time per step (enqueue only)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        step()
    pr.disable()
    torch.cuda.synchronize()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(35)
    print(buf.getvalue())


if __name__ == "__main__":
    main()
