#!/usr/bin/env python
"""Where the host time of one API step goes (distill + backward), on a workload small enough that the loop is
CPU-bound: wall-clock of the pieces in isolation, then a cProfile of the whole step."""
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


def timeit(fn, n=300, sync_every=50):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(n):
        t0 = time.perf_counter()
        fn()
        tot += time.perf_counter() - t0
        if i % sync_every == sync_every - 1:
            torch.cuda.synchronize()     # keep the launch queue short: measure enqueue cost, not back-pressure
    torch.cuda.synchronize()
    return 1e6 * tot / n


def main():
    from mafed_b200 import cabi, node
    dev = torch.device("cuda", 0)
    wl = sys.argv[1] if len(sys.argv) > 1 else "C1"
    desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
    st, te, am = bench.make_device_inputs(wl, 0, dev)
    fd = bench.make_method(n_sel)
    fd.past_model = lambda **kw: bench.Out(tuple(te))
    leaves = [s.detach().requires_grad_(True) for s in st]
    out = bench.Out(tuple(leaves))
    batch = {"attention_mask": am}

    def step():
        for s in leaves:
            s.grad = None
        loss = fd.distill(out, batch)
        loss.backward()

    print(f"{wl}: {timeit(step):.1f} us of host time per step (distill + backward, enqueue only)")

    with torch.autograd.set_multithreading_enabled(False):
        print(f"  the same step with torch.autograd.set_multithreading_enabled(False) (engine runs the nodes on the "
              f"calling thread, no hand-off to the device thread): {timeit(step):.1f} us")

    # marginal cost inside a larger backward: a step whose loss also has another term over the same leaves, with and
    # without the distillation term
    def other_only():
        for s in leaves:
            s.grad = None
        (leaves[0][0, 0, 0] * 2.0).backward()

    def other_plus_distill():
        for s in leaves:
            s.grad = None
        (leaves[0][0, 0, 0] * 2.0 + fd.distill(out, batch)).backward()
    a, b = timeit(other_only), timeit(other_plus_distill)
    print(f"  marginal: (other + distill).backward() {b:.1f} us - other.backward() {a:.1f} us = {b - a:.1f} us")

    def fwd_only():
        fd.distill(out, batch)
    print(f"  fd.distill() alone (graph dropped, no backward):     {timeit(fwd_only):.1f} us")
    ext = node.load()
    layers = list(range(n_sel))
    plan = fd._step_plan(layers)
    nplan = plan.node_plan()
    both = torch.empty((2, B, 256 + txt), dtype=torch.int64, device=dev)

    def raw_node():
        ext.distill(nplan, leaves, te, am, both, 0, None, None, 0)
    print(f"  ext.distill() alone (the compiled node, forward):    {timeit(raw_node):.1f} us")
    with torch.no_grad():
        print(f"  ext.distill() under no_grad (two-pass forward only): {timeit(raw_node):.1f} us")
    losses = []

    def fwd_keep():
        losses.append(fd.distill(out, batch))
    n = 300
    for s in leaves:
        s.grad = None
    for _ in range(n):
        fwd_keep()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for l in losses:
        l.backward()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"  loss.backward() alone (engine + node + {n_sel} AccumulateGrad, grads accumulate): {1e6 * (t1 - t0) / n:.1f} us")
    x = torch.zeros((), device=dev, requires_grad=True)

    def trivial():
        (x * 2.0).backward()
    print(f"  for scale: (x * 2).backward() on a CUDA scalar:      {timeit(trivial):.1f} us")

    def clear():
        for s in leaves:
            s.grad = None
    print(f"  clearing {n_sel} .grad fields:                              {timeit(clear):.1f} us")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        step()
    pr.disable()
    torch.cuda.synchronize()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(25)
    print(buf.getvalue())


if __name__ == "__main__":
    main()
