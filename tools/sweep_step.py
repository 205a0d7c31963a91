#!/usr/bin/env python
"""Steady-state step sweep: times whole steps (one-pass: fused kernel + fix-up; two-pass: forward kernel +
backward kernel; the scalar stages run inside the kernels' last CTA) back to back, >= 100 iterations per point, every point measured `--repeats` times in
shuffled order, over TMA geometry and L2 eviction hints.

    python tools/sweep_step.py [--workloads C4,C2,C3] [--iters 100] [--repeats 2] [--mode one,two]
"""
import argparse
import itertools
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mafed_b200 import cabi  # noqa: E402
from mafed_b200.distill_op import distill_backward, distill_forward, distill_fused  # noqa: E402


def time_ms(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="C4,C2,C3")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--mode", default="one,two")
    ap.add_argument("--out", default=None)
    ap.add_argument("--extra", action="store_true", help="only the 2-CTAs/SM and 16-consumer-warp points")
    args = ap.parse_args()
    lib = cabi.load()
    dev = torch.device("cuda", 0)
    results = []

    def reset():
        for k in range(cabi.N_TUNE_KEYS):
            lib.mafed_distill_set_tuning(k, 0)
        lib.mafed_distill_set_variant(0)

    for wl in args.workloads.split(","):
        desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
        st, te, am = bench.make_device_inputs(wl, 0, dev)
        fd = bench.make_method(n_sel)
        layers = list(range(n_sel))
        coeffs, kind, lang = fd._tables(layers)
        plan = fd._plan(layers, coeffs, 1.0, kind, lang)
        grads = [torch.empty_like(s) for s in st]
        gout = torch.ones((), device=dev)
        row_bytes = D * st[0].element_size()
        units = B * (256 + txt) * n_sel

        def one():
            out, scale, ln = distill_fused(st, te, grads, am, plan, group=False)
            distill_backward(ln, grads, scale, gout, skip_if_equals=1.0)

        def two():
            out, scale, ln = distill_forward(st, te, am, plan, group=False)
            distill_backward(ln, grads, scale, gout)

        max_rows = max(1, min(32, (200 * 1024) // (2 * row_bytes)))
        geoms = [("ldg", 0, 0)]
        for rows in sorted({r for r in (4, 8, 16, 24, 32) if r <= max_rows}):
            for stages in (1, 2, 3, 4, 6):
                if stages * rows * 2 * row_bytes <= 208 * 1024 and stages * rows * 2 * row_bytes >= 96 * 1024:
                    geoms.append(("tma", rows, stages))
        points = []
        if "one" in args.mode:
            for g in geoms:
                points.append(("one", g, 0, 0))
            for lp, sp in itertools.product((0, 1, 3), (0, 1, 2, 3)):
                if (lp, sp) != (0, 0):
                    points.append(("one", ("tma", 0, 0), lp, sp))
                    points.append(("one", ("ldg", 0, 0), lp, sp))
        if "two" in args.mode:
            for g in geoms:
                points.append(("two", g, 0, 0))
            for lp, sp in ((1, 0), (0, 1), (1, 1), (2, 1), (1, 2)):
                points.append(("two", ("tma", 0, 0), lp, sp))
        if args.extra:
            points = []
            for mode in args.mode.split(","):
                points.append((mode, ("tma", 0, 0), 0, 0, 1, 8))            # new default
                for rows, stages in ((2, 2), (4, 2), (2, 3), (8, 2), (4, 3)):
                    if stages * rows * 2 * row_bytes <= 104 * 1024:
                        points.append((mode, ("tma", rows, stages), 0, 0, 2, 8))   # two CTAs per SM
                points.append((mode, ("tma", 0, 0), 0, 0, 1, 16))           # 16 consumer warps
                points.append((mode, ("tma", 0, 0), 0, 1, 1, 8))            # stores evict_first
                points.append((mode, ("tma", 0, 0), 2, 1, 1, 8))
        else:
            points = [pt + (1, 8) for pt in points]
        points = points * args.repeats
        random.Random(0).shuffle(points)
        for mode, (vname, rows, stages), lp, sp, gridmul, ncw in points:
            reset()
            for pid in ((cabi.PASS_FUSED,) if mode == "one" else (cabi.PASS_BWD,)):
                # geometry applies to the write-carrying kernel; the two-pass forward keeps its default
                lib.mafed_distill_set_tuning(cabi.TUNE_VARIANT + pid, 1 if vname == "ldg" else 2)
                if rows:
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_ROWS + pid, rows)
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_STAGES + pid, stages)
            lib.mafed_distill_set_tuning(cabi.TUNE_GRID_MUL, gridmul)
            lib.mafed_distill_set_tuning(cabi.TUNE_TMA_WARPS, ncw)
            lib.mafed_distill_set_tuning(cabi.TUNE_LOAD_POLICY, lp)
            lib.mafed_distill_set_tuning(cabi.TUNE_STORE_POLICY, sp)
            try:
                ms = time_ms(one if mode == "one" else two, args.iters)
                nbytes = (3 if mode == "one" else 5) * row_bytes * units
                rec = dict(workload=wl, mode=mode, variant=vname, rows=rows, stages=stages, load_policy=lp,
                           store_policy=sp, gridmul=gridmul, ncw=ncw, ms=ms, gbs=nbytes / ms / 1e6, units_per_s=units / ms * 1e3)
            except Exception as exc:
                rec = dict(workload=wl, mode=mode, variant=vname, rows=rows, stages=stages, load_policy=lp,
                           store_policy=sp, error=repr(exc))
            results.append(rec)
            print(json.dumps(rec), flush=True)
        del st, te, grads
        torch.cuda.empty_cache()
    reset()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
