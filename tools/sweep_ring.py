#!/usr/bin/env python
"""Steady-state one-pass step over ring granularity: CTAs per SM x consumer warps x rows per stage x stages,
mse and cosine, every point `--repeats` times in shuffled order.

    python tools/sweep_ring.py [--workloads C4,C2,C3] [--iters 100] [--repeats 3] [--out file.json]
"""
import argparse
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mafed_b200 import cabi  # noqa: E402
from mafed_b200.distill_op import DistillPlan, distill_backward, distill_fused  # noqa: E402


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
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--losses", default="mse,cosine")
    ap.add_argument("--out", default=None)
    ap.add_argument("--points", default=None,
                    help="explicit list instead of the grid: 'C4:mse:2,4,2,3;C4:cosine:2,4,4,2;...' (0,0,0,0 = default)")
    args = ap.parse_args()
    explicit = {}
    if args.points:
        for item in args.points.split(";"):
            wl_, loss_, pt = item.split(":")
            explicit.setdefault((wl_, loss_), []).append(tuple(int(x) for x in pt.split(",")))
    lib = cabi.load()
    dev = torch.device("cuda", 0)
    results = []

    def reset():
        for k in range(cabi.N_TUNE_KEYS):
            lib.mafed_distill_set_tuning(k, 0)

    for wl in args.workloads.split(","):
        desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
        st, te, am = bench.make_device_inputs(wl, 0, dev)
        fd = bench.make_method(n_sel)
        layers = list(range(n_sel))
        coeffs, kind, lang = fd._tables(layers)
        grads = [torch.empty_like(s) for s in st]
        gout = torch.ones((), device=dev)
        row_bytes = D * st[0].element_size()
        units = B * (256 + txt) * n_sel
        points = [(0, 0, 0, 0)]   # library default
        for gridmul in (1, 2, 3, 4):
            for ncw in (8, 16):   # (a 4-warp build was tried in profiles/r01b_sweep_ring_box*.json and dropped)
                if gridmul * (ncw + 1) * 32 > 1184:   # registers: ~90 per thread
                    continue
                for rows in (1, 2, 4, 8, 16):
                    for stages in (2, 3, 4, 6, 8):
                        per_sm = gridmul * stages * rows * 2 * row_bytes
                        if 64 * 1024 <= per_sm <= 160 * 1024 and gridmul * (stages * rows * 2 * row_bytes + 14 * 1024) <= 226 * 1024:
                            points.append((gridmul, ncw, rows, stages))
        for loss in args.losses.split(","):
            if explicit:
                if (wl, loss) not in explicit:
                    continue
                points = explicit[(wl, loss)]
            elif loss == "cosine" and wl != "C4":
                continue
            plan = DistillPlan(layers=layers, layer_coeffs=[float(c) for c in coeffs], modality_kind=kind,
                               lang_weights=lang, loss_kind=cabi.LOSS_MSE if loss == "mse" else cabi.LOSS_COSINE)

            def one():
                out, scale, ln = distill_fused(st, te, grads, am, plan, group=False)
                distill_backward(ln, grads, scale, gout, skip_if_equals=1.0)

            pts = points * args.repeats
            random.Random(0).shuffle(pts)
            for gridmul, ncw, rows, stages in pts:
                reset()
                if rows:
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_ROWS + cabi.PASS_FUSED, rows)
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_STAGES + cabi.PASS_FUSED, stages)
                    lib.mafed_distill_set_tuning(cabi.TUNE_GRID_MUL, gridmul)
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_WARPS, ncw)
                try:
                    ms = time_ms(one, args.iters)
                    rec = dict(workload=wl, loss=loss, gridmul=gridmul, ncw=ncw, rows=rows, stages=stages, ms=ms,
                               gbs=3 * row_bytes * units / ms / 1e6)
                except Exception as exc:
                    rec = dict(workload=wl, loss=loss, gridmul=gridmul, ncw=ncw, rows=rows, stages=stages, error=repr(exc))
                    torch.cuda.synchronize()
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
