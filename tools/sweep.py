#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: times the fused forward and backward separately (CUDA events,
inputs >> L2) for each kernel family / tuning knob and prints achieved algorithmic GB/s.

    python tools/sweep.py [--workloads C4,C2] [--iters 30] [--quick]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mafed_b200 import cabi  # noqa: E402
from mafed_b200.distill_op import distill_backward, distill_forward  # noqa: E402


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
    ap.add_argument("--workloads", default="C4,C2")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    lib = cabi.load()
    dev = torch.device("cuda", 0)
    results = []

    # calibration: plain device copy and a read-only reduction through torch
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev).normal_()
    b = torch.empty_like(a)
    ms = time_ms(lambda: b.copy_(a), 10, 3)
    print(f"calib copy_ 2 GiB r+w: {2 * a.numel() * 2 / ms / 1e6:.0f} GB/s", flush=True)
    results.append(dict(kind="calib_copy", gbs=2 * a.numel() * 2 / ms / 1e6))
    del a, b

    configs = [("ldg", {}), ("tma", {})]
    if not args.quick:
        configs += [
            ("ldg", {cabi.TUNE_LDG_BLOCKS_PER_SM: 2}),
            ("tma", {cabi.TUNE_TMA_STAGES: 3}),
            ("tma", {cabi.TUNE_TMA_STAGES: 2}),
            ("tma", {cabi.TUNE_TMA_ROWS: 4, cabi.TUNE_TMA_STAGES: 6}),
            ("tma", {cabi.TUNE_TMA_ROWS: 16, cabi.TUNE_TMA_STAGES: 1 + 0}),
            ("tma", {cabi.TUNE_TMA_WARPS: 16}),
            ("tma", {cabi.TUNE_TMA_WARPS: 16, cabi.TUNE_TMA_STAGES: 3}),
            ("tma", {cabi.TUNE_BWD_REVERSE: 2}),
            ("ldg", {cabi.TUNE_BWD_REVERSE: 2}),
        ]
    for wl in args.workloads.split(","):
        desc, n_tuple, n_sel, B, txt, D, dt = bench.WORKLOADS[wl]
        st, te, am = bench.make_device_inputs(wl, 0, dev)
        fd = bench.make_method(n_sel)
        layers = list(range(n_sel))
        coeffs, kind, lang = fd._tables(layers)
        plan = fd._plan(layers, coeffs, 1.0, kind, lang)
        grads = [torch.empty_like(s) for s in st]
        gout = torch.ones((), device=dev)
        esize = st[0].element_size()
        units = B * (256 + txt) * n_sel
        fwd_bytes, bwd_bytes = 2 * D * esize * units, 3 * D * esize * units
        for vname, tune in configs:
            for k in range(8):
                lib.mafed_distill_set_tuning(k, 0)
            for k, v in tune.items():
                lib.mafed_distill_set_tuning(k, v)
            lib.mafed_distill_set_variant(dict(ldg=1, tma=2)[vname])
            try:
                out, scale, ln = distill_forward(st, te, am, plan, group=False)
                f_ms = time_ms(lambda: distill_forward(st, te, am, plan, group=False), args.iters)
                b_ms = time_ms(lambda: distill_backward(ln, grads, scale, gout), args.iters)

                def both():
                    o, s, l = distill_forward(st, te, am, plan, group=False)
                    distill_backward(l, grads, s, gout)
                s_ms = time_ms(both, args.iters)
                rec = dict(workload=wl, variant=vname, tune={str(k): v for k, v in tune.items()}, fwd_ms=f_ms,
                           bwd_ms=b_ms, step_ms=s_ms, fwd_gbs=fwd_bytes / f_ms / 1e6, bwd_gbs=bwd_bytes / b_ms / 1e6,
                           step_gbs=(fwd_bytes + bwd_bytes) / s_ms / 1e6, units_per_s=units / s_ms * 1e3,
                           loss=float(out[0]))
            except Exception as exc:
                rec = dict(workload=wl, variant=vname, tune={str(k): v for k, v in tune.items()}, error=repr(exc))
            results.append(rec)
            print(json.dumps(rec), flush=True)
        del st, te, grads
        torch.cuda.empty_cache()
    for k in range(8):
        lib.mafed_distill_set_tuning(k, 0)
    lib.mafed_distill_set_variant(0)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
