#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: times the fused forward, backward and one-pass kernels separately
(CUDA events, inputs >> L2) for each kernel family / geometry and prints achieved algorithmic GB/s.

    python tools/sweep.py [--workloads C4,C2,C3] [--iters 30] [--passes fwd,bwd,fused]
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
from mafed_b200.distill_op import distill_backward, distill_forward, distill_fused  # noqa: E402


def time_ms(fn, iters, warm=4):
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
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--passes", default="fwd,bwd,fused")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    lib = cabi.load()
    dev = torch.device("cuda", 0)
    results = []
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev).normal_()
    b = torch.empty_like(a)
    ms = time_ms(lambda: b.copy_(a), 10, 3)
    print(f"calib copy_ 2 GiB r+w: {2 * a.numel() * 2 / ms / 1e6:.0f} GB/s", flush=True)
    results.append(dict(kind="calib_copy", gbs=2 * a.numel() * 2 / ms / 1e6))
    del a, b

    def reset():
        for k in range(cabi.N_TUNE_KEYS):
            lib.mafed_distill_set_tuning(k, 0)
        lib.mafed_distill_set_variant(0)

    passes = args.passes.split(",")
    pass_id = dict(fwd=cabi.PASS_FWD, bwd=cabi.PASS_BWD, fused=cabi.PASS_FUSED)
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
        row_bytes = D * esize
        units = B * (256 + txt) * n_sel
        nbytes = dict(fwd=2 * row_bytes * units, bwd=3 * row_bytes * units, fused=3 * row_bytes * units)
        reset()
        out, scale, ln = distill_forward(st, te, am, plan, group=False)
        fns = dict(fwd=lambda: distill_forward(st, te, am, plan, group=False),
                   bwd=lambda: distill_backward(ln, grads, scale, gout),
                   fused=lambda: distill_fused(st, te, grads, am, plan, group=False))
        max_rows = max(1, min(32, (200 * 1024) // (2 * row_bytes)))
        row_opts = sorted({r for r in (2, 4, 8, 16, 24, 32) if r <= max_rows})
        for ps in passes:
            configs = [("ldg", 0, 0, 0), ("ldg", 0, 0, 2)]
            for rows in row_opts:
                for stages in (1, 2, 3, 4, 6):
                    if stages * rows * 2 * row_bytes <= 208 * 1024:
                        configs.append(("tma", rows, stages, 0))
            configs.append(("tma16w", 0, 0, 0))
            for vname, rows, stages, cap in configs:
                reset()
                pid = pass_id[ps]
                lib.mafed_distill_set_tuning(cabi.TUNE_VARIANT + pid, 1 if vname == "ldg" else 2)
                if vname == "tma16w":
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_WARPS, 16)
                if rows:
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_ROWS + pid, rows)
                    lib.mafed_distill_set_tuning(cabi.TUNE_TMA_STAGES + pid, stages)
                if cap:
                    lib.mafed_distill_set_tuning(cabi.TUNE_LDG_BLOCKS_PER_SM, cap)
                try:
                    ms = time_ms(fns[ps], args.iters)
                    rec = dict(workload=wl, kernel=ps, variant=vname, rows=rows, stages=stages, cap=cap, ms=ms,
                               gbs=nbytes[ps] / ms / 1e6)
                except Exception as exc:
                    rec = dict(workload=wl, kernel=ps, variant=vname, rows=rows, stages=stages, error=repr(exc))
                results.append(rec)
                print(json.dumps(rec), flush=True)
        del st, te, grads, ln
        torch.cuda.empty_cache()
    reset()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
