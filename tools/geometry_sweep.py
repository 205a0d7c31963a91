#!/usr/bin/env python
"""Ring geometry of the fused TMA kernel (stage size x stages) against the kernel-level step time (fused kernel +
gate, back to back), on the library MAFED_B200_LIB selects (default: the in-tree build).

    python tools/geometry_sweep.py [rounds] [points] [geometries] [fused|fwd|bwd]

points: ';'-separated n_layers:B:txt:D:dtype:mask:loss; geometries: ','-separated KBxSTAGES[@PACE_NS] (stage size in KB
of student + teacher rows; 0x0 = the library's default; PACE_NS = the producer's sleep before a refill).
"""
import json
import os
import statistics
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mafed_b200 import cabi  # noqa: E402
from mafed_b200.distill_op import distill_backward, distill_forward, distill_fused  # noqa: E402

POINTS = ("15:64:32:2048:bf16:ones:mse;15:64:32:2048:bf16:ragged:mse;15:64:256:2048:bf16:ones:mse;"
          "15:64:256:2048:bf16:ragged:mse;15:16:32:2048:bf16:ones:mse;15:128:128:2048:bf16:ragged:mse;"
          "11:128:32:768:bf16:ones:mse;11:128:256:768:bf16:ragged:mse;23:256:32:1024:bf16:ones:mse;"
          "15:64:32:2048:fp32:ones:mse;11:128:32:768:fp32:ones:mse;15:64:32:2048:bf16:ones:cosine;"
          "11:128:32:768:bf16:ones:cosine")
GEOS = "0x0,64x2,64x3,48x2,48x3,32x2,32x3,32x4,24x3,24x4,24x5,16x4,16x6"
FUSED = 2       # pass index of the one-pass kernel


def time_ms(fn, iters, warm):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    points = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] else POINTS
    geos = []
    for g in (sys.argv[3] if len(sys.argv) > 3 else GEOS).split(","):
        g, _, pace = g.partition("@")                   # KBxSTAGES[@pace_ns]
        kb, stages = (int(v) for v in g.split("x"))
        geos.append((kb, stages, int(pace or 0)))
    which = sys.argv[4] if len(sys.argv) > 4 else "fused"
    pass_index = {"fwd": 0, "bwd": 1, "fused": 2}[which]
    ctx = bench.Ctx(types.SimpleNamespace())
    dev = ctx.device
    for pt in points.split(";"):
        n, B, txt, D, dt, mask, loss = pt.split(":")
        n, B, txt, D = int(n), int(B), int(txt), int(D)
        dtype = bench.torch_dtype(dt)
        row_bytes = D * (4 if dt == "fp32" else 2)
        st, te = bench.synth(n, B, bench.N_VIS + txt, D, dtype, dev, 77)
        masks = bench.make_masks(B, txt, dev, mask == "ragged")
        fd = bench.make_method(n, loss=loss)
        layers = list(range(n))
        coeffs, kind, lang = fd._tables(layers)
        plan = fd._plan(layers, coeffs, fd.distillation_coeff, kind, lang)
        grads = [torch.empty_like(s) for s in st]
        gout = torch.ones((), dtype=torch.float32, device=dev)

        held = []

        def step(i):
            if which == "fused":
                out, scale, ln = distill_fused(st, te, grads, masks[i % 2], plan, group=False)
                distill_backward(ln, grads, scale, gout, skip_if_equals=plan.assumed_grad_out * plan.grad_multiplier,
                                 grad_out_scale=plan.grad_multiplier)
            elif which == "fwd":
                distill_forward(st, te, masks[i % 2], plan, group=False)
            else:
                if not held:
                    held.append(distill_forward(st, te, masks[0], plan, group=False))
                out, scale, ln = held[0]
                distill_backward(ln, grads, scale, gout, grad_out_scale=plan.grad_multiplier)
        est_ms = 3.0 * n * B * (bench.N_VIS + txt) * row_bytes / 6.5e9
        iters = max(20, min(100, int(40.0 / est_ms)))
        res = {}
        for _ in range(rounds):
            for kb, stages, pace in geos:
                rows = max(1, kb * 1024 // (2 * row_bytes)) if kb else 0
                with cabi.tuning(raw={cabi.TUNE_TMA_ROWS + pass_index: rows, cabi.TUNE_TMA_STAGES + pass_index: stages,
                                      cabi.TUNE_PACE_NS: pace}):
                    held.clear()        # (the launch record carries the tuning it was made under)
                    res.setdefault((kb, stages, rows, pace), []).append(time_ms(step, iters, 8))
        best = min(statistics.median(v) for v in res.values())
        row = {"point": pt, "pass": which, "best_ms": round(best, 4)}
        for (kb, stages, rows, pace), v in res.items():
            row[f"{kb}x{stages}(r{rows})" + (f"@{pace}" if pace else "")] = round(statistics.median(v) / best, 3)
        print(json.dumps(row), flush=True)
        del st, te, grads
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
