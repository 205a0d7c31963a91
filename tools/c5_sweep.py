#!/usr/bin/env python
"""C5 (BASELINE.json config 5) on one GPU: VLPythia-1B distillation over per-GPU batch 8..128 and
visual:text ratio 8:1..1:1 (txt 32..256), all-ones and ragged (left-padded) masks, one-pass and two-pass
steps at kernel level.  Padded rows count only their zero-fill write in the algorithmic bytes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mafed_b200.distill_op import distill_backward, distill_forward, distill_fused  # noqa: E402


def time_ms(fn, iters=30, warm=5):
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
    dev = torch.device("cuda", 0)
    L, D = 15, 2048
    out = []
    fd = bench.make_method(L)
    layers = list(range(L))
    coeffs, kind, lang = fd._tables(layers)
    plan = fd._plan(layers, coeffs, 1.0, kind, lang)
    gout = torch.ones((), device=dev)
    for B in (8, 16, 32, 64, 128):
        for txt in (32, 64, 128, 256):
            T = 256 + txt
            g = torch.Generator(device=dev).manual_seed(1)
            st = [torch.randn(B, T, D, generator=g, device=dev).to(torch.bfloat16) for _ in range(L)]
            te = [(s.float() + 0.1 * torch.randn(B, T, D, generator=g, device=dev)).to(torch.bfloat16) for s in st]
            grads = [torch.empty_like(s) for s in st]
            for mask_kind in ("ones", "ragged"):
                am = torch.ones(B, txt, dtype=torch.int64, device=dev)
                if mask_kind == "ragged":
                    for b in range(B):
                        am[b, : txt - (1 + (7 * b) % txt)] = 0
                valid = B * 256 + int(am.sum())
                padded = B * T - valid

                def one():
                    o, s, l = distill_fused(st, te, grads, am, plan, group=False)
                    distill_backward(l, grads, s, gout, skip_if_equals=1.0)

                def two():
                    o, s, l = distill_forward(st, te, am, plan, group=False)
                    distill_backward(l, grads, s, gout)

                row = D * 2
                for mode, fn, per_valid in (("one", one, 3), ("two", two, 5)):
                    ms = time_ms(fn)
                    nbytes = L * row * (per_valid * valid + padded)
                    out.append(dict(B=B, txt=txt, mask=mask_kind, mode=mode, ms=ms, units_per_s=B * T * L / ms * 1e3,
                                    gbs=nbytes / ms / 1e6, valid_frac=valid / (B * T)))
                    print(json.dumps(out[-1]), flush=True)
            del st, te, grads
            torch.cuda.empty_cache()
    with open(os.path.join(ROOT, "gpurun_out", "c5_sweep.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
