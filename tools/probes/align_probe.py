#!/usr/bin/env python
"""Does the step time depend on where the student / teacher / gradient tensors sit relative to each other?
Kernel-level one-pass step at the 1B shape, default geometry, tensors carved out of one arena with a chosen skew
between the three roles (and between layers).

    python tools/probes/align_probe.py [B] [txt]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mafed_b200.distill_op import distill_backward, distill_fused  # noqa: E402

dev = torch.device("cuda", 0)


def time_ms(fn, iters=60, warm=10):
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    txt = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    L, D, T = 15, 2048, 256 + txt
    n = B * T * D
    layer_bytes = n * 2
    fd = bench.make_method(L)
    layers = list(range(L))
    coeffs, kind, lang = fd._tables(layers)
    plan = fd._plan(layers, coeffs, fd.distillation_coeff, kind, lang)
    am = torch.ones(B, txt, dtype=torch.int64, device=dev)
    gout = torch.ones((), dtype=torch.float32, device=dev)
    src = torch.randn(n, device=dev, dtype=torch.float32).to(torch.bfloat16)
    noise = (0.1 * torch.randn(n, device=dev, dtype=torch.float32)).to(torch.bfloat16)
    cases = [("separate torch allocations", None)]
    KB, MB = 1024, 1 << 20
    for role_skew, layer_gap in ((0, 0), (4 * KB, 0), (64 * KB, 0), (MB, 0), (MB + 4 * KB, 0), (37 * KB, 0), (512, 0),
                                 (0, 4 * KB), (0, 64 * KB), (0, MB + 4 * KB), (37 * KB, 53 * KB), (3 * MB + 28 * KB, 0)):
        cases.append((f"arena role_skew={role_skew} layer_gap={layer_gap}", (role_skew, layer_gap)))
    for rounds in range(2):
        for name, spec in cases:
            if spec is None:
                st = [src.clone().view(B, T, D) for _ in range(L)]
                te = [(src + noise).view(B, T, D) for _ in range(L)]
                gr = [torch.empty(B, T, D, device=dev, dtype=torch.bfloat16) for _ in range(L)]
                arena = None
            else:
                role_skew, gap = spec
                pitch = layer_bytes + gap
                role_bytes = L * pitch
                arena = torch.empty(3 * role_bytes + 2 * role_skew + 4096, dtype=torch.uint8, device=dev)

                def carve(role, l):
                    off = role * (role_bytes + role_skew) + l * pitch
                    return arena[off:off + layer_bytes].view(torch.bfloat16).view(B, T, D)
                st = [carve(0, l) for l in range(L)]
                te = [carve(1, l) for l in range(L)]
                gr = [carve(2, l) for l in range(L)]
                for l in range(L):
                    st[l].view(-1).copy_(src)
                    te[l].view(-1).copy_(src + noise)

            def step():
                out, scale, ln = distill_fused(st, te, gr, am, plan, group=False)
                distill_backward(ln, gr, scale, gout, skip_if_equals=plan.assumed_grad_out * plan.grad_multiplier,
                                 grad_out_scale=plan.grad_multiplier)
            ms = time_ms(step)
            print(json.dumps({"case": name, "round": rounds, "ms": round(ms, 4),
                              "s0": hex(st[0].data_ptr()), "t0": hex(te[0].data_ptr()), "g0": hex(gr[0].data_ptr())}), flush=True)
            del st, te, gr, arena
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
