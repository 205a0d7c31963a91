#!/usr/bin/env python
"""Host<->device copy bandwidth of this box with pinned buffers (the bound of bench.py's `e2e` leg):
H2D alone, D2H alone, and both directions at once, in chunks of one C4 layer (75.5 MB)."""
import json
import torch

dev = torch.device("cuda", 0)
chunk = 64 * 288 * 2048 * 2
n = 16
h_in = torch.empty(n * chunk, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n * chunk, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
d_out = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for i in range(n):
        sl = slice(i * chunk, (i + 1) * chunk)
        if h2d:
            with torch.cuda.stream(s1):
                d_in[sl].copy_(h_in[sl], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


res = {}
for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    best = min(run(h2d, d2h) for _ in range(5))
    res[name + "_GBps_per_direction"] = n * chunk / best / 1e9
print(json.dumps(res))
