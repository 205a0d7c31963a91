#!/usr/bin/env python
"""A/B of two builds of libmafed_distill.so on the same box: kernel-level and API step times of a list of points,
alternating processes (MAFED_B200_LIB selects the build), medians over the rounds.

    python tools/ab_lib.py gpurun_ab/libmafed_distill_old.so [rounds] [points]

points: ';'-separated  n_layers:B:txt:D:dtype:mask:loss   (default: the 1B shape, short / long text, all-ones / ragged)
"""
import collections
import json
import os
import statistics
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = ("15:64:32:2048:bf16:ones:mse;15:64:32:2048:bf16:ragged:mse;15:64:256:2048:bf16:ones:mse;"
           "15:64:256:2048:bf16:ragged:mse;15:128:256:2048:bf16:ragged:mse;15:64:256:2048:bf16:ragged:cosine;"
           "11:128:32:768:bf16:ones:mse;11:128:256:768:bf16:ragged:mse")


def child(points):
    import torch
    sys.path.insert(0, ROOT)
    import bench
    ctx = bench.Ctx(types.SimpleNamespace())
    for pt in points.split(";"):
        n, B, txt, D, dt, mask, loss = pt.split(":")
        rec = bench.measure_point(ctx, int(n), int(B), int(txt), int(D), dt, mask == "ragged", 100, 20, loss=loss,
                                  kernel_level=True, graphed=False)
        print(json.dumps({"point": pt, "api_ms": rec["ms_per_step"], "kernel_ms": rec["kernel_level_ms_per_step"],
                          "frac": rec["frac"]}), flush=True)
        torch.cuda.empty_cache()


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2])
    old = os.path.abspath(sys.argv[1])
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    points = sys.argv[3] if len(sys.argv) > 3 else DEFAULT
    res = collections.defaultdict(list)
    for _ in range(rounds):
        for name, lib in (("old", old), ("new", None)):
            env = dict(os.environ)
            if lib:
                env["MAFED_B200_LIB"] = lib
            else:
                env.pop("MAFED_B200_LIB", None)
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", points], env=env,
                                 capture_output=True, text=True)
            if out.returncode != 0:
                print(out.stderr[-2000:], flush=True)
            for line in out.stdout.splitlines():
                if line.startswith("{"):
                    d = json.loads(line)
                    res[(d["point"], name)].append((d["kernel_ms"], d["api_ms"]))
    for pt in points.split(";"):
        o, n = res[(pt, "old")], res[(pt, "new")]
        if not o or not n:
            continue
        ko, kn = statistics.median(x[0] for x in o), statistics.median(x[0] for x in n)
        ao, an = statistics.median(x[1] for x in o), statistics.median(x[1] for x in n)
        print(f"{pt:40s} kernel old {ko:.4f} new {kn:.4f} new/old {kn / ko:.4f} | api old {ao:.4f} new {an:.4f} "
              f"new/old {an / ao:.4f}", flush=True)


if __name__ == "__main__":
    main()
