#!/usr/bin/env python
"""A/B of two builds of libmafed_distill.so on the same box: the kernel-level one-pass step at the library's
default geometry, alternating processes (MAFED_B200_LIB selects the build).

    python tools/ab_lib.py gpurun_ab/libmafed_distill_old.so [rounds]
"""
import collections, json, os, statistics, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
old = os.path.abspath(sys.argv[1])
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
points = "C2:mse:0,0,0,0;C2:cosine:0,0,0,0;C1:mse:0,0,0,0;C4:mse:0,0,0,0;C4:cosine:0,0,0,0;C3:mse:0,0,0,0"
res = collections.defaultdict(list)
for r in range(rounds):
    for name, lib in (("old", old), ("new", None)):
        env = dict(os.environ)
        if lib:
            env["MAFED_B200_LIB"] = lib
        else:
            env.pop("MAFED_B200_LIB", None)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep_ring.py"), "--workloads", "C2,C1,C4,C3",
                              "--iters", "200", "--repeats", "2", "--points", points], env=env, capture_output=True, text=True)
        for line in out.stdout.splitlines():
            if line.startswith("{"):
                d = json.loads(line)
                if "ms" in d:
                    res[(d["workload"], d["loss"], name)].append(d["ms"])
for (wl, loss) in sorted({k[:2] for k in res}):
    o, n = res[(wl, loss, "old")], res[(wl, loss, "new")]
    print(f"{wl} {loss:6s} old median {statistics.median(o):.4f} (min {min(o):.4f})  new median {statistics.median(n):.4f} "
          f"(min {min(n):.4f})  new/old {statistics.median(n) / statistics.median(o):.4f}", flush=True)
