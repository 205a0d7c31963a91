#!/usr/bin/env python
"""C5 (BASELINE.json config 5) across the GPUs of one box, one process per GPU (run under torchrun):
VLPythia-1B distillation, GLOBAL batch 64..1024 sharded over the ranks, visual:text ratio 8:1 and 1:1
(txt 32 and 256), all-ones masks, one-pass step at kernel level with both cross-rank exchanges inside the fused
kernel.  Times are the max over ranks; units are those of all ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 \
        tools/c5_sweep_dist.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mafed_b200.distill_op import distill_backward, distill_fused  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L, D = 15, 2048
    fd = bench.make_method(L)
    layers = list(range(L))
    coeffs, kind, lang = fd._tables(layers)
    plan = fd._plan(layers, coeffs, 1.0, kind, lang)
    gout = torch.ones((), device=dev)
    out = []
    for global_b in (64, 128, 256, 512, 1024):
        B = global_b // world
        if B < 1:
            continue
        for txt in (32, 256):
            T = 256 + txt
            g = torch.Generator(device=dev).manual_seed(1 + rank)
            st = [torch.randn(B, T, D, generator=g, device=dev).to(torch.bfloat16) for _ in range(L)]
            te = [(s.float() + 0.1 * torch.randn(B, T, D, generator=g, device=dev)).to(torch.bfloat16) for s in st]
            grads = [torch.empty_like(s) for s in st]
            am = torch.ones(B, txt, dtype=torch.int64, device=dev)

            def one():
                o, s, l = distill_fused(st, te, grads, am, plan, group=None)
                distill_backward(l, grads, s, gout, skip_if_equals=1.0)
                return o

            for _ in range(10):
                o = one()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            iters = 100
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                o = one()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            rec = dict(n_gpus=world, global_batch=global_b, per_gpu_batch=B, txt=txt, ms=ms,
                       units_per_s=global_b * T * L / ms * 1e3, gbs_per_gpu=3 * D * 2 * B * T * L / ms / 1e6,
                       loss=float(o[0]))
            out.append(rec)
            if rank == 0:
                print(json.dumps(rec), flush=True)
            del st, te, grads
            torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"c5_sweep_n{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
