#!/usr/bin/env python
"""Batch-sharded parity check, one process per GPU (run under torchrun, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multirank_check.py

Every rank holds a contiguous slice of the batch; the sharded loss and the concatenation of the
per-rank gradients must equal the single-device CPU oracle on the full batch.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gpu_util import Out, make_method, rel_err
    from golden_util import oracle_cfg
    from oracle import distill_oracle as O

    ok = True
    for dtype, loss, modality, single_pass in [(torch.float32, "mse", "equal", True), (torch.bfloat16, "mse", "balanced", True),
                                               (torch.bfloat16, "cosine", "equal", False), (torch.float32, "mse", "equal", False)]:
        B = 4 * world + 2  # uneven shards: the last rank gets the remainder
        st, te, am = O.make_inputs(4, B, 9, 768, n_vis=256, dtype=dtype, seed=31, mask="ragged")
        meta = dict(modality=modality, layer_strategy="discounted", loss=loss, gamma=0.5, num_hidden_layers=3,
                    layer=None, n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
        ref = O.forward_backward(st, te, am, oracle_cfg(meta), grad_out=0.5 if not single_pass else 1.0)
        lo = rank * 4
        hi = B if rank == world - 1 else lo + 4
        fd = make_method(meta, single_pass=single_pass)
        leaves = [s[lo:hi].cuda().contiguous().requires_grad_(True) for s in st]
        teach = [t[lo:hi].cuda().contiguous() for t in te]
        fd.past_model = lambda **kw: Out(tuple(teach))
        loss_v = fd.distill(Out(tuple(leaves)), {"attention_mask": am[lo:hi].cuda()})
        (loss_v * (0.5 if not single_pass else 1.0)).backward()
        torch.cuda.synchronize()
        tol = 1e-5 if dtype == torch.float32 else 2e-3
        e_loss = abs(float(loss_v) - float(ref["loss"])) / abs(float(ref["loss"]))
        e_grad = max(rel_err(leaves[l].grad.float().cpu(), ref["grads"][l][lo:hi].float()) for l in range(3))
        good = e_loss < tol and e_grad < tol
        ok &= good
        print(f"rank {rank}/{world} {dtype} {loss} {modality} single_pass={single_pass}: loss err {e_loss:.2e} "
              f"grad err {e_grad:.2e} {'OK' if good else 'FAIL'}", flush=True)
    from mafed_b200.comm import get_peer_comm
    peer = get_peer_comm(None)
    mode = "nccl" if peer is None else "peer-memory"
    if peer is not None:
        ok &= peer.status() == 0
    if peer is not None:
        ok &= graph_replay_check(rank, world)
        ok &= peer.status() == 0
    if os.environ.get("MAFED_B200_DIST", "peer") == "peer":
        ok &= peer is not None          # on one NVLink box the mailboxes must map
    print(f"rank {rank}: exchange path = {mode}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


def graph_replay_check(rank, world):
    """The sharded one-pass step captured into a CUDA graph (both exchanges inside the fused kernel, epochs on the
    device), replayed with new data, then followed by eager steps: every result equals the full-batch oracle."""
    from golden_util import oracle_cfg
    from gpu_util import rel_err
    from mafed_b200 import cabi
    from mafed_b200.distill_op import DistillPlan, distill_backward, distill_fused
    from oracle import distill_oracle as O
    meta = dict(modality="equal", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=3,
                layer=None, n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
    cfg = oracle_cfg(meta)
    layers, coeffs, _ = O.layer_plan(cfg)
    plan = DistillPlan(layers=layers, layer_coeffs=[float(c) for c in coeffs], modality_kind=cabi.MODW_EQUAL)
    B = 3 * world
    lo, hi = 3 * rank, 3 * rank + 3

    def data(seed):
        st, te, am = O.make_inputs(4, B, 9, 768, n_vis=256, dtype=torch.bfloat16, seed=seed, mask="ragged")
        return st, te, am, O.forward_backward(st, te, am, cfg)

    st, te, am, ref = data(41)
    s = [x[lo:hi].cuda().contiguous() for x in st[:3]]
    t = [x[lo:hi].cuda().contiguous() for x in te[:3]]
    g = [torch.empty_like(x) for x in s]
    mask = am[lo:hi].cuda().contiguous()
    gout = torch.ones((), device="cuda")

    def step():
        out, scale, ln = distill_fused(s, t, g, mask, plan, group=None)
        distill_backward(ln, g, scale, gout, skip_if_equals=1.0)
        return out

    def good(out, ref):
        torch.cuda.synchronize()
        e_loss = abs(float(out[0]) - float(ref["loss"])) / abs(float(ref["loss"]))
        e_grad = max(rel_err(g[l].float().cpu(), ref["grads"][l][lo:hi].float()) for l in range(3))
        return e_loss < 2e-3 and e_grad < 2e-3

    ok = True
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            out = step()
    torch.cuda.current_stream().wait_stream(side)
    ok &= good(out, ref)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for seed in (42, 43, 44):
        st, te, am, ref = data(seed)
        for dst, src in zip(s + t, [x[lo:hi] for x in st[:3]] + [x[lo:hi] for x in te[:3]]):
            dst.copy_(src)
        mask.copy_(am[lo:hi])
        graph.replay()
        ok &= good(out, ref)
    eager = step()                      # eager steps after replays: the device-side epochs simply go on
    ok &= good(eager, ref)
    # a rank that arrives late (data loader hiccup, checkpoint write): the peers' kernels wait for it inside the
    # spin bound (60 s by default) and the step is still exact
    import time
    torch.cuda.synchronize()
    if rank == world - 1:
        time.sleep(3.0)
    late = step()
    ok &= good(late, ref)
    print(f"rank {rank}: CUDA-graph replay of the sharded step {'OK' if ok else 'FAIL'}", flush=True)
    return ok


if __name__ == "__main__":
    main()
