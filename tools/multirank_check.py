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
        fd.grad_multiplier = 1.0      # hidden-state gradients of the GLOBAL loss (the DDP convention is checked below)
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
        ok &= prefetch_check(rank, world)
        ok &= peer.status() == 0
        ok &= host_step_check(rank, world)
        ok &= peer.status() == 0
    ok &= ddp_check(rank, world)
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


def prefetch_check(rank, world):
    """Token counts sent ahead of the step (`fd.prefetch_counts`): one and two batches ahead, uneven ragged shards,
    a late rank -- every step equals the full-batch oracle; a stale ticket (mask edited after the prefetch) is not
    used."""
    import time

    from golden_util import oracle_cfg
    from gpu_util import Out, make_method, rel_err
    from oracle import distill_oracle as O
    meta = dict(modality="equal", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=3,
                layer=None, n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
    cfg = oracle_cfg(meta)
    B = 3 * world
    lo, hi = 3 * rank, 3 * rank + 3
    fd = make_method(meta)
    fd.grad_multiplier = 1.0
    batches = []
    for seed in (51, 52, 53, 54, 55):
        st, te, am = O.make_inputs(4, B, 9, 768, n_vis=256, dtype=torch.float32, seed=seed, mask="ragged")
        batches.append((st, te, am, O.forward_backward(st, te, am, cfg), am[lo:hi].cuda().contiguous()))
    ok = True

    def run(i):
        st, te, am, ref, mask = batches[i]
        leaves = [x[lo:hi].cuda().contiguous().requires_grad_(True) for x in st]
        teach = [x[lo:hi].cuda().contiguous() for x in te]
        fd.past_model = lambda **kw: Out(tuple(teach))
        loss = fd.distill(Out(tuple(leaves)), {"attention_mask": mask})
        loss.backward()
        torch.cuda.synchronize()
        e_loss = abs(float(loss) - float(ref["loss"])) / abs(float(ref["loss"]))
        e_grad = max(rel_err(leaves[l].grad.cpu(), ref["grads"][l][lo:hi]) for l in range(3))
        return e_loss < 1e-5 and e_grad < 1e-5

    # one ahead, two ahead, and a step whose ticket is stale
    assert fd.prefetch_counts({"attention_mask": batches[0][4]}) is not None
    fd.prefetch_counts({"attention_mask": batches[1][4]})
    ok &= run(0)
    fd.prefetch_counts({"attention_mask": batches[2][4]})
    ok &= run(1)
    if rank == world - 1:
        time.sleep(2.0)                 # a late rank: its counts were sent long ago, the sums exchange waits for it
    ok &= run(2)
    fd.prefetch_counts({"attention_mask": batches[3][4]})
    batches[3][4].add_(0)               # in-place touch: the ticket no longer matches the tensor's version
    ok &= run(3)                        # -> counts exchanged inside the kernel instead
    ok &= run(4)                        # no prefetch at all
    print(f"rank {rank}: prefetched token counts {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def host_step_check(rank, world):
    """The C ABI's host-buffer step with a communicator (`mafed_host_step_run(..., comm)`) against the oracle."""
    from golden_util import oracle_cfg
    from gpu_util import make_method, rel_err
    from mafed_b200.host_step import CHostStep
    from oracle import distill_oracle as O
    meta = dict(modality="balanced", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=3,
                layer=None, n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
    B = 2 * world
    lo, hi = 2 * rank, 2 * rank + 2
    st, te, am = O.make_inputs(4, B, 9, 1024, n_vis=256, dtype=torch.bfloat16, seed=61, mask="ragged")
    ref = O.forward_backward(st, te, am, oracle_cfg(meta))
    fd = make_method(meta)
    hs = CHostStep(fd, [x[lo:hi].contiguous() for x in st[:3]], [x[lo:hi].contiguous() for x in te[:3]],
                   am[lo:hi].contiguous(), torch.device("cuda", torch.cuda.current_device()))
    ok = hs.comm is not None
    for _ in range(2):
        loss = hs.step()
        e_loss = abs(float(loss) - float(ref["loss"])) / abs(float(ref["loss"]))
        e_grad = max(rel_err(hs.h_g[l].float(), ref["grads"][l][lo:hi].float()) for l in range(3))
        ok &= e_loss < 2e-3 and e_grad < 2e-3
    hs.close()
    print(f"rank {rank}: sharded host-buffer step {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def ddp_check(rank, world):
    """under DistributedDataParallel the distillation term must weigh what it weighs on one GPU.
    A tiny model whose forward returns a hidden-state tuple, wrapped in DDP, each rank on its shard with the
    strategy's defaults (batch-sharded global loss, gradients multiplied by world_size because DDP averages): the
    averaged parameter gradients equal those of the same model on the full batch on one device."""
    from gpu_util import Out, make_method, rel_err
    from torch.nn.parallel import DistributedDataParallel as DDP

    class Tiny(torch.nn.Module):
        def __init__(self, d_in=24, d=64):
            super().__init__()
            g = torch.Generator().manual_seed(5)
            self.w0 = torch.nn.Parameter(torch.randn(d_in, d, generator=g) * 0.2)
            self.w1 = torch.nn.Parameter(torch.randn(d, d, generator=g) * 0.2)
            self.w2 = torch.nn.Parameter(torch.randn(d, d, generator=g) * 0.2)

        def forward(self, x):
            h0 = x @ self.w0
            h1 = torch.tanh(h0) @ self.w1
            h2 = torch.tanh(h1) @ self.w2
            return h0, h1, h2

    meta = dict(modality="equal", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=3,
                layer=None, n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
    B, txt = 3 * world, 6
    lo, hi = 3 * rank, 3 * rank + 3
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, 256 + txt, 24, generator=g)
    am = torch.zeros(B, txt, dtype=torch.int64)
    for b in range(B):
        am[b, txt - (1 + (5 * b) % txt):] = 1
    student, teacher = Tiny().cuda(), Tiny().cuda()
    with torch.no_grad():
        for p in teacher.parameters():
            p.add_(0.05)
        t_full = [h.detach() for h in teacher(x.cuda())]

    def grads_of(model, xs, teach, mask, group):
        fd = make_method(meta)
        fd.process_group = group
        fd.past_model = lambda **kw: Out(tuple(teach))
        model.zero_grad(set_to_none=True)
        loss = fd.distill(Out(model(xs)), {"attention_mask": mask})
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach()

    ddp = DDP(student, device_ids=[torch.cuda.current_device()])
    loss_d = grads_of(ddp, x[lo:hi].cuda(), [h[lo:hi].contiguous() for h in t_full], am[lo:hi].cuda(), None)
    got = [p.grad.clone() for p in student.parameters()]
    single = Tiny().cuda()                                                   # same seeded weights, never wrapped
    loss_f = grads_of(single, x.cuda(), t_full, am.cuda(), False)            # the full batch on this GPU alone
    want = [p.grad.clone() for p in single.parameters()]
    e_loss = abs(float(loss_d) - float(loss_f)) / abs(float(loss_f))
    e_grad = max(rel_err(a.cpu(), b.cpu()) for a, b in zip(got, want))
    ok = e_loss < 1e-5 and e_grad < 1e-5
    print(f"rank {rank}: DDP parameter gradients vs full batch on one GPU: loss err {e_loss:.2e} grad err {e_grad:.2e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    return ok


if __name__ == "__main__":
    main()
