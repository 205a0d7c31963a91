"""world_size-2 gloo test (CPU) of the host side of the batch-sharded path: the product's collective
wrapper combines per-rank partial sums + counts; finalising them reproduces the full-batch oracle."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partials(st, te, am, n_vis, layers):
    """Per-rank [2L+2] fp64 vector: per layer (text sum, vision sum) of w * ||h-p||^2, then counts --
    exactly what mafed_distill_reduce leaves on the device."""
    B, txt = am.shape
    w_text = torch.cat([torch.zeros(B, n_vis), am.double()], 1)
    w_vis = torch.cat([torch.ones(B, n_vis), torch.zeros(B, txt)], 1).double()
    vec = []
    for l in layers:
        d2 = (st[l].double() - te[l].double()).pow(2).sum(-1)
        vec += [(d2 * w_text).sum(), (d2 * w_vis).sum()]
    vec += [w_text.sum(), w_vis.sum()]
    return torch.stack(vec)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mafed_b200.distill_op import allreduce_sums, resolve_group
    from oracle import distill_oracle as O
    st, te, am = O.make_inputs(4, 6, 5, 32, n_vis=8, seed=41)
    lo, hi = (0, 2) if rank == 0 else (2, 6)          # uneven shards
    local = _partials([s[lo:hi] for s in st], [t[lo:hi] for t in te], am[lo:hi], 8, [0, 1, 2])
    assert resolve_group(None) == (True, None) and resolve_group(False) == (False, None)
    total = allreduce_sums(local.clone())
    full = _partials(st, te, am, 8, [0, 1, 2])
    torch.testing.assert_close(total, full, rtol=1e-12, atol=0)
    # finalise like the epilogue (balanced, equal layer weights) and compare with the oracle
    n_t, n_v, D = float(total[-2]), float(total[-1]), 32
    loss = sum((0.5 * float(total[2 * i]) / (D * n_t) + 0.5 * float(total[2 * i + 1]) / (D * n_v)) / 3 for i in range(3))
    cfg = O.OracleConfig(modality_strategy="balanced", layer_strategy="equal", num_hidden_layers=3,
                         distillation_layer=None, num_vision_tokens=8)
    ref = O.forward_backward(st, te, am, cfg)
    assert abs(loss - float(ref["loss"])) / float(ref["loss"]) < 1e-5
    # with a process group the strategy returns the GLOBAL-batch loss, and DDP will average the
    # parameter gradients over the ranks -- so by default the step multiplies its gradients by world_size; an explicit
    # grad_multiplier (gradients summed, not averaged) or process_group=False (per-rank loss) turns that off
    from mafed_b200.methods import CLMethod

    class Opts:
        tasks = ["a", "b", "c"]; batch_size = 4; seed = 42; pin_mem = False; accumulate_grad_batches = 4

    def make(**kw):
        return CLMethod["featdistill"](memory_size=8, opts=Opts(), model_type="vlpythia",
                                       distillation_modality_weighing_strategy="balanced",
                                       distillation_layer_weighing_strategy="equal", distillation_layer=None,
                                       num_hidden_layers=3, **kw)

    def multiplier(fd):
        coeffs, kind, lang = fd._tables([0, 1, 2])
        return fd._plan([0, 1, 2], coeffs, 1.0, kind, lang).grad_multiplier

    assert multiplier(make()) == float(world)
    assert multiplier(make(grad_multiplier=1.0)) == 1.0
    assert multiplier(make(process_group=False)) == 1.0
    fd = make()
    assert fd.assumed_grad_out == 0.25                      # 1 / accumulate_grad_batches
    plan = fd._step_plan([0, 1, 2])
    assert plan.grad_multiplier == float(world) and plan.assumed_grad_out == 0.25
    ret[rank] = True
    dist.destroy_process_group()


def test_two_rank_gloo_partial_sum_combine():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_resolve_group_without_process_group():
    sys.path.insert(0, ROOT)
    from mafed_b200.distill_op import resolve_group
    assert resolve_group(None) == (False, None) and resolve_group(False) == (False, None)
