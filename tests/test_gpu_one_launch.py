"""The step as ONE kernel launch (mafed_distill_step / mafed_distill_fwd_step: the last CTA to finish runs the
loss algebra, all CTAs write the modality masks) against the same step as separate launches, bit for bit, and
against the oracle."""
import ctypes

import pytest
import torch

from gpu_util import rel_err
from mafed_b200 import cabi
from mafed_b200.distill_op import (DistillPlan, distill_backward, distill_forward, distill_fused, distill_loss,
                                   modality_masks)
from oracle import distill_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(B, txt, D, L, dtype, seed, ragged=True):
    st, te, am = O.make_inputs(L + 1, B, txt, D, n_vis=256, dtype=dtype, seed=seed)
    if not ragged:
        am = torch.ones_like(am)
    return st, te, am


def _plan(L, loss, modality):
    cfg = O.OracleConfig(modality_strategy=modality, layer_strategy="discounted", gamma=0.5, num_hidden_layers=L,
                         distillation_layer=None, loss=loss)
    layers, coeffs, _ = O.layer_plan(cfg)
    kind = cabi.MODW_EQUAL if modality == "equal" else cabi.MODW_TABLE
    lang = None if modality == "equal" else [0.5] * len(layers)
    plan = DistillPlan(layers=layers, layer_coeffs=[float(c) for c in coeffs], modality_kind=kind, lang_weights=lang,
                       loss_kind=cabi.LOSS_MSE if loss == "mse" else cabi.LOSS_COSINE)
    return cfg, plan


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("loss", ["mse", "cosine"])
@pytest.mark.parametrize("modality", ["equal", "balanced"])
def test_one_launch_step_equals_separate_launches(dtype, loss, modality):
    L = 5
    st, te, am = _inputs(6, 9, 768, L, dtype, seed=7)
    cfg, plan = _plan(L, loss, modality)
    s = [st[l].cuda() for l in plan.layers]
    t = [te[l].cuda() for l in plan.layers]
    mask = am.cuda()
    results = []
    for no_tail in (0, 1):
        with cabi.tuning(TUNE_NO_TAIL=no_tail):
            g = [torch.full_like(x, float("nan")) for x in s]
            both = torch.full((2, 6, 256 + 9), -7, dtype=torch.int64, device="cuda")
            out, scale, ln = distill_fused(s, t, g, mask, plan, group=False, mask_out=(both[0], both[1]))
            torch.cuda.synchronize()
            results.append((out.clone(), [x.clone() for x in g], both.clone()))
    (o1, g1, m1), (o2, g2, m2) = results
    assert torch.equal(o1, o2)                                   # same fixed-order reduction, wherever it runs
    assert all(torch.equal(a, b) for a, b in zip(g1, g2))
    lang, image = modality_masks(mask, 256)
    assert torch.equal(m1[0], lang) and torch.equal(m1[1], image)
    assert torch.equal(m2[0], lang) and torch.equal(m2[1], image)
    ref = O.forward_backward(st, te, am, cfg)
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert float(o1[0]) == pytest.approx(float(ref["loss"]), rel=tol)
    for l, gl in zip(plan.layers, g1):
        assert rel_err(gl.float().cpu(), ref["grads"][l].float()) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_forward_step_equals_forward_plus_epilogue(dtype):
    L = 4
    st, te, am = _inputs(5, 7, 1024, L, dtype, seed=11)
    cfg, plan = _plan(L, "mse", "equal")
    s = [st[l].cuda() for l in plan.layers]
    t = [te[l].cuda() for l in plan.layers]
    mask = am.cuda()
    gout = torch.full((), 0.5, device="cuda")
    results = []
    for no_tail in (0, 1):
        with cabi.tuning(TUNE_NO_TAIL=no_tail):
            out, scale, ln = distill_forward(s, t, mask, plan, group=False)
            g = [torch.empty_like(x) for x in s]
            distill_backward(ln, g, scale, gout)
            torch.cuda.synchronize()
            results.append((out.clone(), g))
    assert torch.equal(results[0][0], results[1][0])
    assert all(torch.equal(a, b) for a, b in zip(results[0][1], results[1][1]))
    ref = O.forward_backward(st, te, am, cfg, grad_out=0.5)
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert float(results[0][0][0]) == pytest.approx(float(ref["loss"]), rel=tol)
    for l, gl in zip(plan.layers, results[0][1]):
        assert rel_err(gl.float().cpu(), ref["grads"][l].float()) < tol


def test_large_mask_step_keeps_the_tail():
    """More than 16 Ki mask entries: the counts come from a prologue launch (left in the ws header), the loss
    algebra and the masks still run inside the fused kernel."""
    L, B, txt, D = 2, 70, 256, 256
    st, te, am = O.make_inputs(L + 1, B, txt, D, n_vis=256, dtype=torch.bfloat16, seed=13)
    cfg, plan = _plan(L, "mse", "equal")
    s = [st[l].cuda() for l in plan.layers]
    t = [te[l].cuda() for l in plan.layers]
    mask = am.cuda()
    results = []
    for no_tail in (0, 1):
        with cabi.tuning(TUNE_NO_TAIL=no_tail):
            g = [torch.empty_like(x) for x in s]
            both = torch.empty((2, B, 256 + txt), dtype=torch.int64, device="cuda")
            out, scale, ln = distill_fused(s, t, g, mask, plan, group=False, mask_out=(both[0], both[1]))
            torch.cuda.synchronize()
            results.append((out.clone(), g, both))
    assert torch.equal(results[0][0], results[1][0])
    assert all(torch.equal(a, b) for a, b in zip(results[0][1], results[1][1]))
    assert torch.equal(results[0][2], results[1][2])
    ref = O.forward_backward(st, te, am, cfg)
    assert float(results[0][0][0]) == pytest.approx(float(ref["loss"]), rel=2e-3)
    for l, gl in zip(plan.layers, results[0][1]):
        assert rel_err(gl.float().cpu(), ref["grads"][l].float()) < 2e-3


def test_tail_counters_are_left_clean():
    """The arrival counters are taken round-robin and reset by the last CTA: many more steps than counters,
    alternating shapes and streams, all give the same bits."""
    L = 3
    st, te, am = _inputs(4, 5, 768, L, torch.bfloat16, seed=3)
    _, plan = _plan(L, "mse", "equal")
    s = [st[l].cuda() for l in plan.layers]
    t = [te[l].cuda() for l in plan.layers]
    mask = am.cuda()
    g = [torch.empty_like(x) for x in s]
    first, _, _ = distill_fused(s, t, g, mask, plan, group=False)
    first = first.clone()
    first_fwd, _, _ = distill_forward(s, t, mask, plan, group=False)
    first_fwd = first_fwd.clone()
    # (the forward pass and the one-pass step use different ring geometries, i.e. group the rows differently into
    #  per-CTA partial sums: equal to rounding, each bit-reproducible on its own)
    torch.testing.assert_close(first_fwd, first, rtol=2e-6, atol=0)
    side = torch.cuda.Stream()
    outs = []
    for i in range(700):
        if i % 3 == 2:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                o, _, _ = distill_forward(s, t, mask, plan, group=False)
            torch.cuda.current_stream().wait_stream(side)
        else:
            o, _, _ = distill_fused(s, t, g, mask, plan, group=False)
        outs.append((i % 3 == 2, o))
    torch.cuda.synchronize()
    for is_fwd, o in outs:
        assert torch.equal(o, first_fwd if is_fwd else first)


def test_step_through_the_c_abi_with_every_output():
    """mafed_distill_step called directly: out, sums, bwd_scale, masks, gradients from one call."""
    lib = cabi.load()
    L, B, txt, D = 3, 4, 6, 512
    st, te, am = _inputs(B, txt, D, L, torch.float32, seed=21)
    cfg, plan = _plan(L, "mse", "equal")
    s = [st[l].cuda() for l in plan.layers]
    t = [te[l].cuda() for l in plan.layers]
    n = len(s)
    mask = am.cuda()
    g = [torch.empty_like(x) for x in s]
    shape = cabi.make_shape(n, B, 256 + txt, 256, D, cabi.F32, cabi.LOSS_MSE)
    ws = torch.empty(lib.mafed_distill_ws_bytes(n), dtype=torch.uint8, device="cuda")
    out = torch.empty(1 + 3 * n, device="cuda")
    scale = torch.empty(2 * n, device="cuda")
    sums = torch.empty(2 * n + 2, dtype=torch.float64, device="cuda")
    both = torch.empty((2, B, 256 + txt), dtype=torch.int64, device="cuda")
    rc = lib.mafed_distill_step(ctypes.byref(shape), cabi.ptr_array([x.data_ptr() for x in s]),
                                cabi.ptr_array([x.data_ptr() for x in t]), cabi.ptr_array([x.data_ptr() for x in g]),
                                mask.data_ptr(), ctypes.byref(plan.weights()), 1.0, ws.data_ptr(), out.data_ptr(),
                                scale.data_ptr(), sums.data_ptr(), both[0].data_ptr(), both[1].data_ptr(), None, None,
                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = O.forward_backward(st, te, am, cfg)
    assert float(out[0]) == pytest.approx(float(ref["loss"]), rel=1e-5)
    assert float(sums[2 * n]) == float(am.sum()) and float(sums[2 * n + 1]) == B * 256
    n_text, n_vis = float(am.sum()), B * 256.0
    w_text = n_text / (n_text + n_vis)
    expect = plan.layer_coeffs[0] * w_text * (2.0 / D) / n_text
    assert float(scale[0]) == pytest.approx(expect, rel=1e-6)
    for l, gl in zip(plan.layers, g):
        assert rel_err(gl.cpu(), ref["grads"][l]) < 1e-5
    # argument errors: masks come as a pair; sharded steps need the sums vector
    assert lib.mafed_distill_step(ctypes.byref(shape), None, None, None, None, ctypes.byref(plan.weights()), 1.0,
                                  ws.data_ptr(), out.data_ptr(), scale.data_ptr(), None, both[0].data_ptr(), None,
                                  None, None, None) == -1


def test_autograd_op_fills_the_batch_masks():
    st, te, am = _inputs(3, 5, 768, 2, torch.bfloat16, seed=5)
    _, plan = _plan(2, "mse", "equal")
    s = [st[l].cuda().requires_grad_(True) for l in plan.layers]
    t = [te[l].cuda() for l in plan.layers]
    mask = am.cuda()
    both = torch.zeros((2, 3, 261), dtype=torch.int64, device="cuda")
    total, aux = distill_loss(s, t, mask, plan, group=False, mask_out=both)
    total.backward()
    lang, image = modality_masks(mask, 256)
    assert torch.equal(both[0], lang) and torch.equal(both[1], image)
    assert all(x.grad is not None for x in s)


def test_missing_peer_poisons_the_step_instead_of_hanging():
    """A rank whose peer never arrives: the in-kernel waits run into the spin bound, the status is set and the
    loss comes out NaN (loop-back communicator of world 2 on one GPU, 50 ms bound)."""
    import time
    lib = cabi.load()
    L, B, txt, D = 2, 3, 5, 512
    st, te, am = _inputs(B, txt, D, L, torch.float32, seed=23)
    _, plan = _plan(L, "mse", "equal")
    s = [st[l].cuda() for l in plan.layers]
    t = [te[l].cuda() for l in plan.layers]
    n = len(s)
    mask = am.cuda()
    g = [torch.empty_like(x) for x in s]
    handle = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(lib.mafed_comm_handle_bytes())
    assert lib.mafed_comm_create(2, 0, buf, ctypes.byref(handle)) == 0
    try:
        assert lib.mafed_comm_connect(handle, None) == 0          # loop-back: rank 1 never writes its slots
        assert lib.mafed_comm_set_timeout(handle, 0.05) == 0
        shape = cabi.make_shape(n, B, 256 + txt, 256, D, cabi.F32, cabi.LOSS_MSE)
        ws = torch.empty(lib.mafed_distill_ws_bytes(n), dtype=torch.uint8, device="cuda")
        out = torch.zeros(1 + 3 * n, device="cuda")
        scale = torch.empty(2 * n, device="cuda")
        sums = torch.empty(2 * n + 2, dtype=torch.float64, device="cuda")
        t0 = time.perf_counter()
        rc = lib.mafed_distill_step(ctypes.byref(shape), cabi.ptr_array([x.data_ptr() for x in s]),
                                    cabi.ptr_array([x.data_ptr() for x in t]), cabi.ptr_array([x.data_ptr() for x in g]),
                                    mask.data_ptr(), ctypes.byref(plan.weights()), 1.0, ws.data_ptr(), out.data_ptr(),
                                    scale.data_ptr(), sums.data_ptr(), None, None, handle, None,
                                    torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        assert time.perf_counter() - t0 < 5.0                      # bounded: two waits of 50 ms, not a hang
        status = ctypes.c_int(0)
        assert lib.mafed_comm_status(handle, ctypes.byref(status)) == 0 and status.value == 1
        assert torch.isnan(out[0]) and all(torch.isnan(x).any() for x in g)
        # the strategy notices by itself: the status word is mapped host memory, read (no sync) by every distill()
        from gpu_util import make_method
        from mafed_b200 import comm as C
        key = (0, torch.cuda.current_device())
        saved = C._cache.get(key, "absent")
        C._cache[key] = C.PeerComm(handle, 2, 0)
        try:
            fd = make_method(dict(modality="equal", layer_strategy="discounted", loss="mse", gamma=0.5,
                                  num_hidden_layers=L, layer=None, n_vis=256, coeff=1.0, cls=False, lang_coeff=None))
            with pytest.raises(cabi.MafedDistillError, match="did not reach a distillation exchange"):
                fd.check_exchange(sync=False)
        finally:
            if saved == "absent":
                C._cache.pop(key, None)
            else:
                C._cache[key] = saved
    finally:
        lib.mafed_comm_destroy(handle)


@pytest.mark.parametrize("B,txt", [(5, 9), (70, 256)], ids=["small-mask", "mask-over-16Ki"])
def test_counts_prefetched_on_one_rank(B, txt):
    """`prefetch_counts(force=True)` on a single rank: the step reads the token counts from the ticket instead of
    summing the mask itself -- which also keeps a step with a large mask (> 16 Ki entries) a single launch.  Same
    bits as the step without a ticket; a ticket whose mask was edited afterwards is ignored."""
    from gpu_util import Out, make_method
    from golden_util import oracle_cfg
    L, D = 2, 256
    st, te, am = O.make_inputs(L + 1, B, txt, D, n_vis=256, dtype=torch.bfloat16, seed=17)
    meta = dict(modality="equal", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=L, layer=None,
                n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
    ref = O.forward_backward(st, te, am, oracle_cfg(meta))
    te_c = [t.cuda() for t in te]

    def run(prefetch, touch=False):
        fd = make_method(meta)
        fd.past_model = lambda **kw: Out(tuple(te_c))
        leaves = [s.cuda().requires_grad_(True) for s in st]
        mask = am.cuda()
        batch = {"attention_mask": mask}
        if prefetch:
            ticket = fd.prefetch_counts(batch, force=True)
            assert ticket is not None and len(fd._tickets) == 1
            if touch:
                mask.mul_(1)                                   # bumps the tensor's version: the ticket is stale
        loss = fd.distill(Out(tuple(leaves)), batch)
        assert not fd._tickets
        loss.backward()
        torch.cuda.synchronize()
        if prefetch and not touch:
            tk = ticket.tensor.cpu()
            assert int(tk[0]) == 0                             # no exchange epoch on a single rank
            assert tk[1:3].view(torch.float64).tolist() == [float(am.sum()), float(B * 256)]
        return loss.detach(), [x.grad for x in leaves]

    base_loss, base_grads = run(False)
    for kwargs in (dict(prefetch=True), dict(prefetch=True, touch=True)):
        loss, grads = run(**kwargs)
        assert torch.equal(loss, base_loss)
        assert all(torch.equal(a, b) for a, b in zip(grads[:L], base_grads[:L]))
    assert float(base_loss) == pytest.approx(float(ref["loss"]), rel=2e-3)
    for l in range(L):
        assert rel_err(base_grads[l].float().cpu(), ref["grads"][l].float()) < 2e-3
