"""The step inside a CUDA graph (``mafed_b200.graphed.GraphedDistillStep``): same bits as the eager step, new data
per replay, and the backward gate starting the exact backward from the device inside a replay."""
import pytest
import torch

from gpu_util import Out, make_method, rel_err
from golden_util import oracle_cfg
from oracle import distill_oracle as O

pytestmark = pytest.mark.gpu

META = dict(modality="balanced", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=3, layer=None,
            n_vis=256, coeff=1.0, cls=False, lang_coeff=None)


def _eager(meta, st, te, am, grad_out=1.0):
    fd = make_method(meta)
    te_c = [t.cuda() for t in te]
    fd.past_model = lambda **kw: Out(tuple(te_c))
    leaves = [s.cuda().requires_grad_(True) for s in st]
    loss = fd.distill(Out(tuple(leaves)), {"attention_mask": am.cuda()})
    (loss * grad_out).backward()
    torch.cuda.synchronize()
    return loss.detach().clone(), [l.grad for l in leaves]


@pytest.mark.parametrize("loss_kind", ["mse", "cosine"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_graphed_step_equals_eager_step_and_takes_new_data(loss_kind, dtype):
    from mafed_b200.graphed import GraphedDistillStep
    meta = dict(META, loss=loss_kind)
    st, te, am = O.make_inputs(4, 3, 6, 768, n_vis=256, dtype=dtype, seed=31)
    fd = make_method(meta)
    step = fd.capture([s.cuda() for s in st], [t.cuda() for t in te], am.cuda().clone())
    assert isinstance(step, GraphedDistillStep) and step.layers == [0, 1, 2]
    for trial, seed in enumerate((31, 32, 33)):
        st2, te2, am2 = O.make_inputs(4, 3, 6, 768, n_vis=256, dtype=dtype, seed=seed,
                                      teacher="close" if trial != 1 else "independent")
        with torch.no_grad():
            for dst, src in zip(step.students, st2):
                dst.copy_(src)
            for dst, src in zip(step.teachers, te2):
                dst.copy_(src)
            step.attention_mask.copy_(am2)
        loss = step.replay()
        torch.cuda.synchronize()
        e_loss, e_grads = _eager(meta, st2, te2, am2)
        assert torch.equal(loss, e_loss)                       # same kernels, same fixed-order reductions
        for l, g in zip(step.layers, step.grads):
            assert torch.equal(g, e_grads[l])
        assert step.students[3].grad is None                   # not a distilled layer
        ref = O.forward_backward(st2, te2, am2, oracle_cfg(meta))
        tol = 1e-5 if dtype == torch.float32 else 2e-3
        assert float(loss) == pytest.approx(float(ref["loss"]), rel=tol)
        for l, g in zip(step.layers, step.grads):
            assert rel_err(g.float().cpu(), ref["grads"][l].float()) < tol
        assert torch.allclose(step.layer_losses[:3].cpu(), torch.stack([ref["layer_losses"][l] for l in range(3)]).float(),
                              rtol=tol)


def test_gate_starts_the_backward_from_the_device_inside_a_replay():
    """Upstream gradient 0.25 while the captured step assumed 1: every replay's gate launches the exact backward
    itself (device-side tail launch inside a graph node)."""
    from mafed_b200.graphed import GraphedDistillStep
    st, te, am = O.make_inputs(4, 3, 6, 1024, n_vis=256, dtype=torch.bfloat16, seed=41)
    fd = make_method(META)
    fd.adapt_assumed_grad_out = False
    assert fd.assumed_grad_out == 1.0
    step = GraphedDistillStep(fd, [s.cuda() for s in st], [t.cuda() for t in te], am.cuda().clone(), grad_out=0.25)
    for _ in range(20):
        step.replay()
    torch.cuda.synchronize()
    ref = O.forward_backward(st, te, am, oracle_cfg(META), grad_out=0.25)
    assert float(step.loss) == pytest.approx(float(ref["loss"]), rel=2e-3)
    for l, g in zip(step.layers, step.grads):
        assert rel_err(g.float().cpu(), ref["grads"][l].float()) < 2e-3


def test_assumed_upstream_gradient_follows_what_the_gate_sees():
    """A trainer that scales the loss by a constant the strategy did not expect: the first steps pay the exact
    backward (started by the gate), then `assumed_grad_out` is re-aimed and the one-pass gradients are final."""
    st, te, am = O.make_inputs(4, 3, 6, 512, n_vis=256, dtype=torch.float32, seed=43)
    fd = make_method(META)
    te_c = [t.cuda() for t in te]
    fd.past_model = lambda **kw: Out(tuple(te_c))
    ref = O.forward_backward(st, te, am, oracle_cfg(META), grad_out=0.125)
    for i in range(4):
        leaves = [s.cuda().requires_grad_(True) for s in st]
        loss = fd.distill(Out(tuple(leaves)), {"attention_mask": am.cuda()})
        (loss * 0.125).backward()
        torch.cuda.synchronize()
        for l in range(3):
            assert rel_err(leaves[l].grad.cpu(), ref["grads"][l]) < 1e-5      # exact on every step
    assert fd.assumed_grad_out == pytest.approx(0.125) and fd.single_pass
    # an upstream gradient that never settles (dynamic loss scaling): the strategy falls back to the two-pass form
    for i in range(12):
        leaves = [s.cuda().requires_grad_(True) for s in st]
        loss = fd.distill(Out(tuple(leaves)), {"attention_mask": am.cuda()})
        (loss * float(2 ** i)).backward()
        torch.cuda.synchronize()
    assert not fd.single_pass
    leaves = [s.cuda().requires_grad_(True) for s in st]
    (fd.distill(Out(tuple(leaves)), {"attention_mask": am.cuda()}) * 0.125).backward()
    for l in range(3):
        assert rel_err(leaves[l].grad.cpu(), ref["grads"][l]) < 1e-5


def test_steps_back_to_back_behind_a_gate_that_starts_the_backward():
    """Kernel-level steps enqueued back to back (no host gap), every gate finding a mismatch and starting the backward
    from the device: the next step's fused kernel must not become resident ahead of that backward (the library
    launches the kernel that follows a gate without the programmatic-dependent-launch attribute).  Eager and as two
    steps inside ONE captured graph."""
    from mafed_b200 import cabi
    from mafed_b200.distill_op import DistillPlan, distill_backward, distill_fused
    cfg = oracle_cfg(META)
    layers, coeffs, _ = O.layer_plan(cfg)
    plan = DistillPlan(layers=layers, layer_coeffs=[float(c) for c in coeffs], modality_kind=cabi.MODW_TABLE,
                       lang_weights=[0.5] * len(layers))
    st, te, am = O.make_inputs(4, 6, 9, 2048, n_vis=256, dtype=torch.bfloat16, seed=47)
    s = [x.cuda() for x in st[:3]]
    t = [x.cuda() for x in te[:3]]
    mask = am.cuda()
    gout = torch.full((), 0.5, device="cuda")
    ref = O.forward_backward(st, te, am, cfg, grad_out=0.5)

    def step(g):
        out, scale, ln = distill_fused(s, t, g, mask, plan, group=False)       # assumes an upstream gradient of 1
        distill_backward(ln, g, scale, gout, skip_if_equals=1.0)               # ... finds 0.5: exact backward
        return out

    def check(g):
        torch.cuda.synchronize()
        for l in range(3):
            assert rel_err(g[l].float().cpu(), ref["grads"][l].float()) < 2e-3

    g1 = [torch.empty_like(x) for x in s]
    g2 = [torch.empty_like(x) for x in s]
    for i in range(200):
        step(g1 if i % 2 == 0 else g2)
    check(g1)
    check(g2)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(g1)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    for g in (g1, g2):
        for x in g:
            x.zero_()
    with torch.cuda.graph(graph):
        step(g1)
        step(g2)
    for _ in range(50):
        graph.replay()
    check(g1)
    check(g2)
