"""The fused distillation operator: host tables -> C ABI -> sm_100a kernels, with autograd.

Two ways to run a step, same results:

* two-pass  -- fused forward over all selected layers (the loss algebra runs in the kernel's last CTA),
  later one fused backward: 5*D*e bytes of HBM traffic per token*layer, two launches;
* one-pass  -- the gradient scale depends only on the token counts and the host weight tables, not
  on the loss, so one kernel produces the loss sums AND the gradients from a single read of student
  and teacher (3*D*e bytes).  The upstream gradient is assumed (``plan.assumed_grad_out``, i.e.
  ``1 / accumulate_grad_batches`` under Lightning); ``backward`` launches a fix-up that returns at
  once when the real upstream gradient equals the assumed one and otherwise recomputes exactly.
  The whole step -- modality masks, scale table, loss sums, gradients, losses, and across batch
  shards the NVLink exchange of counts and sums -- is ONE kernel launch (``mafed_distill_step``).

Under ``torch.distributed`` with batch sharding, the per-rank ``[2L+2]`` fp64 partial sums / counts
are combined by allreduce; the backward needs no collective.  Nothing here
synchronises the host with the device.

Host side: ``distill_loss`` hands the whole step to the compiled autograd node (``csrc/torch_node.cpp``,
``mafed_b200.node``) -- pointer tables, workspace and gradient allocation, stream lookup, the launch and the
backward gate all run in C++.  The ctypes functions below (``distill_forward`` / ``distill_fused`` /
``distill_backward``) are the same C-ABI calls one at a time: the NCCL sequence, the kernel-level benchmarks and
the tests use them.

Replaces the per-layer Python loop of ``mafed/methods/distillation.py:105-166`` and the autograd
chains behind ``:226-249``.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import cabi, node
from .comm import get_peer_comm

_DTYPES = {torch.float32: cabi.F32, torch.bfloat16: cabi.BF16, torch.float16: cabi.F16}


@dataclass
class DistillPlan:
    """Host-side description of one distillation step (which layers, which weights)."""
    layers: List[int]                       # indices into the hidden-state tuple
    layer_coeffs: List[float]               # get_layer_loss_weight(layer) per selected layer
    distill_coeff: float = 1.0
    modality_kind: int = cabi.MODW_EQUAL
    lang_weights: Optional[List[float]] = None   # MODW_TABLE: language weight per selected layer
    loss_kind: int = cabi.LOSS_MSE
    cls: bool = False
    n_vis: int = 256
    grad_multiplier: float = 1.0            # e.g. world_size to undo DDP's gradient averaging
    single_pass: bool = True                # one-pass step when gradients are needed
    assumed_grad_out: float = 1.0           # upstream gradient the one-pass step bakes in
    _weights: Optional[cabi.Weights] = field(default=None, repr=False)
    _node_plan: object = field(default=None, repr=False)

    def weights(self) -> cabi.Weights:
        if self._weights is None:
            if len(self.layers) > cabi.MAX_LAYERS:
                raise ValueError(f"at most {cabi.MAX_LAYERS} layers per call")
            self._weights = cabi.make_weights(
                cabi.MODW_CLS if self.cls else self.modality_kind, self.distill_coeff,
                self.layer_coeffs, self.lang_weights)
        return self._weights

    def node_plan(self):
        """The same tables as the compiled node's ``Plan`` (built once per DistillPlan)."""
        if self._node_plan is None:
            self._node_plan = node.load().Plan(
                cabi.MODW_CLS if self.cls else self.modality_kind, float(self.distill_coeff),
                [float(c) for c in self.layer_coeffs],
                None if self.lang_weights is None else [float(c) for c in self.lang_weights],
                self.loss_kind, bool(self.cls), int(self.n_vis), float(self.grad_multiplier), bool(self.single_pass),
                float(self.assumed_grad_out))
        return self._node_plan


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream on `device` (fast path avoids building a Stream object)."""
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` that costs nothing when `dev` is already current (the usual case)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        same = _raw_device is not None and device.index is not None and _raw_device() == device.index
        self.ctx = None if same else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise cabi.MafedDistillError(
            f"{what} is on {t.device}: the distillation path runs only as sm_100a CUDA kernels "
            "(there is no CPU fallback)")


def _prepare(tensors: Sequence[torch.Tensor]):
    return [t if t.is_contiguous() else t.contiguous() for t in tensors]


class _DevicePtr:
    """A raw device address inside a torch allocation that it keeps alive (``.data_ptr()`` like a tensor)."""
    __slots__ = ("owner", "ptr")

    def __init__(self, owner, ptr):
        self.owner, self.ptr = owner, ptr

    def data_ptr(self):
        return self.ptr


class _Launch:
    """Geometry + pointer tables of one step; keeps the tensors alive until the kernels are queued."""

    def __init__(self, students, teachers, attn_mask, plan: DistillPlan):
        s0 = students[0]
        _require_cuda(s0, "hidden_states")
        if s0.dtype not in _DTYPES:
            raise TypeError(f"unsupported hidden-state dtype {s0.dtype} (float32, bfloat16, float16)")
        if s0.dim() != 3:
            raise ValueError("hidden states must be [B, T, D]")
        shape, dtype, device = s0.shape, s0.dtype, s0.device
        self.B, self.T, self.D = shape
        for group in (students, teachers):
            for t in group:
                if t.shape != shape:
                    raise ValueError("all selected hidden states must share one [B, T, D] shape")
                if t.dtype != dtype or t.device != device:
                    raise ValueError("student / teacher dtype or device mismatch")
        self.device = device
        self.dtype = dtype
        n_vis = plan.n_vis
        if plan.cls:
            self.mask = None
        else:
            if attn_mask is None:
                raise ValueError("attention_mask is required")
            _require_cuda(attn_mask, "attention_mask")
            if attn_mask.dtype != torch.int64:
                attn_mask = attn_mask.to(torch.int64)
            if attn_mask.shape != (self.B, self.T - n_vis):
                raise ValueError(
                    f"attention_mask shape {tuple(attn_mask.shape)} != (B, T - n_vis) = {(self.B, self.T - n_vis)}")
            self.mask = attn_mask.contiguous()
        self.n_layers = len(students)
        self.shape = cabi.make_shape(self.n_layers, self.B, self.T, n_vis if not plan.cls else min(n_vis, self.T),
                                     self.D, _DTYPES[self.dtype], plan.loss_kind, plan.cls)
        self.shape_ref = ctypes.byref(self.shape)
        self.students = students
        self.teachers = teachers
        self.s_ptrs = cabi.ptr_array([t.data_ptr() for t in students])
        self.t_ptrs = cabi.ptr_array([t.data_ptr() for t in teachers])
        self.mask_ptr = self.mask.data_ptr() if self.mask is not None else None

    def alloc_scalars(self, lib):
        """(workspace, out, bwd_scale): the partial-sum workspace and the scale table share one scratch
        allocation; `bwd_scale` is returned as a `_DevicePtr` view that keeps the scratch alive."""
        L, dev = self.n_layers, self.device
        ws_bytes = (lib.mafed_distill_ws_bytes(L) + 255) & ~255
        scratch = torch.empty(ws_bytes + 8 * L, dtype=torch.uint8, device=dev)
        out = torch.empty(1 + 3 * L, dtype=torch.float32, device=dev)
        base = scratch.data_ptr()
        return _DevicePtr(scratch, base), out, _DevicePtr(scratch, base + ws_bytes)


def modality_masks(attn_mask: torch.Tensor, n_vis: int):
    """``[B, n_vis + txt]`` int64 language / image masks (``distillation.py:134-144``) in one launch."""
    lib = cabi.load()
    _require_cuda(attn_mask, "attention_mask")
    am = attn_mask if attn_mask.dtype == torch.int64 else attn_mask.to(torch.int64)
    am = am.contiguous()
    B, txt = am.shape
    both = torch.empty((2, B, n_vis + txt), dtype=torch.int64, device=am.device)
    shape = cabi.make_shape(1, B, n_vis + txt, n_vis, 1, cabi.F32, cabi.LOSS_MSE)
    with _on_device(am.device):
        cabi.check(lib.mafed_distill_modality_masks(ctypes.byref(shape), am.data_ptr(), both[0].data_ptr(),
                                                    both[1].data_ptr(), _stream_ptr(am.device)),
                   "mafed_distill_modality_masks")
    lang, image = both[0], both[1]
    if attn_mask.dtype != torch.int64:
        lang, image = lang.to(attn_mask.dtype), image.to(attn_mask.dtype)
    return lang, image


def modality_masks_into(attn_mask: torch.Tensor, n_vis: int, lang: torch.Tensor, image: torch.Tensor):
    """``modality_masks`` into pre-allocated int64 ``[B, n_vis + txt]`` outputs (``attn_mask`` int64, contiguous)."""
    lib = cabi.load()
    B, txt = attn_mask.shape
    shape = cabi.make_shape(1, B, n_vis + txt, n_vis, 1, cabi.F32, cabi.LOSS_MSE)
    with _on_device(attn_mask.device):
        cabi.check(lib.mafed_distill_modality_masks(ctypes.byref(shape), attn_mask.data_ptr(), lang.data_ptr(),
                                                    image.data_ptr(), _stream_ptr(attn_mask.device)),
                   "mafed_distill_modality_masks")


def token_norm_sums(tensors: Sequence[torch.Tensor], attn_mask: torch.Tensor, n_vis: int) -> torch.Tensor:
    """Masked per-modality sums of the per-token L2 norms of ``len(tensors)`` ``[B, T, D]`` tensors in one
    fused pass (``distillation_loss_weights.py:122-137``).  Returns device fp64 ``[2L + 2]``:
    ``(text sum, vision sum)`` per tensor, then ``(n_text, n_vision)``.  No host synchronisation."""
    lib = cabi.load()
    tensors = _prepare([t.detach() for t in tensors])
    plan = DistillPlan(layers=list(range(len(tensors))), layer_coeffs=[1.0] * len(tensors), n_vis=n_vis)
    ln = _Launch(tensors, tensors, attn_mask, plan)
    L, dev = ln.n_layers, ln.device
    with _on_device(dev):
        stream = _stream_ptr(dev)
        ws = torch.empty(lib.mafed_distill_ws_bytes(L), dtype=torch.uint8, device=dev)
        sums = torch.empty(2 * L + 2, dtype=torch.float64, device=dev)
        cabi.check(lib.mafed_distill_token_norm_sums(ln.shape_ref, ln.s_ptrs, ln.mask_ptr, ws.data_ptr(), stream),
                   "mafed_distill_token_norm_sums")
        cabi.check(cabi.reduce_stage(lib, ln.shape_ref, ln.mask_ptr, ws.data_ptr(), sums.data_ptr(), stream),
                   "mafed_distill_scalar_stage(reduce)")
    return sums


def resolve_group(group):
    """``None``: the default process group if one is initialised with more than one rank;
    ``False``: never communicate; otherwise an explicit ``ProcessGroup``."""
    if group is False:
        return False, None
    dist = torch.distributed
    if group is None or group is True:
        on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        return on, None
    return dist.get_world_size(group) > 1, group


def allreduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """The path's collective: SUM-allreduce of fp64 partial sums and/or token counts (<= 2L+2 values)."""
    torch.distributed.all_reduce(sums, op=torch.distributed.ReduceOp.SUM, group=group)
    return sums


def distill_forward(students, teachers, attn_mask, plan: DistillPlan, group=None):
    """Two-pass step, first half: the fused forward; its last CTA reduces the partial sums and forms the
    losses and the backward scale table (one launch).  Returns ``(out, bwd_scale, launch)``.

    ``out`` is a device fp32 vector ``[1 + 3L]``: total loss, L layer losses (what the reference
    logs to W&B, ``distillation.py:165``), then L x (text loss, vision loss).
    """
    lib = cabi.load()
    ln = _Launch(students, teachers, attn_mask, plan)
    L, dev = ln.n_layers, ln.device
    with _on_device(dev):
        stream = _stream_ptr(dev)
        ws, out, bwd_scale = ln.alloc_scalars(lib)
        w = ctypes.byref(plan.weights())
        distributed, pg = resolve_group(group)
        peer = get_peer_comm(pg) if distributed else None
        if distributed and peer is None:
            cabi.check(lib.mafed_distill_fwd(ln.shape_ref, ln.s_ptrs, ln.t_ptrs, ln.mask_ptr, ws.data_ptr(), stream),
                       "mafed_distill_fwd")
            sums = torch.empty(2 * L + 2, dtype=torch.float64, device=dev)
            cabi.check(cabi.reduce_stage(lib, ln.shape_ref, ln.mask_ptr, ws.data_ptr(), sums.data_ptr(), stream),
                       "mafed_distill_scalar_stage(reduce)")
            allreduce_sums(sums, pg)
            cabi.check(cabi.finalize_stage(lib, ln.shape_ref, w, sums.data_ptr(), out.data_ptr(),
                                           bwd_scale.data_ptr(), stream), "mafed_distill_scalar_stage(finalize)")
        else:
            # forward + reduce + counts (+ NVLink peer allreduce) + losses + scale: one launch, no NCCL call
            sums = torch.empty(2 * L + 2, dtype=torch.float64, device=dev) if peer is not None else None
            cabi.check(lib.mafed_distill_fwd_step(
                ln.shape_ref, ln.s_ptrs, ln.t_ptrs, ln.mask_ptr, w, ws.data_ptr(), out.data_ptr(),
                bwd_scale.data_ptr(), sums.data_ptr() if sums is not None else None,
                peer.handle if peer is not None else None, stream), "mafed_distill_fwd_step")
    return out, bwd_scale, ln


def distill_backward(ln: _Launch, grads: Sequence[Optional[torch.Tensor]], bwd_scale: torch.Tensor,
                     grad_out: Optional[torch.Tensor], skip_if_equals: Optional[float] = None,
                     grad_out_scale: float = 1.0, seen: Optional[torch.Tensor] = None):
    """Fused backward into pre-allocated ``grads`` (``None`` entries are skipped).  With ``skip_if_equals`` the
    launch is the one-pass step's gate: a 1-CTA kernel that does nothing when the upstream gradient (times
    ``grad_out_scale``) equals that value and otherwise starts the backward from the device.  ``seen``: optional
    pinned float tensor that receives the upstream gradient the gate saw."""
    lib = cabi.load()
    g_ptrs = cabi.ptr_array([g.data_ptr() if g is not None else None for g in grads])
    skip = ctypes.byref(ctypes.c_float(skip_if_equals)) if skip_if_equals is not None else None
    with _on_device(ln.device):
        cabi.check(lib.mafed_distill_bwd(ln.shape_ref, ln.s_ptrs, ln.t_ptrs, g_ptrs, ln.mask_ptr,
                                         bwd_scale.data_ptr(), grad_out.data_ptr() if grad_out is not None else None,
                                         float(grad_out_scale), skip, seen.data_ptr() if seen is not None else None,
                                         _stream_ptr(ln.device)), "mafed_distill_bwd")


def _ticket_ptr(ticket):
    """Device address of a counts ticket (``node.prefetch_counts``); the current stream is made to wait for the
    side-stream launch that fills it."""
    if ticket is None:
        return None
    if hasattr(ticket, "wait"):
        ticket.wait()
    return ticket.data_ptr()


def distill_fused(students, teachers, grads, attn_mask, plan: DistillPlan, group=None, mask_out=None, ticket=None):
    """One-pass step: counts -> gradient scale, loss sums + gradients from one read of student and teacher,
    losses -- a single launch of the fused kernel (``mafed_distill_step``); ``mask_out = (lang, image)`` int64
    ``[B, T]`` tensors are filled by the same kernel.  Returns ``(out, bwd_scale, launch)``."""
    lib = cabi.load()
    ln = _Launch(students, teachers, attn_mask, plan)
    L, dev = ln.n_layers, ln.device
    fixed = float(plan.assumed_grad_out) * float(plan.grad_multiplier)
    g_ptrs = cabi.ptr_array([g.data_ptr() if g is not None else None for g in grads])
    with _on_device(dev):
        stream = _stream_ptr(dev)
        ws, out, bwd_scale = ln.alloc_scalars(lib)
        w = ctypes.byref(plan.weights())
        distributed, pg = resolve_group(group)
        peer = get_peer_comm(pg) if distributed else None
        if not distributed or peer is not None:
            # one launch, on one GPU and across batch shards alike: with a communicator the counts are exchanged
            # inside the kernel behind its first tiles and the sums by its last CTA (NVLink peer stores, no NCCL;
            # the epochs live on the device, so the step can be captured into a CUDA graph and replayed)
            sums = torch.empty(2 * L + 2, dtype=torch.float64, device=dev) if peer is not None else None
            lang, image = mask_out if mask_out is not None else (None, None)
            cabi.check(lib.mafed_distill_step(
                ln.shape_ref, ln.s_ptrs, ln.t_ptrs, g_ptrs, ln.mask_ptr, w, fixed, ws.data_ptr(), out.data_ptr(),
                bwd_scale.data_ptr(), sums.data_ptr() if sums is not None else None,
                lang.data_ptr() if lang is not None else None, image.data_ptr() if image is not None else None,
                peer.handle if peer is not None else None, _ticket_ptr(ticket), stream),
                "mafed_distill_step")
            return out, bwd_scale, ln
        if mask_out is not None:
            modality_masks_into(ln.mask, plan.n_vis, *mask_out)
        sums = torch.empty(2 * L + 2, dtype=torch.float64, device=dev)
        # NCCL: counts allreduce -> scale table -> fused pass -> sums allreduce -> losses
        cabi.check(lib.mafed_distill_scalar_stage(ln.shape_ref, None, cabi.STAGE_COUNTS, ln.mask_ptr, None,
                                                  sums.data_ptr(), None, None, None, 0, stream), "counts")
        allreduce_sums(sums[2 * L:], pg)
        cabi.check(lib.mafed_distill_scalar_stage(ln.shape_ref, w, cabi.STAGE_SCALE, None, None, sums.data_ptr(), None,
                                                  bwd_scale.data_ptr(), None, 0, stream), "scale")
        cabi.check(lib.mafed_distill_fused(ln.shape_ref, ln.s_ptrs, ln.t_ptrs, g_ptrs, ln.mask_ptr, None,
                                           bwd_scale.data_ptr(), fixed, ws.data_ptr(), None, stream),
                   "mafed_distill_fused")
        cabi.check(lib.mafed_distill_scalar_stage(ln.shape_ref, None, cabi.STAGE_REDUCE, None, ws.data_ptr(),
                                                  sums.data_ptr(), None, None, None, 0, stream), "reduce")
        allreduce_sums(sums[: 2 * L], pg)
        cabi.check(cabi.finalize_stage(lib, ln.shape_ref, w, sums.data_ptr(), out.data_ptr(), None, stream),
                   "mafed_distill_scalar_stage(finalize)")
    return out, bwd_scale, ln


def _alloc_grads(students, needs: Sequence[bool], cls: bool):
    """Gradient buffers for the layers that need them: ONE allocation carved into per-layer views (the
    caching allocator and the Python overhead are paid once, not L times)."""
    idx = [i for i, need in enumerate(needs) if need]
    if not idx:
        return [None] * len(students)
    s0 = students[idx[0]]
    make = torch.zeros if cls else torch.empty   # cls: only row 0 of each sample is written by the kernel
    buf = make((len(idx),) + tuple(s0.shape), dtype=s0.dtype, device=s0.device)
    views = buf.unbind(0)
    grads = [None] * len(students)
    for j, i in enumerate(idx):
        grads[i] = views[j]
    return grads


class _DistillFunction(torch.autograd.Function):
    """The step as a Python autograd Function over the ctypes calls: used when the ranks have no peer-memory
    communicator (NCCL sequence) -- every other step goes through the compiled node (``_node_step``).  Inputs: (plan, attention_mask, group, teachers tuple, mask_out, *students).  The teachers ride in a plain
    tuple: they never need gradients, so autograd does not have to look at them.  ``mask_out``: ``None`` or the
    pre-allocated ``(lang_masks, image_masks)`` pair to fill (``distillation.py:134-144``)."""

    @staticmethod
    def forward(ctx, plan: DistillPlan, attn_mask, group, teachers, mask_out, *students):
        students = _prepare(students)
        needs = ctx.needs_input_grad[5:]
        ctx.plan, ctx.needs, ctx.grads = plan, needs, None
        if plan.single_pass and any(needs):
            grads = _alloc_grads(students, needs, plan.cls)
            out, bwd_scale, ln = distill_fused(students, teachers, grads, attn_mask, plan, group, mask_out)
            ctx.grads = grads
        else:
            out, bwd_scale, ln = distill_forward(students, teachers, attn_mask, plan, group)
            if mask_out is not None:
                modality_masks_into(ln.mask, plan.n_vis, *mask_out)
        ctx.launch, ctx.bwd_scale = ln, bwd_scale
        # registered with autograd so that an in-place modification of a hidden state between forward and
        # backward is detected (the backward / fix-up kernel would otherwise read the modified values)
        ctx.save_for_backward(*students, *teachers)
        total, aux = out[0], out[1:]
        ctx.mark_non_differentiable(aux)
        return total, aux

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_total, _grad_aux):
        ln, plan = ctx.launch, ctx.plan
        if grad_total is None or ln is None or not any(ctx.needs):
            return (None,) * (5 + len(ctx.needs))
        _ = ctx.saved_tensors   # version-counter check of students / teachers
        g = grad_total
        if g.dtype != torch.float32 or g.device != ln.device:
            g = g.to(device=ln.device, dtype=torch.float32)
        if not g.is_contiguous():
            g = g.contiguous()
        mul = float(plan.grad_multiplier)
        if ctx.grads is not None:
            # one-pass step: gradients already exist; fix them up only if the upstream gradient differs
            grads, ctx.grads = ctx.grads, None
            fixed = float(plan.assumed_grad_out) * mul
            distill_backward(ln, grads, ctx.bwd_scale, g, skip_if_equals=fixed, grad_out_scale=mul)
        else:
            grads = _alloc_grads(ln.students, ctx.needs, plan.cls)
            distill_backward(ln, grads, ctx.bwd_scale, g, grad_out_scale=mul)
        return (None, None, None, None, None, *grads)


# Pinned words that receive the upstream gradient a backward gate saw.  The gate stores into its word when it
# RUNS, which can be after every Python object of the step is gone, so the words are never freed: one pinned
# allocation per process, slots handed out round-robin (a slot shared by two strategies after 1024 of them can
# only mislead the `assumed_grad_out` heuristic, never the results -- the gate itself decides on the device).
_SEEN_POOL = None
_seen_next = 0


def seen_slot() -> torch.Tensor:
    """A pinned float32[1] view for ``distill_loss(..., seen=...)``."""
    global _SEEN_POOL, _seen_next
    if _SEEN_POOL is None:
        _SEEN_POOL = torch.zeros(1024, dtype=torch.float32).pin_memory()
    i = _seen_next % 1024
    _seen_next += 1
    slot = _SEEN_POOL[i:i + 1]
    slot.zero_()
    return slot


def _step(plan: DistillPlan, attn_mask, group, teachers, mask_out, students, ticket=None, seen=None):
    """One launch-sized step (<= MAX_LAYERS layers): the compiled node, or the NCCL sequence without a communicator."""
    _require_cuda(students[0], "hidden_states")
    if attn_mask is not None and not plan.cls:
        _require_cuda(attn_mask, "attention_mask")
    distributed, pg = resolve_group(group)
    peer = get_peer_comm(pg) if distributed else None
    if distributed and peer is None:
        if mask_out is True:
            mask_out = torch.empty((2, attn_mask.shape[0], plan.n_vis + attn_mask.shape[1]), dtype=torch.int64,
                                   device=attn_mask.device)
        both = (mask_out[0], mask_out[1]) if mask_out is not None else None
        if len({t.dtype for t in students} | {t.dtype for t in teachers}) > 1:
            students, teachers = [s.float() for s in students], [t.float() for t in teachers]
        total, aux = _DistillFunction.apply(plan, attn_mask, group, tuple(_prepare(teachers)), both, *students)
        return total, aux, mask_out
    ext = node.load()
    return ext.distill(plan.node_plan(), students, teachers, None if plan.cls else attn_mask, mask_out,
                       peer.handle.value if peer is not None else 0, ticket, seen, cabi.active_tuning_address())


def distill_loss(students: Sequence[torch.Tensor], teachers: Sequence[torch.Tensor], attn_mask, plan: DistillPlan,
                 group=None, teachers_detached: bool = False, mask_out=None, ticket=None, seen=None,
                 return_masks: bool = False):
    """Differentiable fused distillation loss over ``len(students)`` selected layers.

    Returns ``(total, aux)``: ``total`` is the 0-dim fp32 loss (gradients flow to ``students``),
    ``aux`` the non-differentiable ``[3L]`` vector of layer losses then (text, vision) losses.
    ``mask_out``: optional pre-allocated int64 ``[2, B, T]`` tensor, or ``True`` to have one allocated; the step's
    own kernel fills ``[0]`` with ``lang_masks`` and ``[1]`` with ``image_masks`` (``attn_mask`` must then be a
    contiguous int64 CUDA tensor); ``return_masks=True`` appends ``(lang_masks, image_masks)`` (or ``None``) to
    the result.
    ``ticket``: the token counts sent ahead of the step (``prefetch_counts``).  ``seen``: pinned float32 tensor
    that receives the upstream gradient the backward gate saw.
    """
    if not teachers_detached:
        teachers = [t.detach() for t in teachers]
    if len(students) != len(teachers) or len(students) != len(plan.layers):
        raise ValueError("students / teachers / plan.layers length mismatch")
    # (layers of different dtypes: the node up-casts both sides to fp32, as the reference does under autocast,
    # distillation.py:90,244)
    n = len(students)
    if n <= cabi.MAX_LAYERS:
        total, aux, both = _step(plan, attn_mask, group, teachers, mask_out, students, ticket, seen)
        if not return_masks:
            return total, aux
        return total, aux, (None if both is None else (both[0], both[1]))
    # more selected layers than one launch carries (MAFED_MAX_LAYERS): chunk and add the partial totals
    total, layer_losses, modal_losses, both = None, [], [], None
    for lo in range(0, n, cabi.MAX_LAYERS):
        hi = min(n, lo + cabi.MAX_LAYERS)
        sub = DistillPlan(layers=plan.layers[lo:hi], layer_coeffs=plan.layer_coeffs[lo:hi],
                          distill_coeff=plan.distill_coeff, modality_kind=plan.modality_kind,
                          lang_weights=None if plan.lang_weights is None else plan.lang_weights[lo:hi],
                          loss_kind=plan.loss_kind, cls=plan.cls, n_vis=plan.n_vis,
                          grad_multiplier=plan.grad_multiplier, single_pass=plan.single_pass,
                          assumed_grad_out=plan.assumed_grad_out)
        part, aux, masks = _step(sub, attn_mask, group, teachers[lo:hi], mask_out if lo == 0 else None,
                                 students[lo:hi], ticket, seen)
        if lo == 0:
            both = masks
        total = part if total is None else total + part
        layer_losses.append(aux[: hi - lo])
        modal_losses.append(aux[hi - lo:])
    aux = torch.cat(layer_losses + modal_losses)
    if not return_masks:
        return total, aux
    return total, aux, (None if both is None else (both[0], both[1]))
