"""The fused distillation operator: host tables -> C ABI -> sm_100a kernels, with autograd.

One step = one fused forward launch over all selected layers (+ a single-CTA epilogue), one fused
backward launch.  Under ``torch.distributed`` with batch sharding, the per-rank ``[2L+2]`` partial
sums are combined by a single allreduce between the forward and the epilogue; the backward needs no
collective.  Nothing here synchronises the host with the device.

Replaces the per-layer Python loop of ``mafed/methods/distillation.py:105-166`` and the autograd
chains behind ``:226-249``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import cabi

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
    _weights: Optional[cabi.Weights] = field(default=None, repr=False)

    def weights(self) -> cabi.Weights:
        if self._weights is None:
            if len(self.layers) > cabi.MAX_LAYERS:
                raise ValueError(f"at most {cabi.MAX_LAYERS} layers per call")
            self._weights = cabi.make_weights(
                cabi.MODW_CLS if self.cls else self.modality_kind, self.distill_coeff,
                self.layer_coeffs, self.lang_weights)
        return self._weights


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise cabi.MafedDistillError(
            f"{what} is on {t.device}: the distillation path runs only as sm_100a CUDA kernels "
            "(there is no CPU fallback)")


def _prepare(tensors: Sequence[torch.Tensor], dtype=None):
    out = []
    for t in tensors:
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        out.append(t if t.is_contiguous() else t.contiguous())
    return out


class _Launch:
    """Geometry + pointer tables of one step; keeps the tensors alive until the call returns."""

    def __init__(self, students, teachers, attn_mask, plan: DistillPlan):
        s0 = students[0]
        _require_cuda(s0, "hidden_states")
        if s0.dtype not in _DTYPES:
            raise TypeError(f"unsupported hidden-state dtype {s0.dtype} (float32, bfloat16, float16)")
        if s0.dim() != 3:
            raise ValueError("hidden states must be [B, T, D]")
        self.B, self.T, self.D = s0.shape
        for s, t in zip(students, teachers):
            if s.shape != s0.shape or t.shape != s0.shape:
                raise ValueError("all selected hidden states must share one [B, T, D] shape")
            if s.dtype != s0.dtype or t.dtype != s0.dtype or t.device != s0.device:
                raise ValueError("student / teacher dtype or device mismatch")
        self.device = s0.device
        self.dtype = s0.dtype
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
        self.students = students
        self.teachers = teachers
        self.s_ptrs = cabi.ptr_array([t.data_ptr() for t in students])
        self.t_ptrs = cabi.ptr_array([t.data_ptr() for t in teachers])
        self.mask_ptr = self.mask.data_ptr() if self.mask is not None else None


def distill_forward(students, teachers, attn_mask, plan: DistillPlan, group=None):
    """Run the fused forward.  Returns ``(out, bwd_scale, launch)``.

    ``out`` is a device fp32 vector ``[1 + 3L]``: total loss, L layer losses (what the reference
    logs to W&B, ``distillation.py:165``), then L x (text loss, vision loss).
    """
    lib = cabi.load()
    ln = _Launch(students, teachers, attn_mask, plan)
    L = ln.n_layers
    dev = ln.device
    with torch.cuda.device(dev):
        stream = _stream_ptr(dev)
        ws = torch.empty(lib.mafed_distill_ws_bytes(L), dtype=torch.uint8, device=dev)
        out = torch.empty(1 + 3 * L, dtype=torch.float32, device=dev)
        bwd_scale = torch.empty(2 * L, dtype=torch.float32, device=dev)
        import ctypes
        cabi.check(lib.mafed_distill_fwd(ctypes.byref(ln.shape), ln.s_ptrs, ln.t_ptrs, ln.mask_ptr, ws.data_ptr(),
                                         stream), "mafed_distill_fwd")
        w = plan.weights()
        distributed, pg = resolve_group(group)
        if distributed:
            sums = torch.empty(2 * L + 2, dtype=torch.float64, device=dev)
            cabi.check(lib.mafed_distill_reduce(ctypes.byref(ln.shape), ln.mask_ptr, ws.data_ptr(), sums.data_ptr(),
                                                stream), "mafed_distill_reduce")
            allreduce_sums(sums, pg)
            cabi.check(lib.mafed_distill_finalize(ctypes.byref(ln.shape), ctypes.byref(w), sums.data_ptr(),
                                                  out.data_ptr(), bwd_scale.data_ptr(), stream),
                       "mafed_distill_finalize")
        else:
            cabi.check(lib.mafed_distill_epilogue(ctypes.byref(ln.shape), ctypes.byref(w), ln.mask_ptr, ws.data_ptr(),
                                                  None, out.data_ptr(), bwd_scale.data_ptr(), stream),
                       "mafed_distill_epilogue")
    return out, bwd_scale, ln


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
    """The path's only collective: SUM-allreduce of the ``[2L+2]`` fp64 partial sums + token counts."""
    torch.distributed.all_reduce(sums, op=torch.distributed.ReduceOp.SUM, group=group)
    return sums


def distill_backward(ln: _Launch, grads: Sequence[Optional[torch.Tensor]], bwd_scale: torch.Tensor,
                     grad_out: Optional[torch.Tensor]):
    """Run the fused backward into pre-allocated ``grads`` (``None`` entries are skipped)."""
    import ctypes
    lib = cabi.load()
    g_ptrs = cabi.ptr_array([g.data_ptr() if g is not None else None for g in grads])
    with torch.cuda.device(ln.device):
        cabi.check(lib.mafed_distill_bwd(ctypes.byref(ln.shape), ln.s_ptrs, ln.t_ptrs, g_ptrs, ln.mask_ptr,
                                         bwd_scale.data_ptr(), grad_out.data_ptr() if grad_out is not None else None,
                                         _stream_ptr(ln.device)), "mafed_distill_bwd")


class _DistillFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: DistillPlan, attn_mask, group, n, *tensors):
        students = _prepare(tensors[:n])
        teachers = _prepare(tensors[n:])
        out, bwd_scale, ln = distill_forward(students, teachers, attn_mask, plan, group)
        ctx.launch = ln
        ctx.plan = plan
        ctx.bwd_scale = bwd_scale
        ctx.n = n
        total = out[0]
        aux = out[1:]
        ctx.mark_non_differentiable(aux)
        return total, aux

    @staticmethod
    def backward(ctx, grad_total, _grad_aux):
        ln, plan, n = ctx.launch, ctx.plan, ctx.n
        if grad_total is None:
            return (None,) * (4 + 2 * n)
        g = grad_total
        if g.dtype != torch.float32 or not g.is_cuda:
            g = g.to(device=ln.device, dtype=torch.float32)
        if plan.grad_multiplier != 1.0:
            g = g * plan.grad_multiplier
        g = g.contiguous()
        grads = []
        for i, s in enumerate(ln.students):
            if not ctx.needs_input_grad[4 + i]:
                grads.append(None)
            elif plan.cls:
                grads.append(torch.zeros_like(s))  # only row 0 of each sample is written by the kernel
            else:
                grads.append(torch.empty_like(s))
        if any(x is not None for x in grads):
            distill_backward(ln, grads, ctx.bwd_scale, g)
        ctx.launch = None
        return (None, None, None, None, *grads, *([None] * n))


def distill_loss(students: Sequence[torch.Tensor], teachers: Sequence[torch.Tensor], attn_mask, plan: DistillPlan,
                 group=None):
    """Differentiable fused distillation loss over ``len(students)`` selected layers.

    Returns ``(total, aux)``: ``total`` is the 0-dim fp32 loss (gradients flow to ``students``),
    ``aux`` the non-differentiable ``[3L]`` vector of layer / modality losses.
    """
    students = list(students)
    teachers = [t.detach() for t in teachers]
    if len(students) != len(teachers) or len(students) != len(plan.layers):
        raise ValueError("students / teachers / plan.layers length mismatch")
    dt = students[0].dtype
    if any(t.dtype != dt for t in students) or any(t.dtype != dt for t in teachers):
        # mixed dtypes: the reference up-casts both sides to fp32 under autocast (distillation.py:90,244)
        students = [s.float() for s in students]
        teachers = [t.float() for t in teachers]
    return _DistillFunction.apply(plan, attn_mask, group, len(students), *students, *teachers)
