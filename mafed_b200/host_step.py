"""Host-buffer entry to the distillation step: inputs and outputs live in (pinned) host memory.

This is the end-to-end form of the path for callers whose hidden states are not already on the
device: every step copies student / teacher / mask host -> device, runs the one-pass step and copies
the gradients and the loss device -> host.  Layers are pipelined over three streams (copy-in, compute,
copy-out) so that PCIe transfers in both directions overlap each other and the kernels; per layer the
work is one C-ABI call (``mafed_distill_step`` with ``n_layers = 1``: one kernel launch, including the
cross-rank exchange of a batch-sharded run).

The backward of a layer needs only the token counts and the host weight tables, not the other
layers' sums, which is what makes the per-layer pipeline legal.
"""
from __future__ import annotations

from typing import List

import torch

from mafed_b200.distill_op import distill_fused


class HostStep:
    def __init__(self, method, students, teachers, attention_mask, device, layers=None):
        """``method``: a ``FeatureDistillation``; tensors may live anywhere -- they are copied once into
        pinned host buffers, which are the step's real inputs."""
        self.method = method
        self.device = torch.device(device)
        self.layers = list(layers) if layers is not None else list(range(len(students)))
        pin = dict(pin_memory=True)
        self.h_s = [torch.empty(s.shape, dtype=s.dtype, **pin).copy_(s) for s in students]
        self.h_t = [torch.empty(t.shape, dtype=t.dtype, **pin).copy_(t) for t in teachers]
        self.h_mask = torch.empty(attention_mask.shape, dtype=torch.int64, **pin).copy_(attention_mask)
        self.h_g = [torch.empty(s.shape, dtype=s.dtype, **pin) for s in students]
        self.h_loss = torch.empty((), dtype=torch.float32, **pin)
        self.h_layer_losses = torch.empty(len(students), dtype=torch.float32, **pin)
        dev = self.device
        self.d_s = [torch.empty(s.shape, dtype=s.dtype, device=dev) for s in students]
        self.d_t = [torch.empty(s.shape, dtype=s.dtype, device=dev) for s in students]
        self.d_g = [torch.empty(s.shape, dtype=s.dtype, device=dev) for s in students]
        self.d_mask = torch.empty(attention_mask.shape, dtype=torch.int64, device=dev)
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in students]
        self.ev_done = [torch.cuda.Event() for _ in students]
        self.gout = torch.ones((), dtype=torch.float32, device=dev)
        nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)
        self.h2d_bytes = nbytes(self.h_s) + nbytes(self.h_t) + nbytes([self.h_mask])
        self.d2h_bytes = nbytes(self.h_g) + 4 * (1 + len(students))
        self.note = "pinned host student/teacher/mask -> device, one-pass fused kernel per layer, gradients+loss -> pinned host; " \
                    "3-stream layer pipeline"
        coeffs, kind, lang = method._tables(self.layers)
        self.plans = []
        for i, l in enumerate(self.layers):
            plan = method._plan([l], [coeffs[i]], method.distillation_coeff, kind, None if lang is None else [lang[i]])
            plan.assumed_grad_out = 1.0
            self.plans.append(plan)

    def step(self) -> torch.Tensor:
        """One end-to-end step.  Returns the pinned host 0-dim loss (valid on return)."""
        cur = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(cur)
        with torch.cuda.stream(self.s_in):
            self.d_mask.copy_(self.h_mask, non_blocking=True)
            for i in range(len(self.layers)):
                self.d_s[i].copy_(self.h_s[i], non_blocking=True)
                self.d_t[i].copy_(self.h_t[i], non_blocking=True)
                self.ev_in[i].record(self.s_in)
        totals: List[torch.Tensor] = []
        outs = []
        for i in range(len(self.layers)):
            cur.wait_event(self.ev_in[i])
            # upstream gradient is a host value here (1.0), so the one-pass kernel's result is final
            out, _, _ = distill_fused([self.d_s[i]], [self.d_t[i]], [self.d_g[i]], self.d_mask, self.plans[i],
                                      group=self.method.process_group)
            self.ev_done[i].record(cur)
            outs.append(out)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_done[i])
                self.h_g[i].copy_(self.d_g[i], non_blocking=True)
        stacked = torch.stack(outs)                     # [L, 4]: total_l, layer_l, text_l, vision_l
        total = stacked[:, 0].double().sum().float()
        self.s_out.wait_stream(cur)
        with torch.cuda.stream(self.s_out):
            self.h_loss.copy_(total, non_blocking=True)
            self.h_layer_losses.copy_(stacked[:, 1], non_blocking=True)
        cur.wait_stream(self.s_out)
        self.s_out.synchronize()                        # results are on the host when step() returns
        return self.h_loss


class CHostStep:
    """The same end-to-end step through the C ABI alone (``mafed_host_step_*``): host pointers in, host
    pointers out, the copy / kernel pipeline runs inside the library on its own three streams.  torch is used
    here only to own the pinned host buffers."""

    def __init__(self, method, students, teachers, attention_mask, device, layers=None):
        import ctypes

        from mafed_b200 import cabi
        from mafed_b200.distill_op import _DTYPES
        self.lib = cabi.load()
        self.device = torch.device(device)
        self.layers = list(layers) if layers is not None else list(range(len(students)))
        pin = dict(pin_memory=True)
        self.h_s = [torch.empty(s.shape, dtype=s.dtype, **pin).copy_(s) for s in students]
        self.h_t = [torch.empty(t.shape, dtype=t.dtype, **pin).copy_(t) for t in teachers]
        self.h_mask = torch.empty(attention_mask.shape, dtype=torch.int64, **pin).copy_(attention_mask)
        self.h_g = [torch.empty(s.shape, dtype=s.dtype, **pin) for s in students]
        L = len(students)
        self.h_out = torch.empty(1 + 3 * L, dtype=torch.float32, **pin)
        B, T, D = students[0].shape
        coeffs, kind, lang = method._tables(self.layers)
        plan = method._plan(self.layers, coeffs, method.distillation_coeff, kind, lang)
        self.weights = plan.weights()
        self.shape = cabi.make_shape(L, B, T, method.num_vision_tokens, D, _DTYPES[students[0].dtype], plan.loss_kind)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            cabi.check(self.lib.mafed_host_step_create(ctypes.byref(self.shape), ctypes.byref(self.handle)),
                       "mafed_host_step_create")
        self.s_ptrs = cabi.ptr_array([t.data_ptr() for t in self.h_s])
        self.t_ptrs = cabi.ptr_array([t.data_ptr() for t in self.h_t])
        self.g_ptrs = cabi.ptr_array([t.data_ptr() for t in self.h_g])
        nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)
        self.h2d_bytes = nbytes(self.h_s) + nbytes(self.h_t) + nbytes([self.h_mask])
        self.d2h_bytes = nbytes(self.h_g) + 4 * (1 + 3 * L)
        self.note = "C ABI mafed_host_step_run: pinned host student/teacher/mask -> device, one-pass fused kernel " \
                    "per layer, gradients + losses -> pinned host; 3-stream layer pipeline inside the library"
        # batch-sharded runs hand the library the same peer-memory communicator the device-resident step uses
        peer = method._peer() if hasattr(method, "_peer") else None
        self.comm = peer.handle if peer is not None else None
        if peer is not None:
            self.note += "; batch-sharded: counts sent ahead once per step, sums exchanged per layer inside the kernels"

    def step(self, grad_out: float = 1.0) -> torch.Tensor:
        import ctypes

        from mafed_b200 import cabi
        with torch.cuda.device(self.device):
            cabi.check(self.lib.mafed_host_step_run(self.handle, ctypes.byref(self.weights), self.s_ptrs, self.t_ptrs,
                                                    self.g_ptrs, self.h_mask.data_ptr(), float(grad_out),
                                                    self.h_out.data_ptr(), self.comm), "mafed_host_step_run")
        return self.h_out[0]

    def close(self):
        if self.handle:
            self.lib.mafed_host_step_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
