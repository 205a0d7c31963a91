"""The distillation step as a replayable CUDA graph.

``FeatureDistillation.distill`` + ``loss.backward()`` enqueue two kernels (the fused step and the backward gate; a
third, the count prefetch, across batch shards), but reaching them through Python, the autograd engine and its
device thread costs the host 100-150 us per step -- more than the kernels of a small batch (BASELINE.json
configs[4]: 8 samples per GPU is ~80 us of HBM traffic).  Every piece of the step is capture-safe: no host
synchronisation, no host-side epoch (the exchange counters live on the device), allocations through torch's
caching allocator.  So a trainer that captures its whole training step with ``torch.cuda.graph`` gets the
distillation term for a graph replay's ~4 us of host time; ``GraphedDistillStep`` is that capture for the path
alone, over static hidden-state buffers: copy new activations into ``students`` / ``teachers`` / ``attention_mask``
(or let the producing kernels write there), ``replay()``, read ``loss`` / ``grads``.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


class _Out:
    def __init__(self, hidden_states):
        self.hidden_states = hidden_states


class GraphedDistillStep:
    """``fd.distill(output, batch)`` then ``(loss * grad_out).backward()`` captured once, replayed per step.

    ``students``: the full hidden-state tuple as the model returns it (static buffers; the distilled entries become
    leaves that receive ``.grad``); ``teachers``: the teacher's tuple (static buffers); ``attention_mask``: int64
    ``[B, txt]`` static buffer.  After ``replay()``: ``loss`` (0-dim), ``layer_losses`` (``[3L]``: layer, then
    (text, vision) losses) and ``grads`` (one tensor per distilled layer, same order as ``layers``) hold the
    step's results; they are overwritten by the next replay.
    """

    def __init__(self, fd, students: Sequence[torch.Tensor], teachers: Sequence[torch.Tensor],
                 attention_mask: torch.Tensor, grad_out: float = None, warmup: int = 3):
        self.fd = fd
        self.layers: List[int] = list(fd.loss_weights.get_distillation_layers())
        self.students = [s.detach().requires_grad_(i in self.layers) for i, s in enumerate(students)]
        self.teachers = [t.detach() for t in teachers]
        self.attention_mask = attention_mask
        self.grad_out = fd.assumed_grad_out if grad_out is None else float(grad_out)
        device = self.students[self.layers[0]].device
        out = _Out(tuple(self.students))
        teacher_out = _Out(tuple(self.teachers))
        saved_past = fd.past_model
        fd.past_model = lambda **kw: teacher_out
        try:
            cur = torch.cuda.current_stream(device)
            side = torch.cuda.Stream(device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):      # outside capture: function attributes, allocator warm-up
                    self._clear()
                    self._step(out)
            cur.wait_stream(side)
            torch.cuda.synchronize(device)
            self._clear()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._step(out)
            self.layer_losses = fd.last_layer_losses
            self.grads = [self.students[l].grad for l in self.layers]
        finally:
            fd.past_model = saved_past

    def _clear(self):
        for l in self.layers:
            self.students[l].grad = None

    def _step(self, out):
        batch = {"attention_mask": self.attention_mask}
        self.fd.prefetch_counts(batch)            # (batch-sharded runs; a no-op on one rank)
        loss = self.fd.distill(out, batch)
        (loss if self.grad_out == 1.0 else loss * self.grad_out).backward()
        return loss.detach()

    def replay(self) -> torch.Tensor:
        """Run the captured step on the current contents of the static buffers; returns ``loss``."""
        self.graph.replay()
        return self.loss
