"""Modality-aware feature distillation on sm_100a kernels.

Mirror of ``mafed/methods/distillation.py:16-257``: same constructor arguments, attributes, hook
methods and return types (``distill`` -> 0-dim loss, ``replay`` -> ``(loss, n_ex)``), so it drops in
behind ``CLMethod["featdistill"]`` (``mafed/train.py:119-134``,
``mafed/model/vqa_cont_learner.py:213-254``).

What changed underneath: the reference loops over layers in Python and, per layer, builds two masks
on the CPU, runs two masked token-loss passes through ~16 ATen kernels and synchronises the host for
W&B.  Here one step is ONE kernel launch over all selected layers (masks, gradient scale, loss sums,
gradients and the loss algebra; ``mafed_distill_step``) plus, in ``backward``, a 1-CTA gate that returns at
once unless the upstream gradient differs from the assumed one (then it starts the exact backward from the
device) -- no host synchronisation; the per-layer values the reference logs (``task_{k}/distill_loss_{layer}``)
stay on the device and are logged a step or two late.  The host side of a step is one call into a compiled
autograd node (``mafed_b200.node``).
"""
from __future__ import annotations

from collections import deque
from copy import deepcopy
from typing import Dict, List, Optional

import numpy as np
import torch

from mafed_b200 import cabi
from mafed_b200 import comm as _comm
from mafed_b200.capture import HiddenStateCapture
from mafed_b200.distill_op import DistillPlan, distill_loss, modality_masks, resolve_group, seen_slot
from mafed_b200.methods.base import CLStrategy
from mafed_b200.methods.distillation_loss_weights import DistillationWeights

try:  # W&B is optional here; the reference requires it
    import wandb as _wandb
except Exception:  # pragma: no cover
    _wandb = None


class _CapturedOutput:
    """Model output whose ``hidden_states`` holds only the captured (distilled) entries."""

    def __init__(self, output, hidden_states):
        self._output = output
        self.hidden_states = hidden_states

    def __getattr__(self, name):
        return getattr(self._output, name)


class FeatureDistillation(CLStrategy):
    """Feature distillation with separate vision / language weights (MAFED)."""

    def __init__(
        self,
        memory_size,
        opts,
        model_type,
        distillation_modality_weighing_strategy="equal",
        distillation_layer_weighing_strategy="single",
        distillation_coeff=1.0,
        replay_coeff=1.0,
        distillation_layer=-1,
        cls_distillation=False,
        distillation_loss="mse",
        gamma: float = 0.8,
        num_hidden_layers: int = 11,
        **kwargs,
    ):
        super().__init__(opts=opts)
        self.opts = opts
        self.model_type = model_type
        # ---- episodic memory bookkeeping (reference :36-46)
        self.memory_size = memory_size
        self.memory_per_task = int(memory_size / (len(opts.tasks) - 1))
        self.batch_size = opts.batch_size
        self.num_workers = 2
        self.seed = 1
        self.datasets = []
        self.rng = np.random.default_rng(opts.seed)
        self.pin_mem = opts.pin_mem
        self.step = 0
        # ---- loss configuration (reference :49-73)
        self.past_model = None
        self.replay_coeff = replay_coeff
        self.distillation_coeff = distillation_coeff
        self.weighing_strategy = distillation_modality_weighing_strategy
        self._cls_distillation = cls_distillation
        self.distillation_loss = "cosine" if distillation_loss == "cosine" else "mse"
        self._loss_kind = cabi.LOSS_COSINE if self.distillation_loss == "cosine" else cabi.LOSS_MSE
        self._compute_distillation_loss = (
            self._compute_cosine_distillation_loss if self._loss_kind == cabi.LOSS_COSINE
            else self._compute_mse_distillation_loss)
        in_range = distillation_layer is not None and 0 <= distillation_layer < num_hidden_layers
        self.loss_weights = DistillationWeights(
            distillation_modality_weighing_strategy=distillation_modality_weighing_strategy,
            distillation_layer_weighing_strategy=distillation_layer_weighing_strategy,
            gamma=gamma,
            num_hidden_layers=num_hidden_layers,
            distillation_layer=distillation_layer if in_range else None,
        )
        self.num_vision_tokens = 256
        # ---- B200 path state
        self.process_group = kwargs.get("process_group")   # None: default group when initialised
        self.populate_batch_masks = True                    # keep the reference's side effect on `batch`
        # one-pass step (loss sums and gradients from a single read of student and teacher); the upstream
        # gradient it bakes in is Lightning's 1/accumulate_grad_batches, checked on the device in backward
        self.single_pass = bool(kwargs.get("single_pass", True))
        self.assumed_grad_out = 1.0 / float(self.update_freq)
        # The backward gate leaves the upstream gradient it really saw in a pinned word; distill() reads it (a plain
        # host read, possibly a step old) and re-aims `assumed_grad_out`, so a trainer that scales the loss
        # differently (a static loss scale, another accumulation factor) pays the exact backward once, not per step.
        # An upstream gradient that keeps changing (dynamic loss scaling) switches the strategy to the two-pass form.
        self.adapt_assumed_grad_out = bool(kwargs.get("adapt_assumed_grad_out", True))
        self._gout_seen = self._gout_seen_np = None
        self._gout_changes = 0
        # Batch-sharded runs return the GLOBAL-batch loss on every rank, so each rank's hidden-state gradient is its
        # share of dL_global/dh.  DistributedDataParallel then AVERAGES parameter gradients over the ranks, which
        # would leave the distillation term world_size times too small next to the (per-rank mean) LM losses.
        # None: multiply by world_size whenever the sharded path is active (right under DDP); pass 1.0 when the
        # gradients are summed, not averaged, across ranks.
        self.grad_multiplier = kwargs.get("grad_multiplier")
        # record only the distilled hidden states with forward hooks instead of output_hidden_states=True
        self.selective_capture = bool(kwargs.get("selective_capture", False))
        # walk one memory-loader iterator instead of spawning a fresh one per replay step
        self.persistent_memory_iterator = bool(kwargs.get("persistent_memory_iterator", False))
        self._mem_iter = self._mem_iter_source = None
        self._mem_epoch = 0
        self._plan_cache = None
        self.last_layer_losses: Optional[torch.Tensor] = None   # device [3L]: layer, then (text, vision)
        self.last_layers: List[int] = []
        self._pending_logs = deque()
        self._tickets = {}                                   # id(attention_mask) -> (mask, its version, ticket)

    # ------------------------------------------------------------------ trainer hooks
    def update(self, dataset, model, dataloader, mask=None, **kwargs):
        self._update_model(model)
        self._update_memory(dataset)
        self.loss_weights.update_weights(model, dataloader, self.task_id)
        self.task_id += 1

    def compute_loss(self, model, loss, batch, **kwargs):
        return loss

    def update_after_new_task(self, model, dataset):
        if self.weighing_strategy != "loss_based":
            return
        self.lang_coeff = self.loss_based_distill.update(
            new_model=model, new_dataset=dataset, memory_dataloader=self.mem_dataloader)

    def update_after_step(self, model, batch_idx=0, on_train_start=False):
        if self.task_id == 0 or self.weighing_strategy != "dynamic":
            return
        if self._is_batch_after_step(batch_idx):
            model.vqa_output_distill_loss_params.update()

    def update_mask(self, mask=None):
        pass

    # ------------------------------------------------------------------ the replay / distillation step
    def replay(self, model):
        """Memory batch -> student forward (hidden states kept) -> replay LM loss + distillation loss.
        Returns ``(loss, n_examples)`` like ``distillation.py:84-103``."""
        batch = self._next_memory_batch()
        n_ex = batch["input_ids"].size(0)
        do_replay = self.replay_coeff > 0 and self.task_id > 0
        loss = None
        if self.distillation_coeff != 0 and not self._cls_distillation:
            # batch-sharded: this rank's token counts leave for the peers now, behind the student forward
            self.prefetch_counts(batch)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if self.selective_capture and self.distillation_coeff != 0:
                # keep only the distilled entries of the hidden-state tuple alive
                with HiddenStateCapture(model, self.loss_weights.get_distillation_layers()) as cap:
                    output = model(**batch, compute_loss=do_replay, output_hidden_states=False, return_dict=True)
                output = _CapturedOutput(output, cap.hidden_states)
            else:
                output = model(**batch, compute_loss=do_replay, output_hidden_states=True, return_dict=True)
            if do_replay:
                loss = self.replay_coeff * output.loss
            if self.distillation_coeff == 0:
                return loss, n_ex
            dloss = self.distill(output=output, batch=batch)
            loss = dloss if loss is None else loss + dloss
        return loss, n_ex

    def _next_memory_batch(self):
        """The reference builds a NEW DataLoader iterator for every replay step
        (``next(iter(self.mem_dataloader))``, ``distillation.py:85``), i.e. the first batch of a fresh
        shuffle -- and a respawn of the loader's worker processes each time.  That stays the default;
        ``persistent_memory_iterator=True`` walks one iterator and re-creates it only when an epoch over
        the memory ends."""
        if not self.persistent_memory_iterator:
            return next(iter(self.mem_dataloader))
        if self._mem_iter is None or self._mem_iter_source is not self.mem_dataloader:
            self._mem_iter, self._mem_iter_source, self._mem_epoch = iter(self.mem_dataloader), self.mem_dataloader, 0
        try:
            return next(self._mem_iter)
        except StopIteration:
            self._mem_epoch += 1
            sampler = getattr(self, "mem_sampler", None)
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(self._mem_epoch)
            self._mem_iter = iter(self.mem_dataloader)
            return next(self._mem_iter)

    def prefetch_counts(self, batch, force: bool = False):
        """Send this rank's token counts to the peers ahead of the step (batch-sharded runs).  The counts depend
        only on ``attention_mask``, which is known as soon as the memory batch is drawn -- before the student
        forward (``distillation.py:85-91``) -- so the step's kernel later finds the global counts in its own
        mailbox instead of exchanging them at its start.  The 1-CTA launch goes to a side stream (ordered behind the
        producer of the mask), so it does not sit between kernels of the compute stream.  ``replay`` calls this itself;
        a trainer that draws batches ahead may call it up to two steps early (every rank alike: the call is part of
        the exchange sequence).  No-op on a single rank unless ``force``."""
        attn = batch.get("attention_mask") if hasattr(batch, "get") else None
        if attn is None or not attn.is_cuda or attn.dtype != torch.int64 or not attn.is_contiguous():
            return None
        peer = self._peer()
        if peer is None and not force:
            return None
        from mafed_b200 import node
        ticket = node.load().prefetch_counts(attn, self.num_vision_tokens, peer.handle.value if peer is not None else 0,
                                             cabi.active_tuning_address())
        if len(self._tickets) >= 3:          # at most 3 batches ahead (the mailbox keeps 4 generations of counts)
            self._tickets.pop(next(iter(self._tickets)))
        self._tickets[id(attn)] = (attn, attn._version, ticket)
        return ticket

    def capture(self, hidden_states, past_hidden_states, attention_mask, grad_out=None, warmup: int = 3):
        """``distill`` + ``backward`` over static hidden-state buffers as a replayable CUDA graph
        (``mafed_b200.graphed.GraphedDistillStep``): every launch of the step is capture-safe, and a replay costs the
        host ~4 us instead of the ~150 us of Python and autograd-engine time of the eager calls.  ``hidden_states`` /
        ``past_hidden_states``: the student's and the teacher's full tuples; fill them (or let the producing kernels
        write there), ``step.replay()``, read ``step.loss`` / ``step.grads``."""
        from mafed_b200.graphed import GraphedDistillStep
        return GraphedDistillStep(self, hidden_states, past_hidden_states, attention_mask, grad_out=grad_out,
                                  warmup=warmup)

    def _peer(self):
        """The peer-memory communicator of a batch-sharded run (None: single rank, or the NCCL sequence)."""
        distributed, pg = resolve_group(self.process_group)
        if not distributed:
            return None
        from mafed_b200.comm import get_peer_comm
        return get_peer_comm(pg)

    def distill(self, output, batch):
        """Sum over the selected layers of ``layer_coeff * distillation_coeff * layer_loss``
        (``distillation.py:105-122``) -- as one fused launch instead of a Python loop."""
        layers = self.loss_weights.get_distillation_layers()
        # the teacher's states come out of a no_grad forward and the node never differentiates them: the per-tensor
        # `.detach()` of the reference (`_get_past_hidden_states`, distillation.py:223) is only paid for the distilled
        # layers, and only if a tensor carries a graph at all
        past_hidden_states = self._past_states(batch, layers)
        if self.adapt_assumed_grad_out and self._gout_seen_np is not None:
            self._adapt_assumed()
        plan = self._step_plan(layers)
        hidden = output.hidden_states
        total, aux = self._launch(plan, batch, [hidden[l] for l in layers], past_hidden_states, teachers_detached=True)
        self._record(aux, layers)
        self.step += 1
        return total

    def _step_plan(self, layers) -> DistillPlan:
        """The plan of the all-layer step, rebuilt only when something it depends on changes (the host
        tables cost ~50 us of Python per step otherwise)."""
        lw = self.loss_weights
        coeff = getattr(lw, "lang_coeff", None)
        key = (tuple(layers), self.distillation_coeff, self._cls_distillation, self._loss_kind, self.num_vision_tokens,
               self.single_pass, self.assumed_grad_out, lw._modality_weighing_strategy, id(coeff),
               getattr(coeff, "_version", None), id(lw.layer_coeffs), self.grad_multiplier, self.process_group)
        if self._plan_cache is None or self._plan_cache[0] != key:
            coeffs, modality_kind, lang_weights = self._tables(layers)
            plan = self._plan(layers, coeffs, self.distillation_coeff, modality_kind, lang_weights)
            if len(layers) <= cabi.MAX_LAYERS:
                plan.weights()
            self._plan_cache = (key, plan)
        return self._plan_cache[1]

    def feature_distillation(self, batch, hidden_states, past_hidden_states, layer: int):
        """Layer loss ``w_text * loss_text + w_vision * loss_vision`` of one layer, without the layer
        coefficient (``distillation.py:124-166``)."""
        coeffs, modality_kind, lang_weights = self._tables([layer])
        plan = self._plan([layer], [1.0], 1.0, modality_kind, lang_weights)
        total, aux = self._launch(plan, batch, [hidden_states], [past_hidden_states])
        if not self._cls_distillation:
            self._record(aux, [layer])
        return total

    # ------------------------------------------------------------------ internals
    def _tables(self, layers):
        if self._cls_distillation:
            coeffs = [float(self.loss_weights.get_layer_loss_weight(l)) for l in layers]
            return coeffs, cabi.MODW_CLS, None
        return self.loss_weights.kernel_tables(layers)

    def _plan(self, layers, coeffs, distill_coeff, modality_kind, lang_weights) -> DistillPlan:
        if self._cls_distillation and self._loss_kind != cabi.LOSS_COSINE:
            # quirk kept from the reference: MSELoss takes two arguments, the CLS branch passes three
            raise TypeError("cls_distillation requires distillation_loss='cosine' "
                            "(MSELoss.forward() takes 3 positional arguments but 4 were given)")
        mul = self.grad_multiplier
        if mul is None:
            distributed, pg = resolve_group(self.process_group)
            mul = float(torch.distributed.get_world_size(pg)) if distributed else 1.0
        return DistillPlan(layers=list(layers), layer_coeffs=list(coeffs), distill_coeff=float(distill_coeff),
                           modality_kind=modality_kind, lang_weights=lang_weights, loss_kind=self._loss_kind,
                           cls=bool(self._cls_distillation), n_vis=self.num_vision_tokens,
                           single_pass=self.single_pass, assumed_grad_out=self.assumed_grad_out,
                           grad_multiplier=float(mul))

    def _adapt_assumed(self):
        """Re-aim the one-pass step at the upstream gradient the gate last saw (see ``__init__``)."""
        seen = float(self._gout_seen_np[0])
        if seen == 0.0 or seen != seen or seen in (float("inf"), float("-inf")):
            return
        mul = self._plan_cache[1].grad_multiplier if self._plan_cache is not None else 1.0
        want = seen / mul
        if abs(want - self.assumed_grad_out) <= 1e-7 * abs(want):
            return
        self.assumed_grad_out = want
        self._gout_changes += 1
        if self._gout_changes > 8 and self.single_pass:
            # the upstream gradient is not a constant of this run: stop guessing
            self.single_pass = False

    def _launch(self, plan: DistillPlan, batch, students, teachers, teachers_detached=False):
        attn, mask_out, ticket = None, None, None
        if not plan.cls:
            attn = batch["attention_mask"]
            if self.populate_batch_masks:
                if attn.is_cuda and attn.dtype == torch.int64 and attn.is_contiguous():
                    # the masks the reference leaves in `batch` are allocated and written by the step itself
                    mask_out = True
                else:
                    batch["lang_masks"], batch["image_masks"] = modality_masks(attn, self.num_vision_tokens)
            if self._tickets:
                tk = self._tickets.pop(id(attn), None)
                if tk is not None and tk[0] is attn and tk[1] == attn._version:
                    ticket = tk[2]
        if plan.single_pass and self._gout_seen is None:
            self._gout_seen = seen_slot()
            self._gout_seen_np = self._gout_seen.numpy()
        total, aux, masks = distill_loss(students, teachers, attn, plan, group=self.process_group,
                                         teachers_detached=teachers_detached, mask_out=mask_out, ticket=ticket,
                                         seen=self._gout_seen if plan.single_pass else None, return_masks=True)
        if masks is not None:
            batch["lang_masks"], batch["image_masks"] = masks
        return total, aux

    def _past_states(self, batch, layers):
        """Teacher hidden states of the distilled layers only (same forward and side effects as
        ``_get_past_hidden_states``)."""
        with torch.no_grad():
            batch.pop("labels", None)
            if self.selective_capture:
                with HiddenStateCapture(self.past_model, layers, detach=True) as cap:
                    self.past_model(**batch, output_hidden_states=False, return_dict=True)
                states = cap.hidden_states
            else:
                states = self.past_model(**batch, output_hidden_states=True, return_dict=True).hidden_states
        # no per-tensor `.detach()` (distillation.py:223): these come out of a no_grad forward, and the node never builds
        # a graph edge to a teacher tensor whatever its flags are
        return [states[l] for l in layers]

    def _get_past_hidden_states(self, batch):
        with torch.no_grad():
            batch.pop("labels", None)
            if self.selective_capture:
                layers = self.loss_weights.get_distillation_layers()
                with HiddenStateCapture(self.past_model, layers, detach=True) as cap:
                    self.past_model(**batch, output_hidden_states=False, return_dict=True)
                return cap.hidden_states
            states = self.past_model(**batch, output_hidden_states=True, return_dict=True).hidden_states
        return [h.detach() for h in states]

    # ---- single-tensor token losses, kept for API compatibility (distillation.py:226-257)
    def _masked_token_loss(self, hidden_states, past_hidden_states, mask, loss_kind):
        dim = hidden_states.shape[-1]
        h = hidden_states.reshape(1, -1, dim)
        p = past_hidden_states.reshape(1, -1, dim)
        plan = DistillPlan(layers=[0], layer_coeffs=[1.0], distill_coeff=1.0, modality_kind=cabi.MODW_TEXT_ONLY,
                           loss_kind=loss_kind, n_vis=0)
        total, _ = distill_loss([h], [p.detach()], mask.reshape(1, -1), plan, group=False)
        return total

    def _compute_cosine_distillation_loss(self, hidden_states, past_hidden_states, mask):
        return self._masked_token_loss(hidden_states, past_hidden_states, mask, cabi.LOSS_COSINE)

    def _compute_mse_distillation_loss(self, hidden_states, past_hidden_states, mask):
        return self._masked_token_loss(hidden_states, past_hidden_states, mask, cabi.LOSS_MSE)

    def _compute_cls_distillation_loss(self, hidden_states, past_hidden_states):
        if self._loss_kind != cabi.LOSS_COSINE:
            raise TypeError("MSELoss.forward() takes 3 positional arguments but 4 were given")
        plan = DistillPlan(layers=[0], layer_coeffs=[1.0], distill_coeff=1.0, modality_kind=cabi.MODW_CLS,
                           loss_kind=self._loss_kind, cls=True, n_vis=self.num_vision_tokens)
        total, _ = distill_loss([hidden_states], [past_hidden_states.detach()], None, plan, group=False)
        return total

    # ------------------------------------------------------------------ logging without host syncs
    def _record(self, aux: torch.Tensor, layers: List[int]):
        """Keep the per-layer losses on the device; hand earlier steps' values to W&B once they have landed in
        pinned memory (the reference calls ``.item()`` per layer, ``distillation.py:165``).  Every step's values
        are logged, in order: a step whose copy has not completed yet stays queued."""
        self.last_layer_losses = aux
        self.last_layers = list(layers)
        if self.process_group is not False and _comm._cache:
            self.check_exchange(sync=False)
        if _wandb is None or getattr(_wandb, "run", None) is None:
            return
        if torch.cuda.is_current_stream_capturing():
            return      # inside a CUDA-graph capture: the values stay in `last_layer_losses` (static per replay)
        self.flush_logs(wait=False)
        host = torch.empty(len(layers), dtype=torch.float32, pin_memory=True)
        host.copy_(aux[: len(layers)], non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(aux.device))
        self._pending_logs.append((host, done, list(layers), self.task_id))
        while len(self._pending_logs) > 64:     # the host is far ahead of the device: wait for the oldest step
            self._flush_one(wait=True)

    def _flush_one(self, wait: bool) -> bool:
        host, done, layers, task_id = self._pending_logs[0]
        if not wait and not done.query():
            return False
        done.synchronize()
        self._pending_logs.popleft()
        if _wandb is not None and getattr(_wandb, "run", None) is not None:
            # one `wandb.log` per layer, like the reference (distillation.py:165): W&B's step axis advances L times
            # per replay step either way
            for l, v in zip(layers, host.tolist()):
                _wandb.log({f"task_{task_id}/distill_loss_{l}": float(v)})
        return True

    def flush_logs(self, wait: bool = True):
        """Send the queued per-layer values to W&B (``task_{id}/distill_loss_{layer}``), oldest first."""
        while self._pending_logs and self._flush_one(wait):
            pass

    def check_exchange(self, sync: bool = True):
        """Batch-sharded runs: raise if a peer missed an in-kernel exchange (that step's loss and gradients are NaN;
        every rank must run the same number of distillation steps).  The status word lives in mapped pinned host
        memory, so reading it costs nothing: ``distill`` does it every step (``sync=False``: it sees every step
        that has FINISHED, i.e. a missed exchange raises one step late at the latest, before the poisoned
        gradients can survive an optimizer step unnoticed); ``sync=True`` waits for the queued steps first."""
        group = self.process_group
        if group is False:
            return
        peer = _comm.peek_peer_comm(None if group is None or group is True else group)
        if peer is not None:
            if sync:
                torch.cuda.synchronize()
            peer.check()

    def layer_loss_dict(self) -> Dict[str, float]:
        """The reference's W&B payload for the last step (synchronises the host)."""
        if self.last_layer_losses is None:
            return {}
        vals = self.last_layer_losses[: len(self.last_layers)].tolist()
        self.check_exchange()
        return {f"task_{self.task_id}/distill_loss_{l}": v for l, v in zip(self.last_layers, vals)}

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self) -> Dict[str, object]:
        """Strategy state that the reference keeps only in the live Python object across the task loop
        (``train.py:207``; SURVEY 5 "checkpoint / resume"): task counter, step counter, the sampled memory
        indices per task, the sampling RNG, and the adaptive language coefficients."""
        coeff = getattr(self.loss_weights, "lang_coeff", None)
        return {
            "task_id": self.task_id,
            "step": self.step,
            "memory_indices": [np.asarray(ds.indices).tolist() for ds in self.datasets],
            "rng_state": self.rng.bit_generator.state,
            "lang_coeff": coeff.detach().cpu() if torch.is_tensor(coeff) else coeff,
        }

    def load_state_dict(self, state: Dict[str, object], datasets=None):
        """Restore ``state_dict()``.  ``datasets`` (one per finished task, in order) lets the episodic memory be
        rebuilt from the saved indices; the teacher snapshot itself comes from the model checkpoint
        (``_update_model``)."""
        self.task_id = int(state["task_id"])
        self.step = int(state["step"])
        self.rng.bit_generator.state = state["rng_state"]
        coeff = state.get("lang_coeff")
        if coeff is not None:
            if torch.is_tensor(coeff) and torch.cuda.is_available():
                coeff = coeff.cuda()     # `update_weights` averages it with importances computed on the device
            self.loss_weights.lang_coeff = coeff
            self.loss_weights._lang_coeff_host = None
        self._plan_cache = None
        if datasets is not None:
            from torch.utils.data import Subset
            self.datasets = [Subset(ds, idx) for ds, idx in zip(datasets, state["memory_indices"])]
            if self.datasets:
                self._rebuild_loader()

    # ------------------------------------------------------------------ between tasks
    def _update_model(self, model):
        """Freeze a copy of the just-trained model as the next task's teacher."""
        self.past_model = deepcopy(model)
        self.past_model.eval()

    def _update_memory(self, dataset):
        """Add ``memory_per_task`` random samples of the finished task to the episodic memory and
        rebuild its loader (``distillation.py:182-209``).  Data loading itself is the reference's
        (``mafed.data``); it is imported lazily because it is outside this package's scope."""
        from torch.utils.data import Subset
        if hasattr(self, "mem_dataloader"):
            del self.mem_dataloader
        picked = self.rng.choice(np.arange(len(dataset)), self.memory_per_task, replace=False)
        assert len(set(picked)) == self.memory_per_task
        self.seed = 1
        self.datasets.append(Subset(dataset, picked))
        self._rebuild_loader()

    def _rebuild_loader(self):
        """(Re)create the memory loader over ``self.datasets`` (``distillation.py:192-209``)."""
        from torch.utils.data import ConcatDataset, DataLoader, RandomSampler
        from torch.utils.data.distributed import DistributedSampler
        import torch.distributed as dist

        memory = ConcatDataset(self.datasets)
        distributed = dist.is_available() and dist.is_initialized()
        sampler = DistributedSampler(memory) if distributed else RandomSampler(memory)
        collate, prefetch = self._data_hooks()
        loader = DataLoader(memory, sampler=sampler, num_workers=self.num_workers, batch_size=self.batch_size,
                            collate_fn=collate)
        self.mem_dataloader = prefetch(loader)
        self.mem_sampler = sampler

    def _data_hooks(self):
        hooks = getattr(self, "data_hooks", None)
        if hooks is not None:
            return hooks
        try:
            from mafed.data import PrefetchLoader, collate_fn
        except ImportError as exc:  # pragma: no cover - depends on the host installation
            raise ImportError(
                "FeatureDistillation._update_memory needs the reference's `mafed.data` (collate_fn, "
                "PrefetchLoader) or `self.data_hooks = (collate_fn, loader_wrapper)`") from exc
        return collate_fn["train"][self.model_type], PrefetchLoader
