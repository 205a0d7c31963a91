"""Step-level loss aggregation around the distillation path, without Lightning.

The reference decides per micro-batch between the task loss and the replay / distillation loss inside its
LightningModule (``mafed/model/vqa_cont_learner.py:213-236``) and lets Lightning divide by
``accumulate_grad_batches`` before ``backward()``.  ``training_step_loss`` is that decision as a plain
function for trainers that do not subclass the reference module; ``backward_step`` applies the same
scaling, which is exactly the upstream gradient the one-pass distillation step assumes
(``FeatureDistillation.assumed_grad_out = 1 / accumulate_grad_batches``), so the backward fix-up is a no-op.
"""
from __future__ import annotations

from typing import Tuple

import torch


def training_step_loss(cl_method, model, batch, batch_idx: int, task_id: int, replay_interval: int
                       ) -> Tuple[torch.Tensor, str]:
    """Returns ``(loss, log_key)`` like ``VLPythiaVQACLearner.training_step``:

    * ``task_id > 0`` and ``(batch_idx + 1) % replay_interval == 0`` -> ``cl_method.replay(model)`` (the task
      batch is skipped on that step), key ``task_{id}/replay_train_loss``;
    * otherwise, or when replay returns no loss -> the model's LM loss passed through
      ``cl_method.compute_loss``, key ``task_{id}/train_loss``.
    """
    loss = None
    if task_id > 0 and (batch_idx + 1) % replay_interval == 0:
        loss, _ = cl_method.replay(model)
    if loss is not None:
        return loss, f"task_{task_id}/replay_train_loss"
    loss = model(**batch, compute_loss=True, return_dict=True).loss
    return cl_method.compute_loss(model, loss, batch=batch), f"task_{task_id}/train_loss"


def backward_step(loss: torch.Tensor, accumulate_grad_batches: int = 1) -> None:
    """Lightning's manual-accumulation scaling followed by ``backward()``."""
    (loss / accumulate_grad_batches).backward()
