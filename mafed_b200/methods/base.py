"""Strategy interface the trainer drives (mirror of the reference's ``mafed/methods/base.py``).

Only the contract is shared with the reference: attribute names (``task_id``, ``reg_lambda``, ``mask``,
``scaler``, ``update_freq``) and hook names / signatures, because ``mafed/train.py:181-213`` and
``mafed/model/vqa_cont_learner.py:211-254`` call them by name.
"""
from __future__ import annotations


def _accumulation_window(options) -> int:
    """``opts.accumulate_grad_batches`` when present and truthy, else 1 (reference ``base.py:12-15``)."""
    window = getattr(options, "accumulate_grad_batches", None) if options else None
    return window if window else 1


class CLStrategy:
    """A continual-learning method as seen by the training loop."""

    def __init__(self, reg_lambda=1.0, mask=None, scaler=None, **kwargs):
        self.task_id, self.reg_lambda, self.mask, self.scaler = 0, reg_lambda, mask, scaler
        self.update_freq = _accumulation_window(kwargs.get("opts"))

    def _hook(self, **kwargs):
        return None

    # between tasks / after model init / after backward / after an optimizer step: no-ops unless overridden
    update_after_new_task = _hook
    update_after_backward = _hook
    update_after_step = _hook

    def update(self, model, **kwargs):
        self.task_id += 1

    def compute_loss(self, model, loss, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} does not define a task loss")

    def replay(self, model, **kwargs):
        return None, 0          # (replay loss, number of memory examples)

    def _is_batch_after_step(self, batch_idx=0):
        """Does this micro-batch close a gradient-accumulation window?"""
        return not (batch_idx + 1) % self.update_freq


class Naive(CLStrategy):
    """Sequential fine-tuning: nothing is added to the task loss."""

    def compute_loss(self, model, loss, **kwargs):
        return loss
