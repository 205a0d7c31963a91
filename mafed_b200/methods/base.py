"""Continual-learning strategy interface -- mirror of ``mafed/methods/base.py:1-57``.

Same attribute names and hook signatures as the reference so that ``mafed/train.py`` and
``mafed/model/vqa_cont_learner.py`` drive either implementation unchanged.
"""


class CLStrategy:
    """Base class of the continual-learning strategies (``mafed/methods/base.py:1-47``).

    Attributes read by the callers: ``task_id``, ``reg_lambda``, ``mask``, ``scaler``,
    ``update_freq`` (= ``opts.accumulate_grad_batches`` when truthy, else 1).
    """

    def __init__(self, reg_lambda=1.0, mask=None, scaler=None, **kwargs):
        opts = kwargs.get("opts")
        accumulate = getattr(opts, "accumulate_grad_batches", None) if opts else None
        self.update_freq = accumulate if (opts and accumulate) else 1
        self.scaler = scaler
        self.mask = mask
        self.reg_lambda = reg_lambda
        self.task_id = 0

    # ---- hooks called by the trainer (train.py:181-213, vqa_cont_learner.py:211-254)
    def update(self, model, **kwargs):
        """Between tasks."""
        self.task_id += 1

    def update_after_new_task(self, **kwargs):
        """After the model for the new task is initialised; nothing to do by default."""

    def update_after_backward(self, **kwargs):
        """After ``loss.backward()``; nothing to do by default."""

    def update_after_step(self, **kwargs):
        """After an optimizer step; nothing to do by default."""

    def compute_loss(self, model, loss, **kwargs):
        raise NotImplementedError

    def replay(self, model, **kwargs):
        """Default: no memory, hence no replay loss and zero examples."""
        return None, 0

    def _is_batch_after_step(self, batch_idx=0):
        """True on the micro-batch that closes a gradient-accumulation window."""
        return (batch_idx + 1) % self.update_freq == 0


class Naive(CLStrategy):
    """Plain fine-tuning: the task loss is returned untouched (``base.py:50-57``)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def compute_loss(self, model, loss, **kwargs):
        return loss
