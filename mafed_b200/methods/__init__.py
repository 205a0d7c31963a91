"""Mirror of ``mafed/methods/__init__.py:6-11``: the strategy registry.

Only the strategies on the B200 hot path live here (``featdistill``) plus the trivial ``naive``
baseline; ``ewc`` and ``replay`` are other continual-learning methods that SURVEY.md section 8 puts
out of scope -- in a drop-in installation they keep coming from the reference.
"""
from mafed_b200.methods.base import CLStrategy, Naive
from mafed_b200.methods.distillation import FeatureDistillation
from mafed_b200.methods.distillation_loss_weights import DistillationWeights

CLMethod = {
    "naive": Naive,
    "featdistill": FeatureDistillation,
}

__all__ = ["CLMethod", "CLStrategy", "Naive", "FeatureDistillation", "DistillationWeights"]
