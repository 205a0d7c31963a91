"""mafed_b200 -- B200-native (sm_100a) implementation of MAFED's modality-aware feature
distillation hot path, behind the reference's ``mafed.methods`` strategy API.

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/mafed_distill.h``); ``cabi`` the ctypes
binding; ``distill_op`` the autograd operator; ``methods/`` the mirror of ``mafed/methods``.
"""
__version__ = "0.1.0"
