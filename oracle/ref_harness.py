"""Import and drive the UNMODIFIED reference (``FeatureDistillation``, ``VLPythiaVQACLearner``) on this machine.

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the golden-vector generators, the parity tests
and ``bench.py``'s reference arm / ``cpu_baseline`` leg use it; no product module imports it.

The reference's files come from ``oracle/_ref/`` (byte-identical copies placed there by ``oracle/make_ref.py``;
they travel to the GPU box) or, in the build container, straight from ``/root/reference``.  Third-party packages
that are absent (``pytorch_lightning``, ``toolz``) and reference sub-packages outside the hot path
(``mafed.data``, ``mafed.model``'s architecture table, ``mafed.optim``, ``mafed.utils.eval_utils``) are replaced
by ``sys.modules`` stubs before the import; ``wandb.log`` is replaced by a recorder.  The reference's own code is
never modified.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types
from typing import Dict, List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(HERE, "_ref")
_SOURCE_TREE = os.environ.get("MAFED_REFERENCE", "/root/reference")

_loaded = None
wandb_log: List[dict] = []          # every dict the reference handed to wandb.log since the last clear()


def reference_root(prefer_source_tree: bool = False) -> Optional[str]:
    """Directory to put on sys.path: oracle/_ref (travels) or the reference's own tree (build container only)."""
    have_ref = os.path.exists(os.path.join(_REF_DIR, "mafed", "methods", "distillation.py"))
    have_tree = os.path.exists(os.path.join(_SOURCE_TREE, "mafed", "methods", "distillation.py"))
    if prefer_source_tree and have_tree:
        return _SOURCE_TREE
    if have_ref:
        return _REF_DIR
    return _SOURCE_TREE if have_tree else None


def available() -> bool:
    return reference_root() is not None


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _importable(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


class _IdentityLoader:
    """Stand-in for ``mafed.data.PrefetchLoader`` (a CUDA side-stream prefetcher, ``data/loader.py:40-80``)."""

    def __init__(self, loader):
        self.loader = loader

    def __iter__(self):
        return iter(self.loader)

    def __len__(self):
        return len(self.loader)

    def __getattr__(self, name):
        return getattr(self.loader, name)


def _default_collate(items):
    from torch.utils.data import default_collate
    return default_collate(items)


def install_stubs():
    """sys.modules stubs for what the reference imports but the path does not need."""
    if not _importable("toolz"):
        _module("toolz")
        _module("toolz.sandbox", unzip=lambda s: zip(*s))
    if not _importable("pytorch_lightning"):
        class WandbLogger:  # placeholder base class (utils/logger.py:5)
            def __init__(self, *a, **k):
                pass

        class LightningModule:  # the hooks VLPythiaVQACLearner inherits; `log` records instead of logging
            def __init__(self, *a, **k):
                pass

            def log(self, name, value, **kwargs):
                self.__dict__.setdefault("logged", []).append((name, value, kwargs))

            def on_before_optimizer_step(self, optimizer):
                pass

            def on_validation_epoch_end(self):
                pass

        _module("pytorch_lightning", LightningModule=LightningModule)
        _module("pytorch_lightning.loggers", WandbLogger=WandbLogger)
        _module("pytorch_lightning.utilities")
        _module("pytorch_lightning.utilities.rank_zero", rank_zero_only=lambda fn: fn,
                rank_zero_warn=lambda *a, **k: None)
    # reference sub-packages outside the hot path: data pipeline, model zoo, optimizers, VQA metric
    collate = {"train": {"vlpythia": _default_collate}, "valid": {"vlpythia": _default_collate}}
    _module("mafed.data", PrefetchLoader=_IdentityLoader, collate_fn=collate)
    _module("mafed.model", model_architecture={})
    _module("mafed.optim")
    _module("mafed.optim.adamw", AdamW=None)
    _module("mafed.optim.sched", get_linear_schedule_with_warmup=None)
    _module("mafed.utils.eval_utils", VQAGenerativeAccuracy=object)


def load(prefer_source_tree: bool = False):
    """Import the unmodified ``mafed.methods`` (and keep ``wandb.log`` recording).  Returns the module."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root(prefer_source_tree)
    if root is None:
        raise RuntimeError("the reference is not available: neither oracle/_ref (python oracle/make_ref.py in the "
                           "build container) nor /root/reference exists")
    for name in [n for n in sys.modules if n == "mafed" or n.startswith("mafed.")]:
        del sys.modules[name]
    sys.path.insert(0, root)
    try:
        importlib.import_module("mafed")           # the real (empty) package first, so that stubs become its children
        install_stubs()
        methods = importlib.import_module("mafed.methods")
    finally:
        sys.path.remove(root)
    import wandb
    wandb.log = lambda d, *a, **k: wandb_log.append(dict(d))
    methods.__reference_root__ = root
    _loaded = methods
    return methods


def learner_class():
    """The unmodified ``VLPythiaVQACLearner`` (``mafed/model/vqa_cont_learner.py``)."""
    load()
    import importlib.util
    root = _loaded.__reference_root__
    path = os.path.join(root, "mafed", "model", "vqa_cont_learner.py")
    spec = importlib.util.spec_from_file_location("mafed.model.vqa_cont_learner", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mafed.model.vqa_cont_learner"] = mod
    spec.loader.exec_module(mod)
    return mod.VLPythiaVQACLearner


class Opts:
    """The five attributes ``FeatureDistillation`` / ``CLStrategy`` read from ``opts`` (distillation.py:37-46)."""
    tasks = ["a", "b", "c"]
    batch_size = 4
    seed = 42
    pin_mem = False
    accumulate_grad_batches = 1


class Output:
    def __init__(self, hidden_states, loss=None):
        self.hidden_states = hidden_states
        self.loss = loss


def make_reference_method(modality="balanced", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=11,
                          layer=None, coeff=1.0, cls=False, n_vis=256, lang_coeff=None, opts=None, **extra):
    """``CLMethod["featdistill"](...)`` of the unmodified reference, constructed like ``mafed/train.py:119-134``."""
    import torch
    methods = load()
    fd = methods.CLMethod["featdistill"](
        memory_size=8, opts=opts or Opts(), model_type="vlpythia",
        distillation_modality_weighing_strategy=modality, distillation_layer_weighing_strategy=layer_strategy,
        distillation_coeff=coeff, distillation_layer=layer, cls_distillation=cls, distillation_loss=loss, gamma=gamma,
        num_hidden_layers=num_hidden_layers, **extra)
    fd.num_vision_tokens = n_vis            # public attribute (distillation.py:73)
    if modality == "adaptive" and lang_coeff is not None:
        fd.loss_weights.lang_coeff = torch.tensor(lang_coeff, dtype=torch.float32)
    return fd


def reference_forward_backward(fd, students, teachers, attention_mask, grad_out: float = 1.0, autocast_bf16: bool = False,
                               need_grads: bool = True) -> Dict[str, object]:
    """``fd.distill(output, batch)`` + ``backward()`` of the unmodified reference on the tensors' own device.
    Returns loss, the per-layer values it sent to W&B, the gradients and the batch it mutated."""
    import torch
    st = [s.detach().clone().requires_grad_(need_grads) for s in students]
    fd.past_model = lambda **kw: Output(tuple(teachers))
    batch = {"attention_mask": attention_mask.clone(), "labels": torch.zeros(1)}
    wandb_log.clear()
    device = students[0].device.type
    ctx = torch.autocast(device, dtype=torch.bfloat16) if autocast_bf16 else contextlib.nullcontext()
    with ctx:
        loss = fd.distill(Output(tuple(st)), batch)
    if need_grads:
        (loss * grad_out).backward()
    logged = {}
    for d in wandb_log:
        logged.update(d)
    return {"loss": loss.detach(), "logged": logged, "grads": [s.grad for s in st], "batch": batch}
