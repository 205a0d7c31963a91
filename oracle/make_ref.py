"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, as a runnable package.

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference is pure Python, so "building" it
means placing byte-identical copies of the few files the path needs where the GPU box can import them
(``/root/reference`` does not exist there).  The copies go ONLY into ``oracle/_ref/`` -- git-ignored, never part
of the repository's history, shipped to the GPU box with the snapshot like the built ``.so`` files -- together with
``MANIFEST.json`` (source path + sha256 of every file, so a reader can check that nothing was edited).

    python oracle/make_ref.py            # in the build container; ``__graft_entry__.build()`` runs it too

What is taken (``/root/reference/mafed/...``):
  methods/__init__.py, base.py, distillation.py, distillation_loss_weights.py, ewc.py, replay.py
        the ``CLMethod`` registry and ``FeatureDistillation``
  model/vqa_cont_learner.py
        the caller, ``VLPythiaVQACLearner.training_step`` (row a12)
  utils/logger.py
        ``LOGGER`` (imported by distillation_loss_weights.py)
Everything those files import from third parties that is absent here / there (``pytorch_lightning``, ``toolz``) or
out of scope (``mafed.data``, ``mafed.model``, ``mafed.optim``, ``mafed.utils.eval_utils``) is stubbed at import
time by ``oracle/ref_harness.py``; the copied files themselves are never touched.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("MAFED_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

FILES = [
    "mafed/__init__.py",
    "mafed/methods/__init__.py",
    "mafed/methods/base.py",
    "mafed/methods/distillation.py",
    "mafed/methods/distillation_loss_weights.py",
    "mafed/methods/ewc.py",
    "mafed/methods/replay.py",
    "mafed/model/vqa_cont_learner.py",
    "mafed/utils/logger.py",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(verbose=True):
    """Copy FILES from the reference into oracle/_ref/ (idempotent).  Returns the manifest, or None if the
    reference is not present (the GPU box: it uses the files that travelled with the snapshot)."""
    if not os.path.isdir(os.path.join(REFERENCE, "mafed")):
        return None
    manifest = {"reference": REFERENCE, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        assert _sha(src) == _sha(dst)
        manifest["files"][rel] = _sha(dst)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} unmodified reference files from {REFERENCE}")
    return manifest


def verify():
    """True if every file of oracle/_ref/ still has the sha256 recorded when it was copied."""
    path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.exists(path):
        return False
    with open(path) as f:
        manifest = json.load(f)
    return all(os.path.exists(os.path.join(OUT, rel)) and _sha(os.path.join(OUT, rel)) == digest
               for rel, digest in manifest["files"].items())


if __name__ == "__main__":
    sys.exit(0 if make() is not None else 1)
