"""Build the C-ABI shared library in-tree with nvcc (sm_100a only)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libmafed_distill.so")
SOURCES = [os.path.join(PKG, "csrc", "distill_abi.cu")]
HEADERS = [os.path.join(PKG, "csrc", n) for n in
           ("distill_common.cuh", "distill_ldg.cuh", "distill_tma.cuh", "distill_epilogue.cuh",
            "distill_comm.cuh", "distill_host.cuh")] + \
          [os.path.join(ROOT, "include", "mafed_distill.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA path cannot be built (there is no CPU fallback)")
    return exe


STAMP_PATH = os.path.join(LIB_DIR, "build.stamp")


def _fingerprint():
    """Hash of every source / header and the compile flags (mtimes do not survive a copy to another box)."""
    import hashlib
    h = hashlib.sha1(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        if os.path.exists(f):
            with open(f, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != _fingerprint()


def build(force=False, verbose=False):
    """Compile libmafed_distill.so for sm_100a if missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
                               "-o", LIB_PATH] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(STAMP_PATH, "w") as f:
        f.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
