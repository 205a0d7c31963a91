"""The C-ABI library builds for sm_100a, loads, and exports every symbol the header declares (not gpu)."""
import ctypes
import os
import re
import subprocess

from mafed_b200 import build, cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mafed_distill.h")).read()
    return sorted(set(re.findall(r"\b(mafed_(?:distill|comm|host)_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = cabi.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(cabi.EXPORTS) == names
    assert lib.mafed_distill_abi_version() == cabi.ABI_VERSION
    assert lib.mafed_distill_sums_len(15) == 32 and lib.mafed_distill_out_len(15) == 46
    assert lib.mafed_distill_ws_bytes(15) >= 148 * 15 * 2 * 4
    assert b"invalid argument" in lib.mafed_distill_error_string(-1)


def test_argument_errors_without_a_gpu():
    lib = cabi.load()
    bad = cabi.make_shape(0, 1, 1, 0, 8, cabi.F32, cabi.LOSS_MSE)
    assert lib.mafed_distill_fwd(ctypes.byref(bad), None, None, None, None, None) == -1
    bad = cabi.make_shape(1, 1, 4, 2, 8, 7, cabi.LOSS_MSE)
    assert lib.mafed_distill_fwd(ctypes.byref(bad), None, None, None, None, None) == -2
    ok = cabi.make_shape(1, 1, 4, 2, 8, cabi.F32, cabi.LOSS_MSE)
    assert lib.mafed_distill_fwd(ctypes.byref(ok), None, None, None, None, None) == -1  # null pointer tables
    assert lib.mafed_distill_set_variant(9) == -1
    # the one-call step: weights and the output vector are mandatory, the two masks come as a pair
    w = cabi.make_weights(cabi.MODW_EQUAL, 1.0, [1.0])
    one = ctypes.c_float(1.0)
    assert lib.mafed_distill_step(ctypes.byref(ok), None, None, None, None, None, one, None, None, None, None, None,
                                  None, None, None) == -1
    assert lib.mafed_distill_step(ctypes.byref(ok), None, None, None, None, ctypes.byref(w), one, None, 8, None, None,
                                  8, None, None, None) == -1
    assert lib.mafed_distill_fwd_step(ctypes.byref(ok), None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.mafed_comm_trace(None, None) == -1


def test_sass_is_sm100_and_uses_bulk_copies():
    out = subprocess.run(["cuobjdump", "-sass", cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "sm_100" in out
    assert "UBLKCP" in out  # cp.async.bulk (TMA engine) in the staged kernels
    assert re.search(r"LDG\.E(\.NA)?\.128", out) and "STG.E.128" in out and "LDS.128" in out


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CUDA library -> the product raises; there is no eager / CPU fallback to fall into."""
    import pytest
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", str(tmp_path / "libmafed_distill.so"))
    with pytest.raises(cabi.MafedDistillError, match="no CPU / eager fallback"):
        cabi.load()


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/mafed_distill.h compiles as C99 (-pedantic) and a C program linked against the library can call
    the ABI -- the non-Python binding route of INTEGRATION.md."""
    build.build()
    src = os.path.join(ROOT, "tests", "c", "cabi_consumer.c")
    exe = str(tmp_path / "cabi_consumer")
    lib_dir = os.path.dirname(cabi.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", lib_dir, "-lmafed_distill", f"-Wl,-rpath,{lib_dir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip() == "OK", run.stdout + run.stderr
