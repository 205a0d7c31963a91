"""Batch-sharded path on >= 2 GPUs (NCCL, one process per GPU) against the single-device oracle."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("path", ["peer", "nccl"])
def test_sharded_equals_full_batch(path):
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "multirank_check.py")]
    env = dict(os.environ, MAFED_B200_DIST=path)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_one_process_two_devices():
    """The same process drives cuda:0 then cuda:1 (per-device kernel attributes, stream and device guards)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import oracle_cfg
    from gpu_util import rel_err, run_product
    from oracle import distill_oracle as O
    st, te, am = O.make_inputs(4, 3, 6, 2048, n_vis=256, dtype=torch.bfloat16, seed=81)
    meta = dict(modality="balanced", layer_strategy="discounted", loss="mse", gamma=0.5, num_hidden_layers=3, layer=None,
                n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
    ref = O.forward_backward(st, te, am, oracle_cfg(meta))
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        out = run_product(meta, st, te, am, dev=dev)
        assert abs(float(out["loss"]) - float(ref["loss"])) / float(ref["loss"]) < 2e-3
        for l in range(3):
            assert rel_err(out["grads"][l].float(), ref["grads"][l].float()) < 2e-3
