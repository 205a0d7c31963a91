"""Host-buffer entry points (pinned host memory in, host memory out) against the CPU oracle."""
import pytest
import torch

from golden_util import oracle_cfg
from gpu_util import make_method, rel_err
from oracle import distill_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["c-abi", "python-pipeline"])
@pytest.mark.parametrize("dtype,loss", [(torch.float32, "mse"), (torch.bfloat16, "mse"), (torch.bfloat16, "cosine")])
def test_host_step_matches_oracle(kind, dtype, loss):
    from mafed_b200.host_step import CHostStep, HostStep
    st, te, am = O.make_inputs(4, 3, 7, 768, n_vis=256, dtype=dtype, seed=61, mask="ragged")
    meta = dict(modality="equal", layer_strategy="discounted", loss=loss, gamma=0.5, num_hidden_layers=3, layer=None,
                n_vis=256, coeff=1.0, cls=False, lang_coeff=None)
    ref = O.forward_backward(st, te, am, oracle_cfg(meta))
    fd = make_method(meta)
    cls = CHostStep if kind == "c-abi" else HostStep
    hs = cls(fd, st[:3], te[:3], am, torch.device("cuda", 0))
    for _ in range(2):                                    # twice: buffers and events are reused correctly
        loss_v = hs.step()
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert float(loss_v) == pytest.approx(float(ref["loss"]), rel=tol)
    for l in range(3):
        assert rel_err(hs.h_g[l].float(), ref["grads"][l].float()) < tol
    if kind == "c-abi":
        for l in range(3):                                # out[1 + l]: per-layer losses, as the device path
            assert float(hs.h_out[1 + l]) == pytest.approx(float(ref["layer_losses"][l]), rel=tol)
        assert hs.lib.mafed_host_step_device_bytes(hs.handle and __import__("ctypes").byref(hs.shape)) > 0
        hs.close()
