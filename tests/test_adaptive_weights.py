"""Adaptive modality importances: oracle vs reference golden (CPU) and the fused
token-norm kernel vs both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import distill_oracle as O
from tiny_model import TinyModel, make_batches

HERE = os.path.dirname(os.path.abspath(__file__))
N_VIS, TXT, D, L = 8, 5, 16, 3


def _load():
    z = np.load(os.path.join(HERE, "golden", "adaptive_case.npz"))
    model = TinyModel(D, L + 1)
    model.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_")})
    batches = []
    i = 0
    while f"b{i}_pixel_values" in z.files:
        batches.append({k: torch.from_numpy(z[f"b{i}_{k}"]) for k in ("pixel_values", "attention_mask", "labels")})
        i += 1
    return z, model, batches


def test_oracle_importances_match_reference():
    z, model, batches = _load()
    regenerated = make_batches(2, 3, N_VIS, TXT, D, seed=11)
    assert all(torch.equal(a["pixel_values"], b["pixel_values"]) for a, b in zip(batches, regenerated))
    imp = O.adaptive_importances(model, [dict(b) for b in batches], list(range(L)), N_VIS)
    np.testing.assert_allclose(imp.numpy(), z["importances"], rtol=1e-6)


def _torch_importances(model, batches):
    """distillation_loss_weights.py:91-146 in plain torch ops on the batches' device."""
    lang = image = None
    n_lang = n_image = 0.0
    for batch in batches:
        batch = dict(batch)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(**batch, compute_loss=True, output_hidden_states=True, allow_input_gradients=True)
            am = batch["attention_mask"]
            B, txt = am.shape
            lm = torch.zeros(B, N_VIS + txt, dtype=am.dtype, device=am.device); lm[:, N_VIS:] = am
            im = torch.zeros_like(lm); im[:, :N_VIS] = 1
            la, ia = [], []
            for l in range(L):
                g = torch.autograd.grad(out.loss, out.hidden_states[l], retain_graph=True)[0]
                n = torch.linalg.norm(g, dim=-1)
                la.append((n * lm).sum()); ia.append((n * im).sum())
            la, ia = torch.stack(la), torch.stack(ia)
            lang = la if lang is None else lang + la
            image = ia if image is None else image + ia
            n_lang = n_lang + lm.sum(); n_image = n_image + im.sum()
    lang = lang / n_lang; image = image / n_image
    return lang / (lang + image)


@pytest.mark.gpu
def test_compute_adaptive_weights_on_gpu_matches_reference():
    from mafed_b200.methods import DistillationWeights
    z, model, batches = _load()
    model = model.cuda()
    cuda_batches = [{k: v.cuda() for k, v in b.items()} for b in batches]
    dw = DistillationWeights("adaptive", "equal", num_hidden_layers=L, distillation_layer=None, num_vision_tokens=N_VIS)
    imp = dw.compute_adaptive_weights(model, [dict(b) for b in cuda_batches])
    # (a) the reference's algorithm in plain torch on the same device under the same autocast: tight
    want = _torch_importances(model, cuda_batches)
    np.testing.assert_allclose(imp.cpu().numpy(), want.cpu().numpy(), rtol=1e-5)
    # (b) the CPU golden from the unmodified reference ran the model in fp32 (autocast("cuda") is inert on
    # CPU tensors); on the GPU the model's matmuls run in bf16, hence the bf16 tolerance
    np.testing.assert_allclose(imp.cpu().numpy(), z["importances"], rtol=2e-3)
    b0 = dict(cuda_batches[0])
    dw.compute_adaptive_weights(model, [b0])
    assert "lang_masks" in b0 and "image_masks" in b0          # side effect kept (:115,121)
    # running average over tasks (:62-69) and the host table the kernels read
    dw.update_weights(model, [dict(b) for b in cuda_batches], 0)
    np.testing.assert_allclose(dw.lang_coeff.cpu().numpy(), z["after_task0"], rtol=2e-3)
    dw.update_weights(model, [dict(b) for b in cuda_batches[:1]], 1)
    np.testing.assert_allclose(dw.lang_coeff.cpu().numpy(), z["after_task1"], rtol=2e-3)
    assert dw.kernel_tables()[2] == pytest.approx([float(x) for x in dw.lang_coeff.cpu()], rel=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2], ids=["ldg", "tma"])
@pytest.mark.parametrize("dtype,dim", [(torch.float32, 768), (torch.bfloat16, 2048), (torch.float32, 2048),
                                       (torch.bfloat16, 100), (torch.float16, 50)])
def test_token_norm_sums_kernel(variant, dtype, dim):
    from mafed_b200 import cabi
    from mafed_b200.distill_op import token_norm_sums
    with cabi.tuning(variant=variant):
        g = torch.Generator(device="cuda").manual_seed(5)
        B, txt, n_layers = 5, 9, 4
        grads = [torch.randn(B, 256 + txt, dim, generator=g, device="cuda").to(dtype) * (0.1 + l) for l in range(n_layers)]
        am = torch.ones(B, txt, dtype=torch.int64, device="cuda")
        for b in range(B):
            am[b, : (3 * b) % txt] = 0
        sums = token_norm_sums(grads, am, 256).cpu()
        lang_mask, image_mask = O.build_masks(am.cpu(), 256)
        tol = 1e-5 if dtype == torch.float32 else 2e-3
        for l, gr in enumerate(grads):
            norm = torch.linalg.norm(gr.float().cpu().double(), dim=-1)
            assert float(sums[2 * l]) == pytest.approx(float((norm * lang_mask).sum()), rel=tol)
            assert float(sums[2 * l + 1]) == pytest.approx(float((norm * image_mask).sum()), rel=tol)
        assert float(sums[-2]) == float(am.sum()) and float(sums[-1]) == B * 256
