"""CPU oracle for MAFED's modality-aware feature-distillation loss (fwd + bwd).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is a CPU
restatement -- torch CPU ops, same op chain and same order of operations as the
reference -- of:

* ``mafed/methods/distillation.py:105-122``   ``FeatureDistillation.distill``
* ``mafed/methods/distillation.py:124-166``   ``feature_distillation``
* ``mafed/methods/distillation.py:226-257``   cosine / mse / cls token losses
* ``mafed/methods/distillation.py:61-64``     distillation-layer resolution
* ``mafed/methods/distillation_loss_weights.py:33-60,81-89``  layer list + coeffs
* ``mafed/methods/distillation_loss_weights.py:71-79,148-174`` modality weights

The arithmetic itself lives in third-party torch (``torch.nn.MSELoss``,
``torch.nn.CosineEmbeddingLoss``, ``Tensor.sum``, autograd, autocast), pinned
``torch==2.2.2`` by the reference (``pyproject.toml:38``); this container has
torch 2.11.  ``closed_form`` below is an independent float64 numpy restatement
of the published formulae used to cross-check the torch chain.

Pinning: the reference ships no tests / golden vectors.
The oracle is pinned against outputs of the *unmodified reference class*
imported from ``/root/reference`` with stubbed third-party modules
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

COS_EPS = 1e-12  # EPSILON in ATen's cosine_embedding_loss


@dataclass
class OracleConfig:
    modality_strategy: str = "equal"  # equal | balanced | adaptive
    layer_strategy: str = "single"  # single | equal | discounted | cumulative
    gamma: float = 0.8
    num_hidden_layers: int = 11
    distillation_layer: Optional[int] = -1
    distillation_coeff: float = 1.0
    loss: str = "mse"  # mse | cosine
    cls_distillation: bool = False
    num_vision_tokens: int = 256
    lang_coeff: Optional[Sequence[float]] = None  # adaptive strategy only
    extra: dict = field(default_factory=dict)


# --------------------------------------------------------------------------- host tables
def resolve_layer(distillation_layer: Optional[int], num_hidden_layers: int) -> Optional[int]:
    """distillation.py:61-64 -- out-of-range / None silently becomes None."""
    if distillation_layer is not None and 0 <= distillation_layer < num_hidden_layers:
        return distillation_layer
    return None


def layer_plan(cfg: OracleConfig) -> Tuple[List[int], Optional[torch.Tensor], str]:
    """distillation_loss_weights.py:33-60,81-84.

    Returns (layers, layer_coeffs or None, effective strategy).
    """
    layer = resolve_layer(cfg.distillation_layer, cfg.num_hidden_layers)
    strategy = cfg.layer_strategy
    if layer is None and strategy == "single":
        raise AssertionError("Invalid layer weighting strategy 'single'.")
    if layer is None and strategy == "cumulative":
        raise AssertionError("Invalid layer weighting strategy 'cumulative'.")
    n = layer if strategy == "cumulative" else cfg.num_hidden_layers
    if layer is not None and strategy != "cumulative":
        strategy = "single"
    if strategy == "single":
        return [layer], None, strategy
    if strategy == "equal":
        coeffs = torch.ones(n) / n
    else:  # discounted, cumulative (and anything else) share the gamma rule (:58-60)
        coeffs = torch.tensor([cfg.gamma**d for d in torch.arange(n, 0, -1)])
        coeffs = coeffs / coeffs.sum()
    return list(range(n)), coeffs, strategy


def build_masks(attention_mask: torch.Tensor, n_vis: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """distillation.py:134-144."""
    bsz, txt = attention_mask.shape
    dev = attention_mask.device  # the reference allocates on the CPU and copies with .to(device)
    lang = torch.zeros((bsz, txt + n_vis), dtype=attention_mask.dtype).to(dev)
    lang[:, n_vis:] = attention_mask
    img = torch.zeros((bsz, txt + n_vis), dtype=attention_mask.dtype).to(dev)
    img[:, :n_vis] = 1
    return lang, img


def modality_weights(cfg: OracleConfig, lang_mask, img_mask, layer: int):
    """distillation_loss_weights.py:71-79,148-174."""
    s = cfg.modality_strategy
    if s == "equal":
        n_t = lang_mask.sum()
        n_v = img_mask.sum()
        tot = n_t + n_v
        return n_t / tot, n_v / tot
    if s == "balanced":
        return 0.5, 1 - 0.5
    if s == "adaptive":
        lc = torch.as_tensor(cfg.lang_coeff, dtype=torch.float32).reshape(-1)
        lw = lc.item() if lc.shape[0] == 1 else lc[layer].item()
        return lw, 1 - lw
    raise NotImplementedError(s)


# --------------------------------------------------------------------------- token losses
def _mse_token_loss(h, p, mask):
    """distillation.py:237-249."""
    dim = h.shape[-1]
    h = h.reshape(-1, dim)
    p = p.reshape(-1, dim)
    mask = mask.reshape(-1)
    loss = torch.nn.MSELoss(reduction="none")(h, p).sum(-1) / dim
    loss = loss * mask
    return loss.sum() / mask.sum()


def _cosine_token_loss(h, p, mask):
    """distillation.py:226-235."""
    dim = h.shape[-1]
    h = h.reshape(-1, dim)
    p = p.reshape(-1, dim)
    mask = mask.reshape(-1)
    targets = torch.ones_like(mask)
    loss = torch.nn.CosineEmbeddingLoss(reduction="none")(h, p, targets)
    loss = loss * mask
    return loss.sum() / mask.sum()


def _cls_loss(h, p, loss_kind: str):
    """distillation.py:251-257 (only valid with cosine; mse raises TypeError)."""
    h0 = h[:, 0]
    p0 = p[:, 0]
    targets = torch.ones(h0.shape[0])
    if loss_kind == "cosine":
        fn = torch.nn.CosineEmbeddingLoss(reduction="none")
    else:
        fn = torch.nn.MSELoss(reduction="none")
    return fn(h0, p0, targets).mean()  # MSELoss(...) with 3 args -> TypeError, as in the reference


# --------------------------------------------------------------------------- the path
def distill(students, teachers, attention_mask, cfg: OracleConfig):
    """distillation.py:105-166: returns (total, {layer: layer_loss}, {layer: (text, vision)})."""
    layers, coeffs, strategy = layer_plan(cfg)
    token_loss = _cosine_token_loss if cfg.loss == "cosine" else _mse_token_loss
    total = 0.0
    per_layer: Dict[int, torch.Tensor] = {}
    per_mod: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
    for layer in layers:
        c = 1.0 if coeffs is None or strategy == "single" else coeffs[layer]
        h, p = students[layer], teachers[layer]
        if cfg.cls_distillation:
            ll = _cls_loss(h, p, cfg.loss)
        else:
            lang_mask, img_mask = build_masks(attention_mask, cfg.num_vision_tokens)
            lw, vw = modality_weights(cfg, lang_mask, img_mask, layer)
            tl = token_loss(h, p, lang_mask)
            vl = token_loss(h, p, img_mask)
            ll = (lw * tl) + (vw * vl)
            per_mod[layer] = (tl.detach(), vl.detach())
        per_layer[layer] = ll.detach()
        total = total + c * cfg.distillation_coeff * ll
    return total, per_layer, per_mod


def forward_backward(students, teachers, attention_mask, cfg: OracleConfig, grad_out: float = 1.0,
                     autocast_bf16: Optional[bool] = None):
    """Run the path on CPU with autograd.

    ``autocast_bf16`` mirrors ``distillation.py:90`` (``torch.autocast(..., bfloat16)``); by
    default it is enabled iff the inputs are not fp32.  Returns a dict with ``loss`` (0-dim fp32),
    ``layer_losses``, ``modality_losses`` and ``grads`` (list aligned with ``students``; ``None``
    for un-selected layers).
    """
    students = [s.detach().clone().requires_grad_(True) for s in students]
    teachers = [t.detach() for t in teachers]
    if autocast_bf16 is None:
        autocast_bf16 = students[0].dtype != torch.float32
    ctx = torch.autocast(students[0].device.type, dtype=torch.bfloat16) if autocast_bf16 else contextlib.nullcontext()
    with ctx:
        total, per_layer, per_mod = distill(students, teachers, attention_mask, cfg)
    (total * grad_out).backward()
    return {
        "loss": total.detach(),
        "layer_losses": per_layer,
        "modality_losses": per_mod,
        "grads": [s.grad for s in students],
    }


# --------------------------------------------------------------------------- independent check
def closed_form(students, teachers, attention_mask, cfg: OracleConfig, grad_out: float = 1.0):
    """Float64 numpy restatement of SURVEY.md section 3.3's formulae (no torch ops).

    loss_m(l) = sum_{n in m} w_n * f(h_n, p_n) / sum_n w_n, with f = ||h-p||^2 / D (mse) or
    1 - h.p / sqrt((|h|^2+eps)(|p|^2+eps)) (cosine); layer = w_t*loss_t + w_v*loss_v;
    total = sum_l c_l * coeff * layer_l.  Gradients are the analytic derivatives.
    """
    layers, coeffs, strategy = layer_plan(cfg)
    am = attention_mask.numpy().astype(np.float64)
    bsz, txt = am.shape
    nv = cfg.num_vision_tokens
    w_text = np.zeros((bsz, txt + nv))
    w_text[:, nv:] = am
    w_vis = np.zeros((bsz, txt + nv))
    w_vis[:, :nv] = 1.0
    n_t, n_v = w_text.sum(), w_vis.sum()
    total = 0.0
    grads: List[Optional[np.ndarray]] = [None] * len(students)
    layer_losses = {}
    for layer in layers:
        c = 1.0 if coeffs is None else float(coeffs[layer])
        h = students[layer].detach().to(torch.float64).numpy()
        p = teachers[layer].detach().to(torch.float64).numpy()
        dim = h.shape[-1]
        if cfg.cls_distillation:
            if cfg.loss != "cosine":
                raise TypeError("cls distillation only works with the cosine loss")
            wv0 = np.zeros_like(w_vis)
            wv0[:, 0] = 1.0
            mods = [(wv0, 1.0, float(bsz))]
        else:
            if cfg.modality_strategy == "equal":
                lw, vw = n_t / (n_t + n_v), n_v / (n_t + n_v)
            elif cfg.modality_strategy == "balanced":
                lw, vw = 0.5, 0.5
            elif cfg.modality_strategy == "adaptive":
                lc = np.asarray(cfg.lang_coeff, dtype=np.float32).reshape(-1)
                lw = float(lc[0] if lc.shape[0] == 1 else lc[layer])
                vw = 1.0 - lw
            else:
                raise NotImplementedError
            mods = [(w_text, lw, n_t), (w_vis, vw, n_v)]
        if cfg.loss == "mse":
            tok = ((h - p) ** 2).sum(-1) / dim
            dtok = 2.0 * (h - p) / dim
        else:
            dot = (h * p).sum(-1)
            a = (h * h).sum(-1) + COS_EPS
            b = (p * p).sum(-1) + COS_EPS
            den = np.sqrt(a * b)
            cos = dot / den
            tok = 1.0 - cos
            dtok = -(p / den[..., None] - (cos / a)[..., None] * h)
        ll = 0.0
        g = np.zeros_like(h)
        for w, mw, cnt in mods:
            ll = ll + mw * (w * tok).sum() / cnt
            g = g + (mw / cnt) * w[..., None] * dtok
        layer_losses[layer] = ll
        total = total + c * cfg.distillation_coeff * ll
        grads[layer] = grad_out * c * cfg.distillation_coeff * g
    return {"loss": total, "layer_losses": layer_losses, "grads": grads}


# --------------------------------------------------------------------------- synthetic inputs
def make_inputs(n_tuple: int, bsz: int, txt: int, dim: int, n_vis: int = 256, dtype=torch.float32,
                seed: int = 1234, teacher: str = "close", mask: str = "ragged"):
    """SURVEY.md section 8(d) synthetic hidden states + left-padded int64 attention mask."""
    g = torch.Generator().manual_seed(seed)
    T = n_vis + txt
    students, teachers = [], []
    for _ in range(n_tuple):
        s = torch.randn(bsz, T, dim, generator=g, dtype=torch.float32)
        if teacher == "close":
            t = s + 0.1 * torch.randn(bsz, T, dim, generator=g, dtype=torch.float32)
        else:
            t = torch.randn(bsz, T, dim, generator=g, dtype=torch.float32)
        students.append(s.to(dtype))
        teachers.append(t.to(dtype))
    am = torch.ones(bsz, txt, dtype=torch.int64)
    if mask == "ragged":
        for b in range(bsz):
            valid = 1 + (7 * b) % txt
            am[b, : txt - valid] = 0  # left padding: zeros first, valid tokens on the right
    return students, teachers, am


# --------------------------------------------------------------------------- adaptive importances
def adaptive_importances(model, batches, layers, n_vis):
    """distillation_loss_weights.py:91-146 restated: per batch, gradient of the LM loss w.r.t. each selected
    hidden state, per-token L2 norm (:131), masked sums per modality (:133-137); then per-token means and
    lang / (lang + image) (:142-144)."""
    model.eval()
    lang = torch.zeros(len(layers))
    image = torch.zeros(len(layers))
    n_lang = n_image = 0.0
    for batch in batches:
        model.zero_grad()
        out = model(**batch, compute_loss=True, output_hidden_states=True, allow_input_gradients=True,
                    return_dict=True)
        lang_mask, image_mask = build_masks(batch["attention_mask"], n_vis)
        for i, layer in enumerate(layers):
            grad = torch.autograd.grad(out.loss, out.hidden_states[layer], retain_graph=True)[0]
            norm = torch.linalg.norm(grad, dim=-1)
            lang[i] += (norm * lang_mask).sum()
            image[i] += (norm * image_mask).sum()
        n_lang += lang_mask.sum()
        n_image += image_mask.sum()
    lang = lang / n_lang
    image = image / n_image
    model.zero_grad()
    return lang / (lang + image)
