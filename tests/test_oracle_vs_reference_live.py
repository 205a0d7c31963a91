"""The oracle, and the host tables of the product's mirror, against the UNMODIFIED reference run live on this machine
(``oracle/_ref`` or ``/root/reference``; not gpu): random configurations beyond the committed golden vectors --
strategies x loss x shapes x masks x upstream gradient x fp32 / bf16-autocast.  Skipped where the reference files are
absent (``python oracle/make_ref.py`` places them in the build container; they travel to the GPU box)."""
import random

import pytest
import torch

from oracle import distill_oracle as O
from oracle import ref_harness as R

pytestmark = pytest.mark.skipif(not R.available(), reason="the reference is not available on this machine")


def _random_case(seed):
    rng = random.Random(seed)
    nh = rng.choice([1, 2, 3, 5, 7])
    layer_strategy = rng.choice(["single", "equal", "discounted", "cumulative"])
    layer = None
    if layer_strategy == "single":
        layer = rng.randrange(nh)
    elif layer_strategy == "cumulative":
        layer = rng.randrange(1, nh) if nh > 1 else None
        if layer is None:
            layer_strategy = "equal"
    modality = rng.choice(["equal", "balanced", "adaptive"])
    loss = rng.choice(["mse", "cosine"])
    n_vis = rng.choice([1, 4, 16])
    cfg = dict(modality=modality, layer_strategy=layer_strategy, loss=loss, gamma=rng.choice([0.3, 0.5, 0.8, 0.9]),
               num_hidden_layers=nh, layer=layer, coeff=rng.choice([1.0, 0.5, 2.0]), n_vis=n_vis,
               lang_coeff=[rng.random() for _ in range(nh)] if modality == "adaptive" else None)
    shape = dict(n_tuple=nh + 2, bsz=rng.choice([1, 2, 5]), txt=rng.choice([1, 3, 8]), dim=rng.choice([8, 24, 40]),
                 teacher=rng.choice(["close", "independent"]), mask=rng.choice(["ragged", "ones"]))
    return cfg, shape, rng.choice([1.0, 0.25, 3.0]), rng.random() < 0.3


def _oracle_cfg(c):
    return O.OracleConfig(modality_strategy=c["modality"], layer_strategy=c["layer_strategy"], gamma=c["gamma"],
                          num_hidden_layers=c["num_hidden_layers"], distillation_layer=c["layer"],
                          distillation_coeff=c["coeff"], loss=c["loss"], num_vision_tokens=c["n_vis"],
                          lang_coeff=c["lang_coeff"])


@pytest.mark.parametrize("seed", range(48))
def test_oracle_equals_the_reference_on_random_configurations(seed):
    c, shape, grad_out, bf16 = _random_case(seed)
    dtype = torch.bfloat16 if bf16 else torch.float32
    st, te, am = O.make_inputs(shape["n_tuple"], shape["bsz"], shape["txt"], shape["dim"], n_vis=c["n_vis"], dtype=dtype,
                               seed=1000 + seed, teacher=shape["teacher"], mask=shape["mask"])
    fd = R.make_reference_method(**c)
    ref = R.reference_forward_backward(fd, st, te, am, grad_out=grad_out, autocast_bf16=bf16)
    got = O.forward_backward(st, te, am, _oracle_cfg(c), grad_out=grad_out, autocast_bf16=bf16)
    # the same torch ops in the same order: equal to rounding of the last place
    assert float(got["loss"]) == pytest.approx(float(ref["loss"]), rel=1e-6, abs=1e-12)
    for l, (g, r) in enumerate(zip(got["grads"], ref["grads"])):
        assert (g is None) == (r is None), l
        if g is not None:
            assert torch.allclose(g.float(), r.float(), rtol=1e-5, atol=1e-9), l
    # the float64 closed form (an independent derivation) agrees to fp32 / bf16 accuracy
    cf = O.closed_form(st, te, am, _oracle_cfg(c), grad_out=grad_out)
    assert float(cf["loss"]) == pytest.approx(float(ref["loss"]), rel=2e-5 if not bf16 else 2e-3)


@pytest.mark.parametrize("seed", range(48))
def test_mirror_host_tables_equal_the_reference(seed):
    """Layer list, layer coefficients and the fixed modality weights of ``mafed_b200.methods.DistillationWeights``
    against the reference class of the same name (``distillation_loss_weights.py:10-89,148-174``)."""
    from mafed_b200.methods import DistillationWeights
    c, _, _, _ = _random_case(seed)
    ref = R.make_reference_method(**c).loss_weights
    mine = DistillationWeights(distillation_modality_weighing_strategy=c["modality"],
                               distillation_layer_weighing_strategy=c["layer_strategy"], gamma=c["gamma"],
                               num_hidden_layers=c["num_hidden_layers"], distillation_layer=c["layer"],
                               num_vision_tokens=c["n_vis"])
    layers = list(ref.get_distillation_layers())
    assert list(mine.get_distillation_layers()) == layers
    for l in layers:
        assert float(mine.get_layer_loss_weight(l)) == pytest.approx(float(ref.get_layer_loss_weight(l)), rel=1e-7)
