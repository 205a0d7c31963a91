"""Golden vector for the adaptive modality importances (distillation_loss_weights.py:91-146), produced by
the UNMODIFIED reference `DistillationWeights.compute_adaptive_weights` driven with a tiny stand-in model.

    python tests/golden/make_golden_adaptive.py      (build container only: reads /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))


def main():
    from make_golden import _install_stubs
    _install_stubs()
    from mafed.methods.distillation_loss_weights import DistillationWeights
    from tiny_model import TinyModel, make_batches

    torch.manual_seed(7)
    n_vis, txt, D, L = 8, 5, 16, 3
    model = TinyModel(D, L + 1)
    batches = make_batches(2, 3, n_vis, txt, D, seed=11)
    dw = DistillationWeights("adaptive", "equal", num_hidden_layers=L, distillation_layer=None, num_vision_tokens=n_vis)
    ref_batches = [dict(b) for b in batches]
    imp = dw.compute_adaptive_weights(model, ref_batches)
    # running average over tasks (:62-69)
    dw.update_weights(model, [dict(b) for b in batches], 0)
    first = dw.lang_coeff.clone()
    dw.update_weights(model, [dict(b) for b in batches[:1]], 1)
    blob = {"importances": imp.detach().numpy(), "after_task0": first.detach().numpy(),
            "after_task1": dw.lang_coeff.detach().numpy()}
    for k, v in model.state_dict().items():
        blob["w_" + k] = v.numpy()
    for i, b in enumerate(batches):
        blob[f"b{i}_pixel_values"] = b["pixel_values"].numpy()
        blob[f"b{i}_attention_mask"] = b["attention_mask"].numpy()
        blob[f"b{i}_labels"] = b["labels"].numpy()
    np.savez_compressed(os.path.join(HERE, "adaptive_case.npz"), **blob)
    print("importances", imp, "task0", first, "task1", dw.lang_coeff)


if __name__ == "__main__":
    main()
