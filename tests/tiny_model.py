"""A tiny stand-in for VLCLIPGPTNeoXForCausalLM.forward (vl_pythia.py:247-326): same keyword interface,
returns `.loss` and a `hidden_states` tuple that is part of the loss graph."""
import torch
from torch import nn


class _Out:
    def __init__(self, loss, hidden_states):
        self.loss = loss
        self.hidden_states = hidden_states


class TinyModel(nn.Module):
    def __init__(self, dim, n_states):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(dim, dim) for _ in range(n_states)])
        self.head = nn.Linear(dim, 1)

    def forward(self, pixel_values=None, attention_mask=None, labels=None, input_ids=None, compute_loss=True,
                output_hidden_states=True, allow_input_gradients=False, return_dict=True, **kwargs):
        h = pixel_values
        states = []
        for i, layer in enumerate(self.layers):
            h = torch.tanh(layer(h)) + (0.5 * h if i else 0.0)
            states.append(h)
        # position-dependent readout so that text and image tokens receive different gradient norms
        weight = torch.linspace(0.2, 2.0, h.shape[1], device=h.device).view(1, -1, 1)
        loss = None
        if labels is not None:
            loss = ((self.head(h) * weight).squeeze(-1) - labels).pow(2).mean()
        return _Out(loss, tuple(states))


def make_batches(n, bsz, n_vis, txt, dim, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        am = torch.ones(bsz, txt, dtype=torch.int64)
        for b in range(bsz):
            am[b, : (b + i) % txt] = 0
        out.append({
            "pixel_values": torch.randn(bsz, n_vis + txt, dim, generator=g).to(device),
            "attention_mask": am.to(device),
            "labels": torch.randn(bsz, n_vis + txt, generator=g).to(device),
        })
    return out
