"""Tiny VL-Pythia look-alike (vision tokens + HF GPTNeoXModel decoder) with the keyword interface of
VLCLIPGPTNeoXForCausalLM.forward (vl_pythia.py:247-326), for end-to-end replay() tests."""
import torch
from torch import nn


class _Out:
    def __init__(self, loss, logits, hidden_states):
        self.loss, self.logits, self.hidden_states = loss, logits, hidden_states


class TinyVL(nn.Module):
    def __init__(self, n_vis=8, dim=64, layers=3, vocab=97, patch=12, fp32_island=False):
        """``fp32_island``: run the whole forward with autocast disabled, so that the model computes the same fp32
        numbers on the CPU (where the reference's ``torch.autocast("cuda", ...)`` region is inert) and on the GPU
        -- what the reference-generated replay goldens need."""
        super().__init__()
        self.fp32_island = fp32_island
        from transformers import GPTNeoXConfig, GPTNeoXModel
        cfg = GPTNeoXConfig(hidden_size=dim, num_hidden_layers=layers, num_attention_heads=4, intermediate_size=2 * dim,
                            vocab_size=vocab, max_position_embeddings=128, hidden_dropout=0.0, attention_dropout=0.0)
        self.gpt_neox = GPTNeoXModel(cfg)
        self.connector = nn.Linear(patch, dim)
        self.embed_out = nn.Linear(dim, vocab, bias=False)
        self.n_vis = n_vis

    def forward(self, input_ids=None, pixel_values=None, attention_mask=None, labels=None, compute_loss=False,
                output_hidden_states=False, allow_input_gradients=False, return_dict=True, **kwargs):
        if self.fp32_island and torch.is_autocast_enabled(pixel_values.device.type):
            with torch.autocast(pixel_values.device.type, enabled=False):
                return self._forward(input_ids, pixel_values, attention_mask, labels, output_hidden_states)
        return self._forward(input_ids, pixel_values, attention_mask, labels, output_hidden_states)

    def _forward(self, input_ids, pixel_values, attention_mask, labels, output_hidden_states):
        vision = self.connector(pixel_values)                         # [B, n_vis, D]
        text = self.gpt_neox.embed_in(input_ids)                      # [B, txt, D]
        embeds = torch.cat([vision, text], dim=1)
        mask = torch.cat([torch.ones_like(attention_mask[:, :1]).expand(-1, self.n_vis), attention_mask], dim=1)
        out = self.gpt_neox(inputs_embeds=embeds, attention_mask=mask, output_hidden_states=output_hidden_states)
        logits = self.embed_out(out.last_hidden_state)
        loss = None
        if labels is not None:
            tgt = torch.cat([torch.full_like(labels[:, :1], -100).expand(-1, self.n_vis), labels], dim=1)
            loss = nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(),
                                               tgt[:, 1:].reshape(-1), ignore_index=-100)
        return _Out(loss, logits, out.hidden_states if output_hidden_states else None)


def make_batch(bsz=4, txt=6, n_vis=8, patch=12, vocab=97, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    am = torch.ones(bsz, txt, dtype=torch.int64)
    for b in range(bsz):
        am[b, : b % txt] = 0                                         # left padding
    ids = torch.randint(0, vocab, (bsz, txt), generator=g)
    labels = torch.where(am.bool(), ids, torch.full_like(ids, -100))
    return {"input_ids": ids.to(device), "pixel_values": torch.randn(bsz, n_vis, patch, generator=g).to(device),
            "attention_mask": am.to(device), "labels": labels.to(device)}
