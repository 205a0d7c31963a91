"""Selective hidden-state capture against HF GPTNeoXModel's own `output_hidden_states=True` tuple (CPU)."""
import pytest
import torch

from mafed_b200.capture import HiddenStateCapture, find_decoder


def _tiny_neox():
    from transformers import GPTNeoXConfig, GPTNeoXModel
    cfg = GPTNeoXConfig(hidden_size=32, num_hidden_layers=4, num_attention_heads=4, intermediate_size=64,
                        vocab_size=128, max_position_embeddings=64, hidden_dropout=0.0, attention_dropout=0.0)
    torch.manual_seed(0)
    return GPTNeoXModel(cfg).eval()


def test_captured_states_equal_hf_tuple_and_carry_gradients():
    model = _tiny_neox()
    embeds = torch.randn(2, 10, 32, requires_grad=True)
    full = model(inputs_embeds=embeds, output_hidden_states=True).hidden_states
    assert len(full) == 5
    with HiddenStateCapture(model, [0, 2, 4]) as cap:
        model(inputs_embeds=embeds, output_hidden_states=False)
    hs = cap.hidden_states
    assert len(hs) == 5 and sorted(cap.states) == [0, 2, 4]
    for i in (0, 2, 4):
        torch.testing.assert_close(hs[i], full[i], rtol=0, atol=0)
    with pytest.raises(KeyError):
        hs[1]
    # hooks are gone afterwards
    cap.states.clear()
    model(inputs_embeds=embeds)
    assert not cap.states
    # the captured tensor is the graph's own: a loss on it back-propagates into the model
    with HiddenStateCapture(model, [-1, 3]) as cap:
        model(inputs_embeds=embeds)
    assert sorted(cap.states) == [3, 4]
    cap.hidden_states[3].pow(2).mean().backward()
    assert model.layers[0].attention.query_key_value.weight.grad is not None
    assert model.layers[3].attention.query_key_value.weight.grad is None   # layer 3 comes after state 3


def test_teacher_capture_detaches_and_find_decoder_on_wrapper():
    model = _tiny_neox()

    class Wrapper(torch.nn.Module):  # mirrors VLCLIPGPTNeoXForCausalLM.gpt_neox (vl_pythia.py:204-237)
        def __init__(self, inner):
            super().__init__()
            self.gpt_neox = inner

        def forward(self, **kw):
            return self.gpt_neox(**kw)

    wrapped = Wrapper(model)
    layers, norm = find_decoder(wrapped)
    assert len(layers) == 4 and norm is model.final_layer_norm
    with torch.no_grad(), HiddenStateCapture(wrapped, [1], detach=True) as cap:
        wrapped(inputs_embeds=torch.randn(1, 6, 32))
    assert not cap.hidden_states[1].requires_grad
    with pytest.raises(IndexError):
        HiddenStateCapture(wrapped, [7])
