"""Gradient checkpointing of the decoder stack (reference modeling_dreamllm.py:994-1003; `gradient_checkpointing=True` in the shipped
training configs): recomputing each layer's forward inside backward gives the same loss and gradients as keeping the activations, with a
smaller activation peak."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _run(ckpt, padded):
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    torch.manual_seed(0)
    cfg = DreamLLMConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2)
    m = DreamLLMForCausalMLM(cfg).to(device="cuda", dtype=BF).train()
    if ckpt:
        m.gradient_checkpointing_enable()
        assert m.is_gradient_checkpointing and m.model.gradient_checkpointing
    ids = torch.randint(1, 512, (2, 160), generator=torch.Generator().manual_seed(1)).cuda()
    mask = torch.ones_like(ids)
    labels = ids.clone()
    if padded:
        mask[1, 100:] = 0
        labels[1, 100:] = -100
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    out = m(input_ids=ids, attention_mask=mask, labels=labels)
    held = torch.cuda.memory_allocated() - base            # activations kept for backward
    out.loss.backward()
    torch.cuda.synchronize()
    return float(out.loss), {k: p.grad.detach().float().clone() for k, p in m.named_parameters()}, held


@pytest.mark.parametrize("padded", [False, True])
def test_checkpointed_backward_equals_plain_backward(padded):
    loss0, g0, held0 = _run(False, padded)
    loss1, g1, held1 = _run(True, padded)
    assert abs(loss0 - loss1) <= 1e-5 * abs(loss0)
    for k in g0:
        torch.testing.assert_close(g1[k], g0[k], rtol=1e-3, atol=1e-6, msg=k)     # same deterministic kernels, recomputed
    assert held1 < 0.6 * held0, (held0, held1)                                     # only layer inputs (+ the lm-head tail) are kept
