"""Whole text-only DreamLLMForCausalMLM step (embedding -> L layers -> norm -> lm_head -> shifted CE, fwd+bwd)
against the CPU oracle (oracle.decoder_oracle.causal_lm, reference modeling_dreamllm.py:846-1043, :1353-1509)."""
import pytest
import torch

from oracle import decoder_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _run(hidden, inter, heads, layers, vocab, B, S, pad):
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    cfg = DreamLLMConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                         num_attention_heads=heads)
    torch.manual_seed(3)
    model = DreamLLMForCausalMLM(cfg).to(device="cuda", dtype=BF)
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(0, vocab, (B, S), generator=g)
    labels = ids.clone()
    am = None
    if pad:
        am = torch.ones(B, S, dtype=torch.long)
        am[-1, S - pad:] = 0
        labels[-1, S - pad:] = -100
    labels[0, :5] = -100
    out = model(input_ids=ids.cuda(), labels=labels.cuda(), attention_mask=am.cuda() if am is not None else None)
    out.loss.backward()
    torch.cuda.synchronize()

    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}

    def oracle(dtype):
        emb = sd["model.embed_tokens.weight"].to(dtype).requires_grad_(True)
        lay = [{k: sd[f"model.layers.{i}.{k}"].to(dtype).requires_grad_(True) for k in O.LAYER_KEYS} for i in range(layers)]
        nw = sd["model.norm.weight"].to(dtype).requires_grad_(True)
        lw = sd["lm_head.weight"].to(dtype).requires_grad_(True)
        loss, _, _ = O.causal_lm(ids, labels, emb, lay, nw, lw, heads, attention_mask=am)
        loss.backward()
        return loss, emb, lay, nw, lw

    l32, e32, lay32, n32, w32 = oracle(torch.float32)
    lbf, ebf, laybf, nbf, wbf = oracle(BF)
    return model, out, (l32, e32, lay32, n32, w32), (lbf, ebf, laybf, nbf, wbf)


@pytest.mark.parametrize("pad", [0, 21])
def test_causal_lm_step_vs_oracle(pad):
    layers = 2
    model, out, o32, obf = _run(256, 512, 2, layers, 1000, 2, 160, pad)
    l32, lbf = float(o32[0]), float(obf[0])
    got = float(out.loss)
    assert abs(got - l32) <= 1.5 * abs(lbf - l32) + 2e-3 * abs(l32), (got, l32, lbf)

    def err(a, ref):
        return float((a.float() - ref.float()).abs().mean())

    named = dict(model.named_parameters())
    pairs = [("model.embed_tokens.weight", o32[1], obf[1]), ("model.norm.weight", o32[3], obf[3]), ("lm_head.weight", o32[4], obf[4])]
    for i in range(layers):
        for k in O.LAYER_KEYS:
            pairs.append((f"model.layers.{i}.{k}", o32[2][i][k], obf[2][i][k]))
    for name, p32, pbf in pairs:
        ours = named[name].grad.cpu()
        scale = float(p32.grad.abs().mean()) + 1e-12
        e_o, e_r = err(ours, p32.grad), err(pbf.grad, p32.grad)
        assert e_o <= 1.3 * e_r + 5e-3 * scale, f"{name}: ours {e_o:.3e} ref-bf16 {e_r:.3e} scale {scale:.3e}"


def test_token_index_paths_bit_exact():
    """embedding rows are exact copies; shifted-label bookkeeping: changing an ignored label must not change the loss."""
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    cfg = DreamLLMConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=2)
    torch.manual_seed(0)
    m = DreamLLMForCausalMLM(cfg).to(device="cuda", dtype=BF)
    ids = torch.randint(0, 512, (2, 64), device="cuda")
    lab = ids.clone()
    lab[:, :10] = -100
    a = m(input_ids=ids, labels=lab).loss
    ids2 = ids.clone()
    lab2 = lab.clone()
    lab2[:, 0] = 7                      # position 0 is never a target after the shift
    b = m(input_ids=ids2, labels=lab2).loss
    assert torch.equal(a, b)
    # argmax path: logits (labels=None) are deterministic and their argmax matches a second run bit-for-bit
    l1 = m(input_ids=ids).logits
    l2 = m(input_ids=ids).logits
    assert torch.equal(l1.argmax(-1), l2.argmax(-1)) and torch.equal(l1, l2)
