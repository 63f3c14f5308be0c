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


def test_argmax_and_greedy_ids_vs_bf16_oracle():
    """Token-id paths against the ORACLE (the r1 tests compared the CUDA model with itself).  The bf16 oracle is
    bit-identical to the reference's bf16 path (tests/test_oracle_pin.py), so wherever the top-1 margin of the fp32 logits exceeds twice
    the bf16 error budget the argmax — and with it every greedily decoded id — must agree exactly:
      * budget e = 2 x max|logits(ref bf16) - logits(ref fp32)|  (the like-for-like criterion of DESIGN.md §2, measured here);
      * our logits must stay inside it, the margin rule must cover most positions (otherwise the test is vacuous), and on those positions
        argmax(ours) == argmax(ref bf16) == argmax(ref fp32);
      * greedy decoding through the kv-cache path reproduces the oracle's greedy ids (full re-forward on the CPU each step) for as long as
        every decoded position is a clear-margin one."""
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    hidden, inter, heads, layers, vocab, B, S = 256, 512, 2, 2, 512, 2, 96
    cfg = DreamLLMConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=heads)
    torch.manual_seed(7)
    model = DreamLLMForCausalMLM(cfg)
    with torch.no_grad():                                  # N(0, 0.02) init gives near-uniform logits; widen the head so margins are real
        model.lm_head.weight.mul_(8.0)
    model = model.to(device="cuda", dtype=BF).eval()
    ids = torch.randint(0, vocab, (B, S), generator=torch.Generator().manual_seed(8))
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}

    def oracle_logits(x_ids, dtype):
        emb = sd["model.embed_tokens.weight"].to(dtype)
        lay = [{k: sd[f"model.layers.{i}.{k}"].to(dtype) for k in O.LAYER_KEYS} for i in range(layers)]
        with torch.no_grad():
            return O.causal_lm(x_ids, None, emb, lay, sd["model.norm.weight"].to(dtype), sd["lm_head.weight"].to(dtype), heads)[1]

    l32, lbf = oracle_logits(ids, torch.float32), oracle_logits(ids, BF)
    with torch.no_grad():
        ours = model(input_ids=ids.cuda()).logits.float().cpu()
    budget = 2.0 * float((lbf - l32).abs().max())
    assert float((ours - l32).abs().max()) <= budget, (float((ours - l32).abs().max()), budget)
    top2 = l32.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 2 * budget
    assert float(clear.float().mean()) > 0.5, f"only {float(clear.float().mean()):.2f} of the positions have a clear top-1 margin"
    assert torch.equal(ours.argmax(-1)[clear], l32.argmax(-1)[clear])
    assert torch.equal(ours.argmax(-1)[clear], lbf.argmax(-1)[clear])

    # greedy decode (kv-cache path) vs the oracle's greedy ids
    prompt = ids[:, :24]
    n_new = 12
    got = model.generate(prompt.cuda(), max_new_tokens=n_new, do_sample=False).cpu()
    assert torch.equal(got[:, :24], prompt)
    cur = prompt.clone()
    compared = 0
    alive = torch.ones(B, dtype=torch.bool)
    for step in range(n_new):
        lg32, lgbf = oracle_logits(cur, torch.float32)[:, -1], oracle_logits(cur, BF)[:, -1]
        t2 = lg32.topk(2, dim=-1).values
        alive &= (t2[:, 0] - t2[:, 1]) > 2 * budget                      # once a near-tie is decoded the continuations may differ
        nxt = lgbf.argmax(-1)
        for b in range(B):
            if alive[b]:
                assert int(got[b, 24 + step]) == int(nxt[b]), (b, step, int(got[b, 24 + step]), int(nxt[b]))
                compared += 1
        cur = torch.cat([cur, got[:, 24 + step: 25 + step]], dim=1)         # follow OUR sequence so later steps stay comparable
    assert compared >= 3, compared
