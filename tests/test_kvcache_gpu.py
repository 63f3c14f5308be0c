"""kv-cache inference path: prefill + decode == full re-forward; two-pass dream-query prompt embedding == single
pass over the concatenated sequence (reference get_prompt_embeds, modeling_dreamllm.py:1598-1673); text -> latents pipeline runs."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _model(layers=2):
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    torch.manual_seed(0)
    cfg = DreamLLMConfig(vocab_size=32008, hidden_size=256, intermediate_size=512, num_hidden_layers=layers, num_attention_heads=2,
                         max_position_embeddings=512)
    return DreamLLMForCausalMLM(cfg)


def test_incremental_decode_matches_full_forward():
    m = _model().to(device="cuda", dtype=BF).eval()
    ids = torch.randint(0, 32000, (2, 70), device="cuda")
    with torch.no_grad():
        full = m(input_ids=ids).logits                       # [B, S, V]
        out = m(input_ids=ids[:, :50], use_cache=True)       # prefill 50
        cache = out.past_key_values
        torch.testing.assert_close(out.logits, full[:, :50], rtol=2e-2, atol=2e-2)
        steps = [m(input_ids=ids[:, 50:64], past_key_values=cache, use_cache=True).logits]     # continuation of 14 tokens
        for t in range(64, 70):                                                                # then 1 token at a time
            steps.append(m(input_ids=ids[:, t:t + 1], past_key_values=cache, use_cache=True).logits)
    inc = torch.cat(steps, 1)
    assert cache.len == 70
    torch.testing.assert_close(inc, full[:, 50:], rtol=2e-2, atol=3e-2)
    assert float((inc.argmax(-1) == full[:, 50:].argmax(-1)).float().mean()) > 0.97


def test_greedy_generate_token_ids_match_recompute():
    m = _model().to(device="cuda", dtype=BF).eval()
    ids = torch.randint(0, 32000, (2, 33), device="cuda")
    gen = m.generate_greedy(ids, max_new_tokens=6)
    assert gen.shape == (2, 39) and torch.equal(gen[:, :33], ids)
    with torch.no_grad():
        for t in range(33, 39):        # every generated token is the argmax of a full re-forward over its prefix
            ref = m(input_ids=gen[:, :t]).logits[:, -1]
            top2 = ref.topk(2).values
            sure = (top2[:, 0] - top2[:, 1]) > 0.05         # ignore near-ties (bf16)
            assert bool(((ref.argmax(-1) == gen[:, t]) | ~sure).all())


def test_prompt_embeds_two_pass_equals_single_pass_and_pipeline_runs():
    from dreamllm_b200.modeling_plugins import DreamEmbedding, StableDiffusionHead
    m = _model()
    Q = 8
    dream = DreamEmbedding(num_dream_queries=Q, embed_hidden_size=256)
    m.stable_diffusion_head = StableDiffusionHead(dict(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4),
                                                       cross_attention_dim=128, vae=dict(block_out_channels=(64, 128, 128, 128))),
                                                  embed_hidden_size=256)
    m.model.attach_plugins(None, dream, image_start_id=32003, dream_start_id=32006)
    m = m.to(device="cuda", dtype=BF).eval()
    ids = torch.randint(3, 32000, (2, 21), device="cuda")
    pe = m.get_prompt_embeds(ids)
    assert pe.shape == (2, Q, 256)
    # single pass: [text, <dream_start>, Q x <im_patch>, <dream_end>] with the dream-query splice (training layout)
    full_ids = torch.cat([ids, torch.tensor([[32006] + [32002] * Q + [32007]] * 2, device="cuda")], 1)
    with torch.no_grad():
        out = m.model(input_ids=full_ids, images_dm=torch.zeros(2, 1), output_hidden_states=True)
    want = out.hidden_states[-1][:, 22:22 + Q]
    torch.testing.assert_close(pe.float(), want.float(), rtol=3e-2, atol=3e-2)
    neg = torch.randint(3, 32000, (2, 5), device="cuda")
    lat = m.stable_diffusion_pipeline(ids, neg, guidance_scale=3.0, num_inference_steps=3, height=128, width=128, scheduler="ddim")
    assert lat.shape == (2, 4, 16, 16) and torch.isfinite(lat).all()
    img = m.stable_diffusion_pipeline(ids, neg, guidance_scale=3.0, num_inference_steps=2, height=128, width=128, scheduler="ddim",
                                      output_type="pt")
    assert img.shape == (2, 3, 128, 128) and float(img.min()) >= 0.0 and float(img.max()) <= 1.0


def test_sampling_generate_and_eos_padding():
    """`generate`: top_k=1 sampling is greedy; sampled tokens come from the top-k of a full re-forward; rows that hit EOS emit pad."""
    m = _model().to(device="cuda", dtype=BF).eval()
    ids = torch.randint(0, 32000, (2, 20), device="cuda")
    greedy = m.generate(ids, max_new_tokens=5)
    assert torch.equal(greedy, m.generate_greedy(ids, max_new_tokens=5))
    g = torch.Generator(device="cuda").manual_seed(0)
    assert torch.equal(m.generate(ids, max_new_tokens=5, do_sample=True, top_k=1, generator=g), greedy)
    samp = m.generate(ids, max_new_tokens=4, do_sample=True, top_k=8, temperature=1.5, generator=g)
    assert samp.shape == (2, 24) and torch.equal(samp[:, :20], ids)
    with torch.no_grad():
        for t in range(20, 24):
            ref = m(input_ids=samp[:, :t]).logits[:, -1]
            top = ref.topk(12).indices                       # top-8 of the cached step, with slack for bf16 near-ties
            assert bool((top == samp[:, t, None]).any(-1).all())
    eos = int(greedy[0, 20])                                 # row 0's first generated token -> that row finishes at once
    out = m.generate(ids, max_new_tokens=5, eos_token_id=eos, pad_token_id=0)
    assert int(out[0, 20]) == eos and bool((out[0, 21:] == 0).all())
    if eos not in greedy[1, 20:].tolist():
        assert torch.equal(out[1], greedy[1])                # the other row is unaffected
    # right-padded batch: every row is decoded on its own valid tokens; HF output layout [padded prompt | new tokens | pad]
    mask = torch.tensor([[1] * 20, [1] * 19 + [0]], device="cuda")
    rag = m.generate(ids, attention_mask=mask, max_new_tokens=3, pad_token_id=0)
    assert rag.shape == (2, 23) and torch.equal(rag[:, :20], ids)
    assert torch.equal(rag[0], m.generate(ids[:1], max_new_tokens=3)[0])
    assert torch.equal(rag[1, 20:], m.generate(ids[1:, :19], max_new_tokens=3)[0, 19:])


def test_cached_layer_vs_oracle_left_padded_batch():
    """kv-cache decode against the ORACLE (the cache tests compared the CUDA path with itself).  Scenario of
    oracle.decoder_oracle.cached_decode_scenario — a LEFT-padded 2-row batch through one decoder layer, prefill + 3 single-token steps,
    full 2-D mask + mask-derived positions every call — whose oracle outputs are pinned to the reference's own `past_key_value` path
    (tests/golden/kvcache_layer.npz, tests/test_oracle_pin.py).  Like-for-like: err(ours bf16, ref fp32) <= 1.5 x err(ref bf16, ref fp32)."""
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMDecoderLayer, KVCache
    from oracle import decoder_oracle as O
    hidden, inter, heads = 256, 512, 2
    p, calls = O.cached_decode_scenario(hidden, inter, heads)
    ref32 = O.run_cached_scenario(p, calls, heads, torch.float32)
    refbf = O.run_cached_scenario(p, calls, heads, BF)
    layer = DreamLLMDecoderLayer(DreamLLMConfig(hidden_size=hidden, intermediate_size=inter, num_attention_heads=heads, num_hidden_layers=1))
    sd = {k: v.clone() for k, v in p.items()}
    sd["self_attn.rotary_emb.inv_freq"] = layer.self_attn.rotary_emb.inv_freq.clone()
    layer.load_state_dict(sd)
    layer = layer.to(device="cuda", dtype=BF).eval()
    cache = KVCache(1, 2, 128, heads, hidden // heads, "cuda")
    for i, (x, am, pos) in enumerate(calls):
        cache.set_mask(am.cuda())
        with torch.no_grad():
            y = layer(x.cuda().to(BF), position_ids=pos.cuda(), past_key_value=(cache, 0), use_cache=True)[0]
        cache.len += x.shape[1]
        valid = am[:, -x.shape[1]:].bool()
        ours, r32, rbf = y.cpu().float()[valid], ref32[i][valid], refbf[i].float()[valid]
        e_o, e_r = float((ours - r32).abs().mean()), float((rbf - r32).abs().mean())
        assert e_o <= 1.5 * e_r + 1e-3 * float(r32.abs().mean()), (i, e_o, e_r)
        if i == 0:                       # pad QUERY rows of the prefill: zeros from the attention core, as flash-attn's pad_input (:545)
            assert bool(torch.isfinite(y).all())


def test_left_padded_batch_generate_matches_unpadded_rows():
    """Batched greedy / beam decode of a LEFT-padded prompt batch (what vqa_inference.py:112-130 does) == decoding each prompt alone."""
    m = _model().to(device="cuda", dtype=BF).eval()
    g = torch.Generator().manual_seed(3)
    a, b = torch.randint(3, 32000, (1, 30), generator=g), torch.randint(3, 32000, (1, 19), generator=g)
    ids = torch.zeros(2, 30, dtype=torch.long)
    ids[0], ids[1, 11:] = a[0], b[0]
    mask = torch.ones(2, 30, dtype=torch.long)
    mask[1, :11] = 0
    out = m.generate(ids.cuda(), attention_mask=mask.cuda(), max_new_tokens=6, pad_token_id=0)
    assert out.shape == (2, 36) and torch.equal(out[:, :30].cpu(), ids)
    alone = [m.generate(a.cuda(), max_new_tokens=6), m.generate(b.cuda(), max_new_tokens=6)]
    with torch.no_grad():                # compare where the single-prompt decode itself is not a bf16 near-tie
        for row, (prompt, solo) in enumerate(zip((a, b), alone)):
            for t in range(6):
                ref = m(input_ids=solo[:, :prompt.shape[1] + t]).logits[0, -1].float()
                top2 = ref.topk(2).values
                if float(top2[0] - top2[1]) <= 0.05:
                    break
                assert int(out[row, 30 + t]) == int(solo[0, prompt.shape[1] + t]), (row, t)
    beams = m.generate(ids.cuda(), attention_mask=mask.cuda(), max_new_tokens=4, num_beams=3, pad_token_id=0)
    assert beams.shape[0] == 2 and torch.equal(beams[:, :30].cpu(), ids)
