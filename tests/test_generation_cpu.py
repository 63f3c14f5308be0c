"""`warp_logits` / `pick_next_token` pinned to the installed transformers logits processors (the ones HF `generate` applies in the
reference's eval scripts, omni/eval/vqa/vqa_inference.py:112-130)."""
import pytest
import torch

from dreamllm_b200.generation import pick_next_token, warp_logits


def _hf(scores, ids, temperature, top_k, top_p, rep):
    from transformers.generation.logits_process import (RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper, TopKLogitsWarper,
                                                        TopPLogitsWarper)
    if rep != 1.0:
        scores = RepetitionPenaltyLogitsProcessor(rep)(ids, scores)
    if temperature != 1.0:
        scores = TemperatureLogitsWarper(temperature)(ids, scores)
    if top_k:
        scores = TopKLogitsWarper(top_k)(ids, scores)
    if top_p < 1.0:
        scores = TopPLogitsWarper(top_p)(ids, scores)
    return scores


@pytest.mark.parametrize("temperature,top_k,top_p,rep", [(1.0, 0, 1.0, 1.0), (0.7, 0, 1.0, 1.0), (1.0, 5, 1.0, 1.0), (1.0, 0, 0.9, 1.0),
                                                         (0.8, 50, 0.95, 1.1), (1.3, 3, 0.5, 1.0), (1.0, 1000, 0.01, 1.2)])
def test_warp_logits_equals_transformers_processors(temperature, top_k, top_p, rep):
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(4, 257, generator=g) * 3
    ids = torch.randint(0, 257, (4, 9), generator=g)
    got = warp_logits(scores.clone(), ids, temperature=temperature, top_k=top_k, top_p=top_p, repetition_penalty=rep)
    want = _hf(scores.clone(), ids, temperature, top_k, top_p, rep)
    assert torch.equal(torch.isinf(got), torch.isinf(want))
    keep = ~torch.isinf(want)
    torch.testing.assert_close(got[keep], want[keep], rtol=1e-6, atol=1e-6)


def test_pick_next_token():
    g = torch.Generator().manual_seed(1)
    scores = torch.randn(3, 100, generator=g)
    assert torch.equal(pick_next_token(scores), scores.argmax(-1))                                   # greedy = argmax, bit-exact
    assert torch.equal(pick_next_token(scores, do_sample=True, top_k=1, generator=g), scores.argmax(-1))
    top5 = scores.topk(5).indices
    for _ in range(20):
        t = pick_next_token(scores, do_sample=True, top_k=5, temperature=2.0, generator=g)
        assert bool((top5 == t[:, None]).any(-1).all())
    a = pick_next_token(scores, do_sample=True, generator=torch.Generator().manual_seed(7))
    b = pick_next_token(scores, do_sample=True, generator=torch.Generator().manual_seed(7))
    assert torch.equal(a, b)                                                                           # seeded sampling is reproducible
    with pytest.raises(ValueError):
        warp_logits(scores, temperature=0.0)


# ------------------------------------------------------------------------------------------------ decode loop on a scripted model
class _Scripted:
    """Stands in for DreamLLMForCausalMLM in `generation.generate`: next-token logits are a fixed function of (row, position), so the
    loop's bookkeeping (cache hand-over, EOS / pad, stopping criteria, lengths) can be checked on CPU."""

    def __init__(self, table):
        self.table = table                     # [B, T] token to emit at each decode step
        self.calls = []

    def __call__(self, input_ids=None, images=None, past_key_values=None, use_cache=None, last_token_logits_only=None):
        from types import SimpleNamespace
        if past_key_values is None:
            past_key_values = {"step": 0}                 # like KVCache: created at prefill, then advanced IN PLACE
        else:
            past_key_values["step"] += 1
        step = past_key_values["step"]
        self.calls.append((tuple(input_ids.shape), step > 0, images is not None))
        B = input_ids.shape[0]
        logits = torch.full((B, 1, 50), -10.0)
        logits[torch.arange(B), 0, self.table[:, step]] = 10.0
        return SimpleNamespace(logits=logits, past_key_values=past_key_values)


def test_generate_loop_eos_pad_and_cache_handover():
    from dreamllm_b200.generation import generate
    table = torch.tensor([[5, 6, 2, 7, 8, 9], [11, 12, 13, 14, 2, 15]])
    m = _Scripted(table)
    ids = torch.tensor([[1, 3, 4], [1, 3, 3]])
    out = generate(m, ids, images="img", max_new_tokens=6, eos_token_id=2, pad_token_id=0)
    assert out.tolist() == [[1, 3, 4, 5, 6, 2, 0, 0], [1, 3, 3, 11, 12, 13, 14, 2]]        # stops when the last row hits EOS
    assert m.calls[0] == ((2, 3), False, True)                                               # prefill: whole prompt + images, no cache
    assert all(c == ((2, 1), True, False) for c in m.calls[1:]) and len(m.calls) == 5        # then one token per call on the cache
    m = _Scripted(table)
    out = generate(m, ids, max_new_tokens=3)
    assert out.shape == (2, 6) and len(m.calls) == 3                                         # no model call after the last token
    m = _Scripted(table)
    out = generate(m, ids, max_new_tokens=6, eos_token_id=[2, 13])                           # several EOS ids; pad defaults to the first
    assert out.tolist() == [[1, 3, 4, 5, 6, 2], [1, 3, 3, 11, 12, 13]]


def test_generate_stopping_criteria():
    from dreamllm_b200.generation import generate
    table = torch.tensor([[5, 6, 7, 8, 9, 10]])
    seen = []

    def stop_on_7(input_ids, scores):
        seen.append(input_ids.shape[1])
        return bool((input_ids[0, -1] == 7))
    out = generate(_Scripted(table), torch.tensor([[1, 3]]), max_new_tokens=6, stopping_criteria=[stop_on_7])
    assert out.tolist() == [[1, 3, 5, 6, 7]] and seen == [3, 4, 5]
    per_row = lambda ids, scores: ids[:, -1] == 6                                            # BoolTensor[B] form
    out = generate(_Scripted(torch.tensor([[5, 6, 7, 8], [6, 9, 9, 6]])), torch.tensor([[1], [1]]), max_new_tokens=4,
                   stopping_criteria=[per_row])
    assert out.tolist() == [[1, 5, 6], [1, 6, 0]]


# ------------------------------------------------------------------------------------------------ beam search pinned to transformers
class _HFCache:
    """KVCache protocol (in-place advance, `repeat_interleave`, `reorder`) over a transformers DynamicCache."""

    def __init__(self, cache):
        self.cache = cache

    def repeat_interleave(self, k):
        self.cache.batch_repeat_interleave(k)

    def reorder(self, idx):
        self.cache.reorder_cache(idx)


class _HFAdapter:
    """A transformers causal LM behind the call signature `generation.generate` / `beam_search` use."""

    def __init__(self, hf):
        self.hf = hf

    def __call__(self, input_ids=None, images=None, past_key_values=None, use_cache=None, last_token_logits_only=None):
        from types import SimpleNamespace
        pkv = past_key_values.cache if past_key_values is not None else None
        out = self.hf(input_ids=input_ids, past_key_values=pkv, use_cache=True)
        cache = past_key_values if past_key_values is not None else _HFCache(out.past_key_values)
        return SimpleNamespace(logits=out.logits[:, -1:].float(), past_key_values=cache)


def _tiny_llama(seed, vocab=40):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(seed)
    cfg = LlamaConfig(vocab_size=vocab, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                      max_position_embeddings=64, eos_token_id=2, bos_token_id=1, pad_token_id=0)
    m = LlamaForCausalLM(cfg).eval()
    with torch.no_grad():
        m.lm_head.weight.mul_(8.0)            # peaked next-token distributions: hypotheses of different lengths and EOS picks occur
    return m


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("num_beams,length_penalty,early", [(3, 1.0, False), (5, 1.0, False), (4, 0.6, False), (3, 1.0, True)])
def test_beam_search_equals_transformers_generate(seed, num_beams, length_penalty, early):
    from dreamllm_b200.generation import beam_search
    hf = _tiny_llama(seed)
    g = torch.Generator().manual_seed(100 + seed)
    ids = torch.cat([torch.ones(2, 1, dtype=torch.long), torch.randint(3, 40, (2, 4), generator=g)], 1)
    with torch.no_grad():
        want = hf.generate(ids, num_beams=num_beams, do_sample=False, max_new_tokens=10, length_penalty=length_penalty, early_stopping=early,
                           eos_token_id=2, pad_token_id=39, return_dict_in_generate=True, output_scores=True)
    # (pad id != 0: the installed transformers fills with `pad_token_id or eos_token_id`, i.e. treats pad id 0 as unset)
    got, scores = beam_search(_HFAdapter(hf), ids, num_beams=num_beams, max_new_tokens=10, length_penalty=length_penalty,
                              early_stopping=early, eos_token_id=2, pad_token_id=39, return_scores=True)
    assert got.tolist() == want.sequences.tolist()
    torch.testing.assert_close(scores, want.sequences_scores, rtol=1e-4, atol=1e-4)


def test_beam_search_modes_and_kvcache_beam_ops():
    from dreamllm_b200.generation import beam_search
    from dreamllm_b200.modeling_dreamllm import DreamLLMForCausalMLM, KVCache
    hf = _tiny_llama(1)
    ids = torch.tensor([[1, 7, 9]])
    a = beam_search(_HFAdapter(hf), ids, num_beams=1, max_new_tokens=6, eos_token_id=2, pad_token_id=0)
    with torch.no_grad():
        greedy = hf.generate(ids, do_sample=False, max_new_tokens=6, eos_token_id=2, pad_token_id=0)
    assert a.tolist() == greedy.tolist()                                   # one beam = greedy
    t = beam_search(_HFAdapter(hf), ids, num_beams=3, max_new_tokens=6, eos_token_id=2, pad_token_id=0, length_normalization="total")
    assert t.shape[0] == 1 and t[0, :3].tolist() == [1, 7, 9]
    with pytest.raises(ValueError):
        beam_search(_HFAdapter(hf), ids, length_normalization="words")
    # KVCache beam ops (CPU tensors): repeat per beam, then re-order only the valid prefix
    c = KVCache(2, 2, 8, 1, 4, "cpu", dtype=torch.float32)
    c.len = 3
    for li in range(2):
        c.k[li][:, :3] = torch.arange(2.)[:, None, None, None] + 1        # row b holds b + 1
        c.v[li][:, :3] = -(torch.arange(2.)[:, None, None, None] + 1)
    c.repeat_interleave(3)
    assert c.k[0].shape[0] == 6 and c.k[0][:, 0, 0, 0].tolist() == [1, 1, 1, 2, 2, 2]
    c.reorder(torch.tensor([3, 3, 0, 5, 4, 1]))
    assert c.k[1][:, 2, 0, 0].tolist() == [2, 2, 1, 2, 2, 1] and c.v[1][:, 0, 0, 0].tolist() == [-2, -2, -1, -2, -2, -1]
    assert float(c.k[0][:, 3:].abs().sum()) == 0                           # rows past the valid length stay zero
    assert DreamLLMForCausalMLM._reorder_cache(c, torch.arange(6)) is c


def test_padded_prompt_batches_keep_the_hf_output_layout():
    """Padded prompt batches.  Left-padded (HF's convention for decoder-only models): ONE batched prefill + batched decode with the 2-D
    attention_mask and mask-derived position_ids handed to the model every step (modeling_dreamllm.py:1511-1547).  Right-padded: each row
    decoded on its own.  Either way the output is `[original padded prompt | new tokens | pad]` — `out[:, :S] == input_ids` and
    `out[:, S:]` holds only generated tokens, also when rows stop at different lengths (the old re-padding shifted prompts)."""
    from types import SimpleNamespace

    from dreamllm_b200.generation import generate

    class ByLength:          # row b emits (number of REAL prompt tokens + step) % 50; token 9 = EOS for rows whose real length is 2
        def __init__(self):
            self.calls = []

        def __call__(self, input_ids=None, images=None, past_key_values=None, use_cache=None, last_token_logits_only=None,
                     attention_mask=None, position_ids=None):
            B = input_ids.shape[0]
            self.calls.append((tuple(input_ids.shape), None if attention_mask is None else tuple(attention_mask.shape),
                               None if position_ids is None else position_ids[:, -1].tolist()))
            if past_key_values is None:
                real = attention_mask.sum(-1) if attention_mask is not None else torch.full((B,), input_ids.shape[1])
                past_key_values = {"n": real.clone(), "len0": real.clone()}
            else:
                past_key_values["n"] = past_key_values["n"] + 1
            logits = torch.full((B, 1, 50), -10.0)
            for b in range(B):
                tok = int(past_key_values["n"][b]) % 50
                if int(past_key_values["len0"][b]) == 2 and int(past_key_values["n"][b]) == 3:
                    tok = 9                                   # the short row stops after one token
                logits[b, 0, tok] = 10.0
            return SimpleNamespace(logits=logits, past_key_values=past_key_values)

    ids = torch.tensor([[0, 0, 7, 8], [5, 6, 7, 8]])
    mask = torch.tensor([[0, 0, 1, 1], [1, 1, 1, 1]])
    m = ByLength()
    out = generate(m, ids, attention_mask=mask, max_new_tokens=3, pad_token_id=0, eos_token_id=9)
    assert out[:, :4].tolist() == ids.tolist()                                   # the prompt block is untouched
    assert out[:, 4:].tolist() == [[2, 9, 0], [4, 5, 6]]                         # row 0: EOS after 2 tokens, then pad; row 1 runs on
    assert m.calls[0] == ((2, 4), (2, 4), [1, 3])                                # batched prefill, positions from the mask
    assert m.calls[1] == ((2, 1), (2, 5), [2, 4]) and m.calls[2] == ((2, 1), (2, 6), [3, 5])   # full mask + next position every step
    # right-padded: row-by-row fallback, same output layout
    ids_r = torch.tensor([[7, 8, 0, 0], [5, 6, 7, 8]])
    mask_r = torch.tensor([[1, 1, 0, 0], [1, 1, 1, 1]])
    m = ByLength()
    out = generate(m, ids_r, attention_mask=mask_r, max_new_tokens=3, pad_token_id=0, eos_token_id=9)
    assert out[:, :4].tolist() == ids_r.tolist()
    assert out[:, 4:].tolist() == [[2, 9, 0], [4, 5, 6]]
    assert all(c[0][0] == 1 for c in m.calls)                                    # batch 1 at a time
    with pytest.raises(NotImplementedError):
        generate(ByLength(), ids_r, images="x", attention_mask=mask_r, max_new_tokens=1)
