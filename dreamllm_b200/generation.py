"""Token sampling for `DreamLLMForCausalMLM.generate` (the kv-cache decode path the reference reaches through HF
`GenerationMixin.generate`, omni/eval/vqa/vqa_inference.py:112-130, modeling_dreamllm.py:1511-1547).

The per-token model work (prefill + cached decode) is the CUDA path in modeling_dreamllm.py; what lives here is the [B, V] logits
post-processing of one step — the semantics of transformers' `RepetitionPenaltyLogitsProcessor`, `TemperatureLogitsWarper`,
`TopKLogitsWarper`, `TopPLogitsWarper` in HF's order — as a handful of torch ops on the device holding the logits (plumbing, not a hot
path: 32 008 floats per sequence per token).  Greedy (`do_sample=False`) is a plain argmax: bit-exact token ids.
tests/test_generation_cpu.py pins `warp_logits` to the installed transformers processors.
"""
from __future__ import annotations

import torch


def warp_logits(logits: torch.Tensor, input_ids: torch.Tensor | None = None, temperature: float = 1.0, top_k: int = 0,
                top_p: float = 1.0, repetition_penalty: float = 1.0, min_tokens_to_keep: int = 1) -> torch.Tensor:
    """[B, V] fp32 scores -> filtered scores (-inf = removed), HF order: repetition penalty, temperature, top-k, top-p."""
    scores = logits.float()
    if repetition_penalty != 1.0 and input_ids is not None:
        picked = torch.gather(scores, 1, input_ids)
        picked = torch.where(picked < 0, picked * repetition_penalty, picked / repetition_penalty)
        scores = scores.scatter(1, input_ids, picked)
    if temperature != 1.0:
        if not temperature > 0:
            raise ValueError(f"`temperature` has to be a strictly positive float, but is {temperature}")
        scores = scores / temperature
    if top_k and top_k > 0:
        k = min(max(top_k, min_tokens_to_keep), scores.shape[-1])
        kth = torch.topk(scores, k)[0][..., -1, None]
        scores = scores.masked_fill(scores < kth, float("-inf"))
    if top_p is not None and top_p < 1.0:
        if not 0 <= top_p <= 1.0:
            raise ValueError(f"`top_p` has to be a float > 0 and < 1, but is {top_p}")
        sorted_logits, sorted_indices = torch.sort(scores, descending=False)
        cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        sorted_remove = cumulative_probs <= (1 - top_p)
        sorted_remove[..., -min_tokens_to_keep:] = False
        remove = sorted_remove.scatter(1, sorted_indices, sorted_remove)
        scores = scores.masked_fill(remove, float("-inf"))
    return scores


def pick_next_token(logits: torch.Tensor, input_ids: torch.Tensor | None = None, do_sample: bool = False, generator=None, **warp) -> torch.Tensor:
    """One decoding step over [B, V] logits -> [B] token ids."""
    if not do_sample:
        if warp.get("repetition_penalty", 1.0) != 1.0:
            logits = warp_logits(logits, input_ids, repetition_penalty=warp["repetition_penalty"])
        return logits.argmax(-1)
    probs = warp_logits(logits, input_ids, **warp).softmax(-1)
    return torch.multinomial(probs, 1, generator=generator).squeeze(1)


@torch.no_grad()
def generate(model, input_ids, images=None, max_new_tokens: int = 16, do_sample: bool = False, temperature: float = 1.0, top_k: int = 0,
             top_p: float = 1.0, repetition_penalty: float = 1.0, eos_token_id=None, pad_token_id=None, generator=None,
             attention_mask=None, stopping_criteria=None):
    """Prefill once with the kv-cache, then one cached decode step per token.  Finished rows keep emitting `pad_token_id` (HF semantics);
    stops when every row has produced an EOS or a `stopping_criteria` callable `(input_ids, scores) -> bool | BoolTensor[B]` fires for it
    (HF `StoppingCriteria` semantics; the reference's VQA eval passes a keyword criterion, omni/eval/vqa/vqa_inference.py:105-106).
    Returns [B, S + n_generated] ids (prompt included, like HF for decoder-only models)."""
    if attention_mask is not None and not bool(attention_mask.all()):
        raise NotImplementedError("padded prompt batches are not supported by the kv-cache path; generate prompts of different lengths one at a time")
    eos = None
    if eos_token_id is not None:
        eos = torch.as_tensor([eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id), device=input_ids.device)
        if pad_token_id is None:
            pad_token_id = int(eos[0])
    if stopping_criteria and pad_token_id is None:
        pad_token_id = 0
    warp = dict(temperature=temperature, top_k=top_k, top_p=top_p, repetition_penalty=repetition_penalty)
    seq = input_ids
    unfinished = torch.ones(input_ids.shape[0], dtype=torch.bool, device=input_ids.device)
    out = model(input_ids=input_ids, images=images, use_cache=True, last_token_logits_only=True)
    cache = out.past_key_values
    for i in range(max_new_tokens):
        nxt = pick_next_token(out.logits[:, -1], seq, do_sample=do_sample, generator=generator, **warp)
        if pad_token_id is not None:
            nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_token_id))       # finished rows keep emitting pad
        if eos is not None:
            unfinished = unfinished & ~torch.isin(nxt, eos)
        seq = torch.cat([seq, nxt[:, None]], 1)
        for crit in (stopping_criteria or ()):
            done = crit(seq, out.logits[:, -1])
            done = torch.as_tensor(done, device=seq.device, dtype=torch.bool)
            unfinished = unfinished & ~(done.expand_as(unfinished))
        if i + 1 == max_new_tokens or ((eos is not None or stopping_criteria) and not bool(unfinished.any())):
            break
        out = model(input_ids=nxt[:, None], past_key_values=cache, use_cache=True, last_token_logits_only=True)
    return seq
