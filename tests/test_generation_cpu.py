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
