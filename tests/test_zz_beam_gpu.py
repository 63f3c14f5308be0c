"""Beam search over the CUDA kv-cache decode path (`generate(num_beams=k)`, the reference's default VQA decoding:
omni/eval/vqa/vqa_inference.py:111-119).  The search logic is pinned to transformers on CPU (tests/test_generation_cpu.py); here: the
cache repeat / re-order on device and the consistency of the returned hypothesis with a full re-forward.  (Written after this round's GPU
budget was spent — first hardware run is the round-end suite; sorts last on purpose.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _model():
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    torch.manual_seed(0)
    cfg = DreamLLMConfig(vocab_size=32008, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                         max_position_embeddings=512)
    m = DreamLLMForCausalMLM(cfg)
    with torch.no_grad():
        m.lm_head.weight.mul_(4.0)         # peaked distributions: beams separate by more than bf16 noise
    return m.to(device="cuda", dtype=BF).eval()


def test_beam_search_on_the_cuda_decode_path():
    m = _model()
    ids = torch.randint(3, 32000, (2, 17), device="cuda")
    from dreamllm_b200.generation import beam_search
    one = beam_search(m, ids, num_beams=1, max_new_tokens=5)
    assert one.shape == (2, 22) and torch.equal(one[:, :17], ids)
    with torch.no_grad():                                     # one beam = greedy: every token is an argmax of a full re-forward (up to bf16 ties)
        for t in range(17, 22):
            ref = m(input_ids=one[:, :t]).logits[:, -1]
            assert bool((ref.gather(-1, one[:, t, None]).squeeze(-1) >= ref.max(-1).values - 0.2).all()), t
    assert m.generate(ids, max_new_tokens=3, num_beams=2).shape == (2, 20)        # the public entry point routes num_beams > 1 here
    seqs, scores = beam_search(m, ids, num_beams=4, max_new_tokens=6, return_scores=True)
    assert seqs.shape == (2, 23) and torch.equal(seqs[:, :17], ids) and torch.isfinite(scores).all()
    with torch.no_grad():                                     # the reported score is the hypothesis' mean log-prob under a full re-forward
        logp = torch.log_softmax(m(input_ids=seqs).logits.float(), -1)
    tok_lp = logp[:, 16:22].gather(-1, seqs[:, 17:23, None]).squeeze(-1)
    torch.testing.assert_close(scores, tok_lp.sum(-1) / 6.0, rtol=2e-2, atol=0.15)        # cached-decode vs full-forward bf16 logits
    assert bool((scores >= -12.0).all())
