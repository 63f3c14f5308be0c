"""Pins the splice / conditioning-gather / loss-combination restatements to the LIVE reference model wrappers: `DreamLLMModel.forward`
(modeling_dreamllm.py:1045-1158) and `DreamLLMForCausalMLM.forward` (:1353-1509) are exec'd verbatim from /root/reference and run on CPU
with stand-in sub-modules (an nn.Embedding, callables for the plugins, a recorder for `_forward` / `stable_diffusion_head`).

Chain of custody this closes:  live reference == oracle/splice_oracle.py == SplicePlan index maps (tests/test_collator_cpu.py, CPU)
== CUDA copy_rows / segment_sum_rows / gather_rows kernels (tests/test_clip_splice_gpu.py, GPU).
Build container only: skipped where /root/reference does not exist (GPU box)."""
import math
import os
import textwrap
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch.nn import CrossEntropyLoss

from oracle import decoder_oracle as O
from oracle import splice_oracle as SO

REF = "/root/reference/omni/models/dreamllm/modeling_dreamllm.py"
pytestmark = pytest.mark.skipif(not os.path.isfile(REF), reason="reference checkout not present (GPU box)")
TOK = {"<im_start>": 90, "<im_patch>": 91, "<im_end>": 92, "<dream_start>": 93, "<dream_end>": 94}
ST = {"additional_special_tokens": TOK, "<s>": 1, "</s>": 2}
P, Q, H, V = 5, 3, 16, 96


class _Out(tuple):
    """BaseModelOutputWithPast stand-in: indexable + the attributes the reference reads."""
    def __new__(cls, hidden):
        o = super().__new__(cls, (hidden,))
        o.past_key_values = o.hidden_states = o.attentions = None
        o.additional_log_info = {}
        return o


def _method(start_marker, end_marker, nth=0):
    src = open(REF).read()
    a = -1
    for _ in range(nth + 1):
        a = src.index(start_marker, a + 1)
    b = src.index(end_marker, a)
    ns = dict(torch=torch, F=F, math=math, np=np, CrossEntropyLoss=CrossEntropyLoss, BaseModelOutputWithPast=None,
              CausalLMOutputWithPast=lambda **kw: SimpleNamespace(**kw), DEFAULT_IMAGE_START_TOKEN="<im_start>",
              DEFAULT_DREAM_START_TOKEN="<dream_start>", DEFAULT_DREAM_END_TOKEN="<dream_end>", DEFAULT_IMAGE_PATCH_TOKEN="<im_patch>",
              DEFAULT_BOS_TOKEN="<s>", DEFAULT_EOS_TOKEN="</s>",
              logger=SimpleNamespace(warning=lambda *a, **k: None, warning_once=lambda *a, **k: None, error=lambda *a, **k: None))
    exec("from __future__ import annotations\n" + textwrap.dedent(src[a:b]), ns)
    return ns["forward"]


def _ids(g):
    rows = [[1, 90] + [91] * P + [92, 7, 8, 93] + [91] * Q + [94, 2],
            [1, 5, 90] + [91] * P + [92, 90] + [91] * P + [92, 6, 2],          # two <im_start>, only one image left for this row
            [1, 93] + [91] * Q + [94, 9, 93] + [91] * Q + [94, 2]]             # two dreams
    S = max(len(r) for r in rows)
    return torch.tensor([r + [0] * (S - len(r)) for r in rows])


@pytest.mark.parametrize("n_images,with_dream", [(2, True), (3, True), (1, False), (0, True)])
def test_splice_oracle_equals_live_reference_model_forward(n_images, with_dream):
    fwd = _method("    def forward(\n        self,\n        input_ids: torch.LongTensor = None,\n        images:", "    # `DreamEmbedding`")
    g = torch.Generator().manual_seed(0)
    ids = _ids(g)
    emb = torch.nn.Embedding(V, H)
    dq = torch.randn(1, Q, H, generator=g)
    feats = torch.randn(max(n_images, 1), P, H, generator=g)
    seen = {}

    class M:
        training = False
        config = SimpleNamespace(special_tokens2ids_dict=ST)
        embed_tokens = emb
        dream_embedding = staticmethod(lambda bs=1: dq.repeat(bs, 1, 1))
        clip_vision_embedding = staticmethod(lambda images: feats[:n_images] if images is not None else torch.zeros(()))

        def _forward(self, **kw):
            seen.update(kw)
            return kw["inputs_embeds"]
    M.dream_embedding.embed_len = Q
    M.forward = fwd
    images = torch.zeros(n_images, 3, 2, 2) if n_images else None
    images_dm = torch.zeros(2, 3, 2, 2) if with_dream else None
    with torch.no_grad():
        got_ref = M().forward(input_ids=ids, images=images, images_dm=images_dm)
        want = SO.splice(ids, emb(ids), feats[:n_images] if n_images else None, dq if with_dream else None, 90, 93)
    assert torch.equal(got_ref, want)
    assert seen["input_ids"] is None                                   # the reference hands `_forward` embeddings only


@pytest.mark.parametrize("drop_prob,n_dm", [(None, 3), (0.1, 2)])
def test_conditioning_gather_null_prompt_and_loss_equal_live_reference_causal_lm_forward(drop_prob, n_dm):
    fwd = _method("    def forward(\n        self,\n        input_ids: torch.LongTensor = None,\n        images:",
                  "    def prepare_inputs_for_generation", nth=1)
    g = torch.Generator().manual_seed(1)
    ids = _ids(g)
    B, S = ids.shape
    hidden = torch.randn(B, S, H, generator=g)
    u_hidden = torch.randn(1, Q + 4, H, generator=g)
    labels = ids.clone()
    labels[ids >= 90] = -100
    labels[ids == 0] = -100
    head_w = torch.randn(V, H, generator=g) * 0.1
    calls, sd_calls = [], []

    def model(**kw):
        calls.append(kw)
        return _Out(hidden if len(calls) == 1 else u_hidden)
    model.config = SimpleNamespace(special_tokens2ids_dict=ST)
    model.dream_embedding = SimpleNamespace(embed_len=Q)

    def sd_head(images_dm, enc, u_enc, *rest):
        sd_calls.append((enc, u_enc))
        return enc.float().pow(2).mean()
    sd_head.drop_prob = drop_prob
    self = SimpleNamespace(training=True, model=model, stable_diffusion_head=sd_head, lm_head=lambda h: F.linear(h, head_w),
                           loss_weight_lm=1.0, loss_weight_vm=10.0, vocab_size=V,
                           config=SimpleNamespace(max_position_embeddings=2048, output_attentions=False, output_hidden_states=False,
                                                  use_return_dict=True, special_tokens2ids_dict=ST, pretraining_tp=1, vocab_size=V,
                                                  loss_scale_schedule="none"))
    images_dm = torch.zeros(n_dm, 3, 2, 2)
    out = fwd(self, input_ids=ids, images_dm=images_dm, labels=labels)
    enc, u_enc = sd_calls[0]
    # (1) conditioning gather == oracle (== SplicePlan.cond_rows, tests/test_collator_cpu.py)
    assert torch.equal(enc, SO.gather_conditioning(ids, hidden, 93, Q, n_dm))
    # (2) null prompt: the id layout our `_null_prompt_states` builds, hidden rows [2, 2+Q), broadcast over the batch (:1420-1439)
    if drop_prob is None:
        assert u_enc is None and len(calls) == 1
    else:
        assert calls[1]["input_ids"].tolist() == [[1, 93] + [91] * Q + [94, 2]]
        assert torch.equal(u_enc, u_hidden[:, 2:2 + Q].repeat(n_dm, 1, 1))
    # (3) losses: masked-mean CE over shifted labels (:1456-1470) and vm * w_vm + lm * w_lm (:1486-1488)
    lm = O.lm_loss(F.linear(hidden, head_w).float(), labels)
    torch.testing.assert_close(torch.as_tensor(out.additional_log_info["lm_loss"]), lm, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(out.loss, enc.float().pow(2).mean() * 10.0 + lm * 1.0, rtol=1e-6, atol=1e-7)
