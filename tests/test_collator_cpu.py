"""CPU tests of the index-map collator (dreamllm_b200/collator.py).

* same keys / values as the reference's `DataCollatorForDreamLLMDataset.__call__` (omni/data/builders/builder_dreamllm.py:467-482,
  restated inline: three `pad_sequence` calls + concatenation of the non-None images);
* the emitted `SplicePlan`, applied with plain index ops (what `copy_rows` / `segment_sum_rows` / `gather_rows` do on the GPU), reproduces
  the oracle's restatement of the reference splice loops (oracle/splice_oracle.py <- modeling_dreamllm.py:1082-1141, :1401-1418), forward
  and gradient, including surplus <im_start> tokens and ragged (right-padded) batches.
"""
from types import SimpleNamespace

import pytest
import torch

from dreamllm_b200.collator import DataCollatorForDreamLLMDataset, to_device
from oracle import splice_oracle as SO

PAD, IM_START, IM_PATCH, IM_END, DREAM_START, DREAM_END = 0, 500, 501, 502, 503, 504
P, Q, H = 6, 4, 8
TOK = SimpleNamespace(pad_token_id=PAD)


def _sample(g, n_text, image=False, dream=False, extra_im_start=False):
    ids = [1]
    lab = [-100]
    if image:
        ids += [IM_START] + [IM_PATCH] * P + [IM_END]
        lab += [-100] * (P + 2)
    t = torch.randint(5, 400, (n_text,), generator=g).tolist()
    ids += t
    lab += t
    if extra_im_start:                      # an <im_start> with no image behind it: left as a token embedding (:1122-1123)
        ids += [IM_START] + [IM_PATCH] * P + [IM_END]
        lab += [-100] * (P + 2)
    if dream:
        ids += [DREAM_START] + [IM_PATCH] * Q + [DREAM_END]
        lab += [-100] * (Q + 2)
    ids.append(2)
    lab.append(2)
    return dict(input_ids=torch.tensor(ids), attention_mask=torch.ones(len(ids), dtype=torch.long), labels=torch.tensor(lab),
                images=torch.randn(1, 3, 4, 4, generator=g) if image else None,
                images_dm=torch.randn(1, 3, 8, 8, generator=g) if dream else None)


def _reference_collate(examples):
    """builder_dreamllm.py:467-482, restated."""
    pad = torch.nn.utils.rnn.pad_sequence
    b = {k: [e[k] for e in examples] for k in examples[0]}
    b["input_ids"] = pad(b["input_ids"], batch_first=True, padding_value=PAD)
    b["attention_mask"] = pad(b["attention_mask"], batch_first=True, padding_value=0)
    b["labels"] = pad(b["labels"], batch_first=True, padding_value=-100)
    for k in ("images", "images_dm"):
        xs = [x for x in b[k] if x is not None]
        b[k] = torch.cat(xs, 0) if xs else None
    return b


def _collator(**kw):
    return DataCollatorForDreamLLMDataset(TOK, image_start_id=IM_START, dream_start_id=DREAM_START, clip_embed_len=P, dream_embed_len=Q, **kw)


def _batch(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [_sample(g, 11, image=True, dream=True), _sample(g, 3, image=True, extra_im_start=True), _sample(g, 17, dream=True),
            _sample(g, 5)]


def test_same_batch_as_the_reference_collator():
    ex = _batch()
    got, want = _collator()(ex), _reference_collate(ex)
    for k in ("input_ids", "attention_mask", "labels", "images", "images_dm"):
        assert torch.equal(got[k], want[k]), k
    assert got["attention_mask_has_padding"] is True
    assert torch.equal(got["seqlens"], want["attention_mask"].sum(-1).to(torch.int32))
    assert got["cu_seqlens"].tolist() == [0] + torch.cumsum(got["seqlens"], 0).tolist()
    sh = torch.full_like(want["labels"], -100)
    sh[:, :-1] = want["labels"][:, 1:]                                        # modeling_dreamllm.py:1456-1459
    assert torch.equal(got["shifted_labels"], sh)
    assert got["num_tokens"] == int(want["attention_mask"].sum()) and got["num_label_tokens"] == int((sh != -100).sum())
    assert got["input_ids_cpu"] is got["input_ids"]
    nopad = _collator()([ex[0], ex[0]])
    assert nopad["attention_mask_has_padding"] is False
    text_only = _collator()([ex[3], ex[3]])
    assert text_only["splice_plan"] is None and text_only["images"] is None


def test_pad_to_multiple_and_left_padding_rejected():
    ex = _batch(1)
    got = _collator(pad_to_multiple_of=64)(ex)
    assert got["input_ids"].shape[1] % 64 == 0
    ref = _reference_collate(ex)
    S = ref["input_ids"].shape[1]
    assert torch.equal(got["input_ids"][:, :S], ref["input_ids"]) and bool((got["input_ids"][:, S:] == PAD).all())
    assert bool((got["attention_mask"][:, S:] == 0).all()) and bool((got["labels"][:, S:] == -100).all())
    bad = dict(ex[0])
    bad["attention_mask"] = bad["attention_mask"].clone()
    bad["attention_mask"][0] = 0
    with pytest.raises(ValueError, match="right-padded"):
        _collator()([bad, ex[1]])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_plan_reproduces_the_reference_splice_and_gather(seed):
    ex = _batch(seed)
    b = _collator()(ex)
    plan = b["splice_plan"]
    ids = b["input_ids"]
    B, S = ids.shape
    g = torch.Generator().manual_seed(10 + seed)
    emb = torch.randn(B, S, H, generator=g, requires_grad=True)
    feats = torch.randn(b["images"].shape[0], P, H, generator=g, requires_grad=True)
    dq = torch.randn(1, Q, H, generator=g, requires_grad=True)
    assert plan.n_images_used == 2 and plan.n_dreams == b["images_dm"].shape[0] == 2

    # what the kernels do with the maps: row copies, then a row gather
    flat = emb.reshape(B * S, H)
    out = flat.index_copy(0, plan.dq_dst.long(), dq[0][plan.dq_src.long()])
    out = out.index_copy(0, plan.img_dst.long(), feats.reshape(-1, H)[plan.img_src.long()])
    cond = out[plan.cond_rows.long()].view(plan.n_dreams, Q, H)
    want = SO.splice(ids, emb, feats, dq, IM_START, DREAM_START)
    assert torch.equal(out.view(B, S, H), want)
    want_cond = SO.gather_conditioning(ids, want, DREAM_START, Q, plan.n_dreams)
    assert torch.equal(cond, want_cond)

    # gradient of the dream-query broadcast = CSR segment sum over (dq_seg, dq_rows); rows overwritten by a splice get no gradient
    w = torch.randn(B, S, H, generator=g)
    (want * w).sum().backward()
    dflat = w.reshape(B * S, H)
    ddq = torch.stack([dflat[plan.dq_rows[plan.dq_seg[q]:plan.dq_seg[q + 1]].long()].sum(0) for q in range(Q)])
    torch.testing.assert_close(ddq, dq.grad[0], rtol=1e-6, atol=1e-6)
    demb = dflat.clone()
    demb[plan.dq_dst.long()] = 0
    demb[plan.img_dst.long()] = 0
    torch.testing.assert_close(demb.view(B, S, H), emb.grad, rtol=0, atol=0)
    torch.testing.assert_close(dflat[plan.img_dst.long()].view(-1, P, H), feats.grad[: plan.n_images_used], rtol=0, atol=0)


def test_host_plan_moves_with_to_device():
    b = _collator()(_batch())
    moved = to_device(b, "cpu")
    assert moved["splice_plan"].device.type == "cpu" and moved["input_ids_cpu"] is b["input_ids_cpu"]
    assert torch.equal(moved["splice_plan"].cond_rows, b["splice_plan"].cond_rows) and moved["splice_plan"].n_dreams == 2
    assert all(getattr(b["splice_plan"], f).dtype == torch.int32 for f in b["splice_plan"]._FIELDS)
