"""C3 rows: CLIP ViT tower (a11), projectors (a12), embedding splice + conditioning gather (a9/a10/a13) vs CPU oracles.

CLIP oracle = the installed `transformers.CLIPVisionModel` (eager attention, fp32) with the SAME weights — the arithmetic the
reference calls at modeling_plugins.py:321-323.  Splice oracle = `oracle/splice_oracle.py`, a restatement of the
reference's torch.where / torch.cat loops (modeling_dreamllm.py:1082-1141, :1401-1418); integer paths must be bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _hf_clip(cfg_kw):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(**cfg_kw)
    try:
        cfg._attn_implementation = "eager"
    except Exception:
        pass
    m = CLIPVisionModel(cfg).eval()
    for p in m.parameters():          # non-trivial LN / bias values
        if p.dim() == 1:
            p.data.add_(0.05 * torch.randn_like(p))
    return m


@pytest.mark.parametrize("cfg_kw,n_img", [
    (dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4, image_size=56, patch_size=14), 3),
    (dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=336, patch_size=14), 1),
])
def test_clip_tower_vs_transformers(cfg_kw, n_img):
    from dreamllm_b200.clip_vision import CLIPVisionConfigLite, CLIPVisionModel
    hf = _hf_clip(cfg_kw)
    ours = CLIPVisionModel(CLIPVisionConfigLite(**cfg_kw))
    ours.load_state_dict(hf.state_dict())            # identical keys
    ours = ours.to(device="cuda", dtype=BF)
    g = torch.Generator().manual_seed(1)
    img = torch.randn(n_img, 3, cfg_kw["image_size"], cfg_kw["image_size"], generator=g)
    with torch.no_grad():
        hfb = _hf_clip(cfg_kw).to(BF)
        hfb.load_state_dict(hf.state_dict())
        ref32 = hf(img, output_hidden_states=True).hidden_states[-2]
        refbf = hfb(img.to(BF), output_hidden_states=True).hidden_states[-2].float()
    got = ours.hidden_state(img.cuda().to(BF), -2).cpu().float()
    assert got.shape == ref32.shape
    e_o = (got - ref32).abs().mean()
    e_r = (refbf - ref32).abs().mean()
    assert float(e_o) <= 1.3 * float(e_r) + 1e-3 * float(ref32.abs().mean()), (float(e_o), float(e_r))
    # select_layer semantics: hidden_states[0] is the post-pre_layrnorm embedding
    with torch.no_grad():
        ref0 = hf(img, output_hidden_states=True).hidden_states[0]
        ref0b = hfb(img.to(BF), output_hidden_states=True).hidden_states[0].float()
    got0 = ours.hidden_state(img.cuda().to(BF), 0).cpu().float()
    e0_o, e0_r = float((got0 - ref0).abs().mean()), float((ref0b - ref0).abs().mean())
    assert e0_o <= 1.3 * e0_r + 1e-3 * float(ref0.abs().mean()), (e0_o, e0_r)          # like-for-like, as for the selected layer


@pytest.mark.parametrize("kind,depth,bias", [("linear", 1, True), ("linear", 1, False), ("mlp", 2, True)])
def test_projectors_fwd_bwd(kind, depth, bias):
    from dreamllm_b200.projector import build_projector
    torch.manual_seed(2)
    cfg = dict(projector=kind, freeze_projector=False, depth=depth, save_model_name="clip_vision_embedding", model_name_or_path=None)
    pr = build_projector(cfg, in_hidden_size=256, out_hidden_size=512, bias=bias)
    ref = [torch.nn.Linear(256, 512, bias=bias)] if kind == "linear" else [torch.nn.Linear(256, 512, bias=bias), torch.nn.GELU(), torch.nn.Linear(512, 512, bias=bias)]
    ref = torch.nn.Sequential(*ref)
    sd = pr.projector.state_dict()
    (ref[0] if kind == "linear" else ref).load_state_dict(sd)
    pr = pr.to(device="cuda", dtype=BF)
    x = torch.randn(3, 40, 256)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    xc = x.cuda().to(BF).requires_grad_(True)
    out = pr(xc)
    assert isinstance(out, list) and len(out) == 1               # list contract (mlp_projector.py:23-27)
    out[-1].backward(gy.cuda().to(BF))
    torch.testing.assert_close(out[-1].cpu().float(), yr.detach(), rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(xc.grad.cpu().float(), xr.grad, rtol=5e-2, atol=5e-2)
    lin0 = pr.projector if kind == "linear" else pr.projector[0]
    rl0 = ref[0]
    torch.testing.assert_close(lin0.weight.grad.cpu().float(), rl0.weight.grad, rtol=5e-2, atol=0.3)
    if bias:
        torch.testing.assert_close(lin0.bias.grad.cpu().float(), rl0.bias.grad, rtol=5e-2, atol=0.3)


def test_splice_and_conditioning_gather_bit_exact():
    from dreamllm_b200.modeling_plugins import build_splice_plan, gather_rows, splice_embeddings
    from oracle import splice_oracle as SO
    IM_START, DREAM_START, PATCH = 32003, 32006, 32002
    P, Q, H, S = 5, 4, 64, 40
    g = torch.Generator().manual_seed(5)
    rows = []
    # sample 0: two image spans; sample 1: dream span + image span (interleaved layout, builder_dreamllm.py:259-264); sample 2: text only
    rows.append([1, IM_START] + [PATCH] * P + [32004, 11, 12, IM_START] + [PATCH] * P + [32004, 13])
    rows.append([1, 21, DREAM_START] + [PATCH] * Q + [32007, IM_START] + [PATCH] * P + [32004, 22, 2])
    rows.append([1, 31, 32, 33, 2])
    ids = torch.full((3, S), 32000)
    for i, r in enumerate(rows):
        ids[i, :len(r)] = torch.tensor(r)
    emb = torch.randn(3, S, H, generator=g).to(BF)
    feats = torch.randn(3, P, H, generator=g).to(BF)
    dq = torch.randn(1, Q, H, generator=g).to(BF)
    want = SO.splice(ids, emb, feats, dq, IM_START, DREAM_START)
    plan = build_splice_plan(ids, IM_START, DREAM_START, P, Q, n_images=3, n_dream_images=1, device="cuda")
    e_c = emb.cuda().requires_grad_(True)
    f_c = feats.cuda().requires_grad_(True)
    d_c = dq.cuda().requires_grad_(True)
    got = splice_embeddings(e_c, f_c, d_c, plan)
    assert torch.equal(got.detach().cpu(), want)
    cond_want = SO.gather_conditioning(ids, want, DREAM_START, Q, n_dm=1)
    cond = gather_rows(got, plan.cond_rows)
    assert torch.equal(cond.detach().cpu().view(1, Q, H), cond_want)
    # gradients: compare with autograd through the oracle's torch.cat implementation (fp32)
    e32, f32, d32 = (t.float().requires_grad_(True) for t in (emb, feats, dq))
    w32 = SO.splice(ids, e32, f32, d32, IM_START, DREAM_START)
    gy = torch.randn(3, S, H, generator=g)
    gc = torch.randn(1, Q, H, generator=g)
    (w32 * gy).sum().backward(retain_graph=True)
    (SO.gather_conditioning(ids, w32, DREAM_START, Q, 1) * gc).sum().backward()
    ((got.float() * gy.cuda()).sum() + (cond.float().view(1, Q, H) * gc.cuda()).sum()).backward()
    torch.testing.assert_close(e_c.grad.cpu().float(), e32.grad, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(f_c.grad.cpu().float(), f32.grad, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(d_c.grad.cpu().float(), d32.grad, rtol=1e-2, atol=2e-2)
    # more <im_start> tokens than images: extra spans keep their token embeddings (reference :1122-1123)
    plan2 = build_splice_plan(ids, IM_START, DREAM_START, P, Q, n_images=2, n_dream_images=1, device="cuda")
    got2 = splice_embeddings(emb.cuda(), feats[:2].cuda(), dq.cuda(), plan2)
    assert torch.equal(got2.cpu(), SO.splice(ids, emb, feats[:2], dq, IM_START, DREAM_START))


def test_comprehension_step_with_images_runs_and_matches_manual_splice():
    """DreamLLMForCausalMLM.forward(images=...) == manual [CLIP -> projector -> splice -> inputs_embeds] path."""
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    from dreamllm_b200.modeling_plugins import CLIPVisionEmbedding, DreamEmbedding
    torch.manual_seed(0)
    cfg = DreamLLMConfig(vocab_size=32008, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2)
    m = DreamLLMForCausalMLM(cfg)
    clip = CLIPVisionEmbedding(dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4, image_size=56,
                                    patch_size=14), embed_hidden_size=256)
    dream = DreamEmbedding(num_dream_queries=8, embed_hidden_size=256)
    m.model.attach_plugins(clip, dream, image_start_id=32003, dream_start_id=32006)
    m = m.to(device="cuda", dtype=BF)
    P = clip.embed_len
    ids = torch.full((2, 64), 32000)
    ids[0, :P + 6] = torch.tensor([1, 32003] + [32002] * P + [32004, 5, 6, 2])
    ids[1, :P + 5] = torch.tensor([1, 32003] + [32002] * P + [32004, 7, 2])
    labels = ids.clone()
    labels[ids >= 32000] = -100
    am = (ids != 32000).long()
    imgs = torch.randn(2, 3, 56, 56).to(BF)
    out = m(input_ids=ids.cuda(), images=imgs.cuda(), attention_mask=am.cuda(), labels=labels.cuda(), input_ids_cpu=ids)
    out.loss.backward()
    assert torch.isfinite(out.loss) and clip.projector.projector.weight.grad is not None
    assert all(p.grad is None for p in clip.clip_vision_model.parameters())       # frozen tower
    loss2 = m(input_ids=ids.cuda(), images=imgs.cuda(), attention_mask=am.cuda(), labels=labels.cuda()).loss   # device-ids path
    assert torch.equal(out.loss, loss2)
