"""C5 rows (a14, VAE f-row): VAE encoder vs its from-spec oracle; StableDiffusionHead.forward loss + gradients vs oracle autograd;
full creation step (LLM -> dream-query conditioning -> SD head) wiring."""
import pytest
import torch

from oracle import unet_oracle as UO
from oracle import vae_oracle as VO

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
SMALL_UNET = dict(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128)
SMALL_VAE = dict(block_out_channels=(64, 128, 128, 128))


def _rel(a, ref):
    return float((a - ref).abs().mean() / ref.abs().mean())


def test_vae_encoder_param_count():
    from dreamllm_b200.vae import AutoencoderKLEncoder
    with torch.device("meta"):
        m = AutoencoderKLEncoder()
    n = sum(p.numel() for p in m.encoder.parameters())
    assert abs(n - 34.16e6) < 0.05e6, n          # SD VAE encoder = 34.1 M


@pytest.mark.parametrize("cfg,B,R", [(SMALL_VAE, 2, 64), (None, 1, 256)])
def test_vae_encode_vs_oracle(cfg, B, R):
    from dreamllm_b200.vae import AutoencoderKLEncoder
    torch.manual_seed(0)
    ref = VO.AutoencoderKLEncoder(cfg).eval()
    for n, p in ref.named_parameters():
        if p.dim() == 1:
            p.data.add_(0.05 * torch.randn_like(p))
    ours = AutoencoderKLEncoder(cfg)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(device="cuda", dtype=BF)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(B, 3, R, R, generator=g) * 2 - 1
    z = torch.randn(B, 4, R // 8, R // 8, generator=g)
    with torch.no_grad():
        want = ref.encode_sample(img, z)
    got = ours.encode_sample(img.cuda(), z=z.cuda()).cpu()
    e = _rel(got, want)
    print(f"vae rel err {e:.4f}")
    assert got.shape == want.shape and e < 4e-2, e


def test_sd_head_loss_and_grads_vs_oracle_autograd():
    from dreamllm_b200.modeling_plugins import StableDiffusionHead
    torch.manual_seed(2)
    cfg = dict(SMALL_UNET)
    cfg["vae"] = SMALL_VAE
    head = StableDiffusionHead(cfg, embed_hidden_size=256)
    ref_unet = UO.UNet2DConditionModel(SMALL_UNET).eval()
    ref_unet.load_state_dict(head.unet.state_dict())
    ref_vae = VO.AutoencoderKLEncoder(SMALL_VAE).eval()
    ref_vae.load_state_dict(head.vae.state_dict())
    wproj = head.projector.projector.weight.detach().clone()
    head = head.to(device="cuda", dtype=BF)
    g = torch.Generator().manual_seed(3)
    Nd, Q = 2, 8
    img = torch.rand(Nd, 3, 128, 128, generator=g) * 2 - 1
    enc = torch.randn(Nd, Q, 256, generator=g)
    vz = torch.randn(Nd, 4, 16, 16, generator=g)
    noise = torch.randn(Nd, 4, 16, 16, generator=g)
    t = torch.tensor([17, 803])
    # oracle (fp32): the reference's forward, restated
    e32 = enc.clone().requires_grad_(True)
    w32 = wproj.clone().requires_grad_(True)
    with torch.no_grad():
        lat = ref_vae.encode_sample(img, vz)
    noisy = UO.add_noise(lat, noise, t, UO.alphas_cumprod())
    pred = ref_unet(noisy, t, torch.nn.functional.linear(e32, w32))
    want = torch.nn.functional.mse_loss(pred.float(), noise.float(), reduction="mean")
    want.backward()
    # ours
    ec = enc.cuda().to(BF).requires_grad_(True)
    loss = head(img.cuda().to(BF), ec, vae_noise=vz.cuda(), noise=noise.cuda(), timesteps=t.cuda())
    loss.backward()
    assert abs(float(loss) - float(want)) < 5e-2 * float(want), (float(loss), float(want))
    assert _rel(ec.grad.cpu().float(), e32.grad) < 0.15
    assert _rel(head.projector.projector.weight.grad.cpu().float(), w32.grad) < 0.15
    assert all(p.grad is None for p in head.unet.parameters()) and all(p.grad is None for p in head.vae.parameters())


def test_creation_step_end_to_end():
    """stage-1 creation layout (builder_dreamllm.py:210-218): [bos, text, <dream_start>, Q x <im_patch>, <dream_end>, eos]; trainable =
    dream queries + SD projector; LLM / UNet / VAE frozen.  Gradients must reach dream_queries through UNet + all LLM layers."""
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    from dreamllm_b200.modeling_plugins import CLIPVisionEmbedding, DreamEmbedding, StableDiffusionHead
    torch.manual_seed(0)
    cfg = DreamLLMConfig(vocab_size=32008, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2)
    m = DreamLLMForCausalMLM(cfg)
    Q = 8
    dream = DreamEmbedding(num_dream_queries=Q, embed_hidden_size=256)
    sdc = dict(SMALL_UNET)
    sdc["vae"] = SMALL_VAE
    m.stable_diffusion_head = StableDiffusionHead(sdc, embed_hidden_size=256)
    m.model.attach_plugins(None, dream, image_start_id=32003, dream_start_id=32006)
    m = m.to(device="cuda", dtype=BF)
    for p in m.parameters():
        p.requires_grad_(False)
    dream.dream_queries.requires_grad_(True)
    m.stable_diffusion_head.projector.requires_grad_(True)
    m.train()
    B, S = 2, 32
    ids = torch.full((B, S), 32000)
    for b in range(B):
        row = [1, 11 + b, 12, 32006] + [32002] * Q + [32007, 2]
        ids[b, :len(row)] = torch.tensor(row)
    labels = torch.full((B, S), -100)
    am = (ids != 32000).long()
    imgs = (torch.rand(B, 3, 128, 128) * 2 - 1).to(BF)
    out = m(input_ids=ids.cuda(), images_dm=imgs.cuda(), attention_mask=am.cuda(), labels=labels.cuda(), input_ids_cpu=ids)
    out.loss.backward()
    assert torch.isfinite(out.loss) and float(out.additional_log_info["vm_loss"]) > 0
    assert float(out.additional_log_info["lm_loss"]) == 0.0                       # all labels -100 (creation-only)
    gq = dream.dream_queries.grad
    assert gq is not None and torch.isfinite(gq.float()).all() and float(gq.float().abs().sum()) > 0
    assert m.stable_diffusion_head.projector.projector.weight.grad is not None
    assert m.model.layers[0].mlp.down_proj.weight.grad is None                    # frozen LLM: dgrad only


@pytest.mark.parametrize("cfg,B,R", [(SMALL_VAE, 2, 8), (None, 1, 32)])
def test_vae_decode_vs_oracle(cfg, B, R):
    from dreamllm_b200.vae import AutoencoderKLDecoder
    torch.manual_seed(4)
    ref = VO.AutoencoderKLDecoder(cfg).eval()
    for n, p in ref.named_parameters():
        if p.dim() == 1:
            p.data.add_(0.05 * torch.randn_like(p))
    ours = AutoencoderKLDecoder(cfg)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(device="cuda", dtype=BF)
    lat = torch.randn(B, 4, R, R, generator=torch.Generator().manual_seed(5)) * 0.18215
    with torch.no_grad():
        want = ref.decode(lat)
    got = ours.decode(lat.cuda()).cpu()
    e = _rel(got, want)
    print(f"vae decode rel err {e:.4f}")
    assert got.shape == want.shape == (B, 3, 8 * R, 8 * R) and e < 5e-2, e


def _small_head(seed, **kw):
    from dreamllm_b200.modeling_plugins import StableDiffusionHead
    torch.manual_seed(seed)
    cfg = dict(SMALL_UNET)
    cfg["vae"] = SMALL_VAE
    head = StableDiffusionHead(cfg, embed_hidden_size=256, **kw)
    ref_unet = UO.UNet2DConditionModel(SMALL_UNET).eval()
    ref_unet.load_state_dict(head.unet.state_dict())
    wproj = head.projector.projector.weight.detach().clone()
    return head.to(device="cuda", dtype=BF), ref_unet, wproj


@pytest.mark.parametrize("opts", [dict(noise_offset=0.1), dict(input_perturbation=0.1), dict(snr_gamma=5.0),
                                  dict(noise_offset=0.05, input_perturbation=0.1, snr_gamma=1.0)],
                         ids=["offset", "perturb", "minsnr", "all"])
def test_sd_head_noise_options_vs_oracle(opts):
    """reference :521-525 / :533-534 / :561-572 — the training-noise options, each against the fp32 restatement."""
    head, ref_unet, wproj = _small_head(6, **opts)
    g = torch.Generator().manual_seed(7)
    Nd, Q = 2, 8
    lat = torch.randn(Nd, 4, 16, 16, generator=g) * 0.5
    enc = torch.randn(Nd, Q, 256, generator=g)
    noise = torch.randn(Nd, 4, 16, 16, generator=g)
    off = torch.randn(Nd, 4, 1, 1, generator=g)
    pert = torch.randn(Nd, 4, 16, 16, generator=g)
    t = torch.tensor([950, 41])                     # snr(950) << gamma (weight 1), snr(41) >> gamma (weight gamma / snr)
    e32 = enc.clone().requires_grad_(True)
    want = UO.diffusion_loss(ref_unet, lat, torch.nn.functional.linear(e32, wproj), noise, t, UO.alphas_cumprod(),
                             noise_offset=opts.get("noise_offset", 0.0), offset_noise=off,
                             input_perturbation=opts.get("input_perturbation", 0.0), perturbation_noise=pert,
                             snr_gamma=opts.get("snr_gamma"))
    want.backward()
    ec = enc.cuda().to(BF).requires_grad_(True)
    loss = head(None, ec, latents=lat.cuda(), noise=noise.cuda(), timesteps=t.cuda(), offset_noise=off.cuda(),
                perturbation_noise=pert.cuda())
    loss.backward()
    print(f"{opts}: loss {float(loss):.5f} want {float(want):.5f}")
    assert abs(float(loss) - float(want)) < 5e-2 * float(want), (float(loss), float(want))
    assert _rel(ec.grad.cpu().float(), e32.grad) < 0.15


def test_mse_minsnr_kernel_vs_formula():
    from dreamllm_b200 import ops
    g = torch.Generator().manual_seed(8)
    B = 5
    pred = torch.randn(B, 4, 8, 8, generator=g)
    tgt = torch.randn(B, 4, 8, 8, generator=g)
    t = torch.tensor([0, 100, 500, 900, 999], dtype=torch.int32)
    ac = UO.alphas_cumprod()
    p32 = pred.clone().requires_grad_(True)
    snr = ac[t.long()] / (1 - ac[t.long()])
    w = torch.minimum(snr, torch.tensor(5.0)) / snr
    want = (((p32 - tgt) ** 2).mean(dim=(1, 2, 3)) * w).mean()
    want.backward()
    loss, dpred = ops.mse_minsnr_fwd_bwd(pred.cuda(), tgt.cuda(), t.cuda(), ac.cuda(), 5.0)
    assert abs(float(loss) - float(want)) < 1e-3 * float(want)
    assert torch.allclose(dpred.cpu(), p32.grad, rtol=2e-3, atol=1e-7)


def test_sd_head_cfg_dropout_mix():
    """reference :539-543: dropped samples take the null-prompt states; loss and gradients equal a run on the hand-mixed conditioning."""
    head, _, _ = _small_head(9, drop_prob=0.1)
    g = torch.Generator().manual_seed(10)
    Nd, Q = 3, 8
    lat = (torch.randn(Nd, 4, 16, 16, generator=g) * 0.5).cuda()
    noise = torch.randn(Nd, 4, 16, 16, generator=g).cuda()
    t = torch.tensor([10, 500, 900]).cuda()
    enc = torch.randn(Nd, Q, 256, generator=g).cuda().to(BF)
    u = torch.randn(1, Q, 256, generator=g).cuda().to(BF)
    mask = torch.tensor([1.0, 0.0, 1.0])
    e1, u1 = enc.clone().requires_grad_(True), u.clone().requires_grad_(True)
    l1 = head(None, e1, u1, latents=lat, noise=noise, timesteps=t, drop_mask=mask)
    l1.backward()
    mixed = torch.stack([u[0], enc[1], u[0]]).clone().requires_grad_(True)
    l2 = head(None, mixed, None, latents=lat, noise=noise, timesteps=t)
    l2.backward()
    assert abs(float(l1) - float(l2)) <= 1e-4 * abs(float(l2)), (float(l1), float(l2))
    assert _rel(e1.grad[1].float(), mixed.grad[1].float()) < 1e-2
    assert float(e1.grad[0].abs().sum()) == 0 and float(e1.grad[2].abs().sum()) == 0
    want_u = (mixed.grad[0].float() + mixed.grad[2].float())
    assert _rel(u1.grad[0].float(), want_u) < 1e-2
    # [Nd, Q, H]-shaped null states (the reference's `.repeat`, modeling_dreamllm.py:1439) give the same loss
    e3, u3 = enc.clone().requires_grad_(True), u.repeat(Nd, 1, 1).clone().requires_grad_(True)
    l3 = head(None, e3, u3, latents=lat, noise=noise, timesteps=t, drop_mask=mask)
    l3.backward()
    assert abs(float(l3) - float(l1)) <= 1e-4 * abs(float(l1))
    assert _rel(u3.grad[0].float(), mixed.grad[0].float()) < 1e-2 and float(u3.grad[1].abs().sum()) == 0
    # drop_prob set but nothing dropped -> identical to the plain path
    e4 = enc.clone().requires_grad_(True)
    l4 = head(None, e4, u, latents=lat, noise=noise, timesteps=t, drop_mask=torch.zeros(Nd))
    e5 = enc.clone().requires_grad_(True)
    l5 = head(None, e5, None, latents=lat, noise=noise, timesteps=t)
    assert abs(float(l4) - float(l5)) <= 1e-4 * abs(float(l5))


def test_creation_step_with_cfg_dropout_null_prompt_pass():
    """reference modeling_dreamllm.py:1420-1441: with `drop_prob` set the LLM runs a second pass over the null prompt
    [bos, <dream_start>, Q x <im_patch>, <dream_end>, eos]; its states replace the conditioning of dropped samples.  The literal
    <im_patch> embedding row only gets a gradient through that pass (main-pass patch positions are overwritten by dream queries)."""
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    from dreamllm_b200.modeling_plugins import DreamEmbedding, StableDiffusionHead
    torch.manual_seed(0)
    cfg = DreamLLMConfig(vocab_size=32008, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2)
    m = DreamLLMForCausalMLM(cfg)
    Q = 8
    dream = DreamEmbedding(num_dream_queries=Q, embed_hidden_size=256)
    sdc = dict(SMALL_UNET)
    sdc["vae"] = SMALL_VAE
    m.stable_diffusion_head = StableDiffusionHead(sdc, embed_hidden_size=256, drop_prob=0.1)
    m.model.attach_plugins(None, dream, image_start_id=32003, dream_start_id=32006)
    m = m.to(device="cuda", dtype=BF)
    m.train()
    B, S = 2, 32
    ids = torch.full((B, S), 32000)
    for b in range(B):
        row = [1, 11 + b, 12, 32006] + [32002] * Q + [32007, 2]
        ids[b, :len(row)] = torch.tensor(row)
    labels = torch.full((B, S), -100)
    am = (ids != 32000).long()
    imgs = (torch.rand(B, 3, 128, 128) * 2 - 1).to(BF)

    def step(mask):
        m.zero_grad(set_to_none=True)
        torch.manual_seed(5)
        out = m(input_ids=ids.cuda(), images_dm=imgs.cuda(), attention_mask=am.cuda(), labels=labels.cuda(), input_ids_cpu=ids,
                sd_kwargs=dict(drop_mask=mask, timesteps=torch.tensor([100, 700]).cuda()))
        out.loss.backward()
        assert torch.isfinite(out.loss)
        return m.model.embed_tokens.weight.grad[32002].float().abs().sum().item(), dream.dream_queries.grad.float().abs().sum().item()

    patch_g, dq_g = step(torch.tensor([1.0, 0.0]))
    assert patch_g > 0 and dq_g > 0                  # sample 0 conditioned on the null prompt, sample 1 on its dream queries
    patch_g0, dq_g0 = step(torch.tensor([0.0, 0.0]))
    assert patch_g0 == 0 and dq_g0 > 0               # nothing dropped: the null pass contributes no gradient
    patch_g1, dq_g1 = step(torch.tensor([1.0, 1.0]))
    assert patch_g1 > 0 and dq_g1 == 0               # everything dropped: dream queries are cut off from the diffusion loss
