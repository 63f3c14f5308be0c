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
