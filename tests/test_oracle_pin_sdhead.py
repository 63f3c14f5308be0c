"""Pins the diffusion-loss restatement (oracle/unet_oracle.py `diffusion_loss`, `add_noise`, and the CFG-dropout row mix) to the LIVE
reference code: `StableDiffusionHead._compute_snr` and `.forward` are exec'd verbatim from
/root/reference/omni/models/dreamllm/modeling_plugins.py (:468-577) and run on CPU with stand-in `vae` / `noise_scheduler` / `projector` /
`unet` objects built from the oracle's modules.  Same seed => the reference's own RNG draws (randn_like, randn, randint, bernoulli — in
its order) are reproduced and fed to the restatement, so every branch (`noise_offset`, `input_perturbation`, `snr_gamma`, `drop_prob`) is
compared to the reference's arithmetic exactly.  What stays from-spec is only the inside of diffusers' UNet / VAE / scheduler classes
(not installable here — DESIGN.md §2).  Build container only: skipped where /root/reference does not exist (GPU box)."""
import os
import textwrap
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from oracle import unet_oracle as UO

REF = "/root/reference/omni/models/dreamllm/modeling_plugins.py"
pytestmark = pytest.mark.skipif(not os.path.isfile(REF), reason="reference checkout not present (GPU box)")
SMALL_UNET = dict(block_out_channels=(32, 64), attention_head_dim=(2, 4), cross_attention_dim=48, down_attn=(True, False),
                  up_attn=(False, True), norm_num_groups=8)
T = 1000


def _reference_head_class():
    src = open(REF).read()
    a = src.index("    def _compute_snr(self, timesteps):")
    b = src.index("    def check_inputs(", a)
    ns = {"torch": torch, "F": F}
    exec("from __future__ import annotations\n" + textwrap.dedent(src[a:b]), ns)
    return type("RefStableDiffusionHead", (), {"_compute_snr": ns["_compute_snr"], "forward": ns["forward"]})


class _Sched:
    """DDPMScheduler stand-in: the attributes / methods the reference forward touches (:529, :534-536, :551-554, :473)."""

    def __init__(self):
        self.config = SimpleNamespace(num_train_timesteps=T, prediction_type="epsilon")
        self.alphas_cumprod = UO.alphas_cumprod(T)

    def add_noise(self, x0, noise, t):
        return UO.add_noise(x0, noise, t, self.alphas_cumprod)


def _make(noise_offset, input_perturbation, snr_gamma, drop_prob):
    torch.manual_seed(0)
    unet = UO.UNet2DConditionModel(SMALL_UNET).eval()
    proj = torch.nn.Linear(40, 48)
    lat = torch.randn(3, 4, 8, 8)
    head = _reference_head_class()()
    head.vae = SimpleNamespace(encode=lambda images: SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda: lat / 0.18215)),
                               config=SimpleNamespace(scaling_factor=0.18215))
    head.noise_scheduler = _Sched()
    head.projector = lambda x: [proj(x)]
    head.unet = lambda x, t, c: SimpleNamespace(sample=unet(x, t, c))
    head.noise_offset, head.input_perturbation, head.snr_gamma, head.drop_prob = noise_offset, input_perturbation, snr_gamma, drop_prob
    head.embed_hidden_size, head.device, head.dtype = 40, torch.device("cpu"), torch.float32
    return head, unet, proj, lat


@pytest.mark.parametrize("noise_offset,input_perturbation,snr_gamma,drop_prob", [
    (0.0, 0.0, None, None), (0.1, 0.0, None, None), (0.0, 0.1, None, None), (0.0, 0.0, 5.0, None), (0.0, 0.0, None, 0.5),
    (0.05, 0.1, 5.0, 0.5)])
def test_diffusion_loss_restatement_equals_live_reference_forward(noise_offset, input_perturbation, snr_gamma, drop_prob):
    head, unet, proj, lat = _make(noise_offset, input_perturbation, snr_gamma, drop_prob)
    g = torch.Generator().manual_seed(1)
    images = torch.zeros(3, 3, 64, 64)
    enc = torch.randn(3, 5, 40, generator=g)
    u_enc = torch.randn(1, 5, 40, generator=g).repeat(3, 1, 1) if drop_prob is not None else None
    def replay(seed):
        """the reference's RNG draws in its order (:520-541)"""
        torch.manual_seed(seed)
        noise = torch.randn_like(lat)
        offset = torch.randn((3, 4, 1, 1)) if noise_offset else None
        pert = torch.randn_like(noise) if input_perturbation else None
        t = torch.randint(0, T, (3,)).long()
        mask = torch.bernoulli(torch.zeros(3) + drop_prob)[:, None, None] if drop_prob is not None else None
        return noise, offset, pert, t, mask

    seed = 1234
    if drop_prob is not None:                                          # a seed whose Bernoulli mask drops some rows and keeps others
        seed = next(s for s in range(1234, 1334) if 0 < float(replay(s)[4].sum()) < 3)
    torch.manual_seed(seed)
    with torch.no_grad():
        want = head.forward(images, enc, u_enc)
    noise, offset, pert, t, mask = replay(seed)
    latents = lat                                                      # vae stand-in: sample() * scaling_factor
    cond = enc
    if mask is not None:
        cond = (1.0 - mask) * enc + mask * u_enc                       # (:539-542) — the row select `_CfgDropFn` performs
    with torch.no_grad():
        got = UO.diffusion_loss(unet, latents, proj(cond), noise, t, UO.alphas_cumprod(T), noise_offset=noise_offset,
                                offset_noise=None if offset is None else offset.view(3, 4), input_perturbation=input_perturbation,
                                perturbation_noise=pert, snr_gamma=snr_gamma)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7)


def test_dummy_forward_of_the_reference_is_a_zero():
    """(:500-509) the reference's images=None branch only feeds DDP's unused-parameter check; our reducer needs no such pass."""
    head, _, _, _ = _make(0.0, 0.0, None, None)
    out = head.forward(None, None, None, dream_embeddings=torch.randn(1, 5, 40))
    assert float(out) == 0.0
