"""TEST INFRASTRUCTURE ONLY — from-spec CPU restatement of the SD-2.1-base `UNet2DConditionModel` and schedulers.

PARITY UNPINNED against diffusers: the arithmetic of reference call sites modeling_plugins.py:556 / :815-821 lives in
`diffusers==0.24.0` (pyproject.toml:74), which is not vendored, not installed and not downloadable here.
This file restates the public architecture (models/unet_2d_condition.py, unet_2d_blocks.py,
resnet.py, transformer_2d.py, attention.py, embeddings.py, schedulers/scheduling_ddim.py / scheduling_ddpm.py) with
diffusers' module / state-dict key names; it is validated by (1) parameter count == 865 910 724 (the published SD-2.x UNet
size) and (2) state-dict key/shape agreement with the native module; compare against diffusers whenever it is available.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

SD21 = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
            attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024, norm_num_groups=32, norm_eps=1e-5,
            down_attn=(True, True, True, False), up_attn=(False, True, True, True))


def timestep_embedding(t, dim=320, max_period=10000):
    """Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb=1280, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Attention(nn.Module):
    def __init__(self, dim, heads, kv_dim=None):
        super().__init__()
        self.heads = heads
        kv_dim = dim if kv_dim is None else kv_dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(kv_dim, dim, bias=False)
        self.to_v = nn.Linear(kv_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, S, C = x.shape
        d = C // self.heads
        q = self.to_q(x).view(B, S, self.heads, d).transpose(1, 2)
        k = self.to_k(ctx).view(B, -1, self.heads, d).transpose(1, 2)
        v = self.to_v(ctx).view(B, -1, self.heads, d).transpose(1, 2)
        w = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
        o = (w @ v).transpose(1, 2).reshape(B, S, C)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, ctx_dim, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.norm(x).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.proj_in(h)
        h = self.transformer_blocks[0](h, ctx)
        h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        return h + x


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self):
        super().__init__()


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        c = dict(SD21)
        c.update(cfg or {})
        self.cfg = c
        ch = c["block_out_channels"]
        heads = c["attention_head_dim"]
        ctx = c["cross_attention_dim"]
        G = c["norm_num_groups"]
        L = c["layers_per_block"]
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(c["in_channels"], ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, co in enumerate(ch):
            b = _Block()
            cin = out
            out = co
            b.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else out, out, temb, G) for j in range(L)])
            if c["down_attn"][i]:
                b.attentions = nn.ModuleList([Transformer2DModel(out, heads[i], ctx, G) for _ in range(L)])
            if i < len(ch) - 1:
                b.downsamplers = nn.ModuleList([Downsample2D(out)])
            self.down_blocks.append(b)
        self.mid_block = _Block()
        self.mid_block.resnets = nn.ModuleList([ResnetBlock2D(ch[-1], ch[-1], temb, G), ResnetBlock2D(ch[-1], ch[-1], temb, G)])
        self.mid_block.attentions = nn.ModuleList([Transformer2DModel(ch[-1], heads[-1], ctx, G)])
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        rheads = list(reversed(heads))
        prev = rev[0]
        for i, co in enumerate(rev):
            b = _Block()
            skip_in = rev[min(i + 1, len(ch) - 1)]
            res = []
            for j in range(L + 1):
                skip = skip_in if j == L else co
                rin = prev if j == 0 else co
                res.append(ResnetBlock2D(rin + skip, co, temb, G))
            b.resnets = nn.ModuleList(res)
            if c["up_attn"][i]:
                b.attentions = nn.ModuleList([Transformer2DModel(co, rheads[i], ctx, G) for _ in range(L + 1)])
            if i < len(ch) - 1:
                b.upsamplers = nn.ModuleList([Upsample2D(co)])
            prev = co
            self.up_blocks.append(b)
        self.conv_norm_out = nn.GroupNorm(G, ch[0], eps=c["norm_eps"])
        self.conv_out = nn.Conv2d(ch[0], c["out_channels"], 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states):
        """sample [B,4,H,W] (NCHW), timestep scalar or [B], encoder_hidden_states [B,Q,1024] -> eps [B,4,H,W]."""
        B = sample.shape[0]
        t = torch.as_tensor(timestep)
        t = t.expand(B) if t.dim() == 0 else t
        temb = self.time_embedding(timestep_embedding(t, self.cfg["block_out_channels"][0]).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for b in self.down_blocks:
            for j, r in enumerate(b.resnets):
                x = r(x, temb)
                if hasattr(b, "attentions"):
                    x = b.attentions[j](x, encoder_hidden_states)
                skips.append(x)
            if hasattr(b, "downsamplers"):
                x = b.downsamplers[0](x)
                skips.append(x)
        x = self.mid_block.resnets[0](x, temb)
        x = self.mid_block.attentions[0](x, encoder_hidden_states)
        x = self.mid_block.resnets[1](x, temb)
        for b in self.up_blocks:
            for j, r in enumerate(b.resnets):
                x = r(torch.cat([x, skips.pop()], dim=1), temb)
                if hasattr(b, "attentions"):
                    x = b.attentions[j](x, encoder_hidden_states)
            if hasattr(b, "upsamplers"):
                x = b.upsamplers[0](x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


# ------------------------------------------------------------------------------------------------ schedulers (A.2)
def alphas_cumprod(num_train=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def set_timesteps(n, num_train=1000, steps_offset=1):
    ratio = num_train // n
    return (torch.arange(n) * ratio).flip(0) + steps_offset


def add_noise(x0, eps, t, ac):
    a = ac[t].view(-1, 1, 1, 1)
    return a.sqrt() * x0 + (1 - a).sqrt() * eps


def cfg_combine(eps_u, eps_c, g):
    return eps_u + g * (eps_c - eps_u)


def ddim_step(x_t, eps, t, ratio, ac):
    """eta = 0, set_alpha_to_one=False, prediction_type epsilon, clip_sample False."""
    a_t = ac[t]
    t_prev = t - ratio
    a_prev = ac[t_prev] if t_prev >= 0 else ac[0]
    x0 = (x_t - (1 - a_t).sqrt() * eps) / a_t.sqrt()
    return a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * eps


def ddpm_step(x_t, eps, t, ratio, ac, noise):
    """variance_type fixed_small; what the reference's sampler actually runs (modeling_plugins.py:379, :833)."""
    a_t = ac[t]
    t_prev = t - ratio
    a_prev = ac[t_prev] if t_prev >= 0 else torch.tensor(1.0)
    a_cur = a_t / a_prev
    b_cur = 1 - a_cur
    x0 = (x_t - (1 - a_t).sqrt() * eps) / a_t.sqrt()
    mu = (a_prev.sqrt() * b_cur / (1 - a_t)) * x0 + (a_cur.sqrt() * (1 - a_prev) / (1 - a_t)) * x_t
    var = torch.clamp((1 - a_prev) / (1 - a_t) * b_cur, min=1e-20)
    return mu + (var.sqrt() * noise if t > 0 else 0.0)


def diffusion_loss(unet, latents, cond, noise, t, ac, *, noise_offset=0.0, offset_noise=None, input_perturbation=0.0,
                   perturbation_noise=None, snr_gamma=None):
    """Restatement of `StableDiffusionHead.forward` after the VAE (reference modeling_plugins.py:520-572, epsilon prediction):
    offset noise (:521-523), input perturbation (:524-525, :533-534), add_noise (:534-536), UNet (:556), plain MSE (:559) or
    min-SNR weighted MSE (:561-572 with `_compute_snr` :468-491).  fp32 throughout; `cond` is already projected."""
    noise = noise.clone()
    if noise_offset:
        noise = noise + noise_offset * offset_noise.view(noise.shape[0], noise.shape[1], 1, 1)
    new_noise = noise + input_perturbation * perturbation_noise if input_perturbation else noise
    noisy = add_noise(latents, new_noise, t, ac)
    pred = unet(noisy, t, cond)
    if snr_gamma is None:
        return torch.nn.functional.mse_loss(pred.float(), noise.float(), reduction="mean")
    alpha = (ac ** 0.5)[t].float()
    sigma = ((1.0 - ac) ** 0.5)[t].float()
    snr = (alpha / sigma) ** 2
    w = torch.stack([snr, snr_gamma * torch.ones_like(snr)], dim=1).min(dim=1)[0] / snr
    loss = torch.nn.functional.mse_loss(pred.float(), noise.float(), reduction="none")
    loss = loss.mean(dim=list(range(1, len(loss.shape)))) * w
    return loss.mean()
