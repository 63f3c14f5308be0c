"""TEST INFRASTRUCTURE ONLY — from-spec CPU restatement of the SD `AutoencoderKL` ENCODER (+ quant_conv, latent sampling).

PARITY UNPINNED against diffusers (not installed / vendored): restates diffusers 0.24 `models/autoencoder_kl.py`,
`models/vae.py` (Encoder, DiagonalGaussianDistribution), `unet_2d_blocks.py` (DownEncoderBlock2D, UNetMidBlock2D with one
single-head attention) per SURVEY.md Appendix A.3, with diffusers' module / state-dict key names.  Reference call site:
`self.vae.encode(images).latent_dist.sample() * scaling_factor`, modeling_plugins.py:511-512."""
import torch
import torch.nn as nn
import torch.nn.functional as F

SDVAE = dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_num_groups=32,
             scaling_factor=0.18215)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, groups, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class Attention(nn.Module):
    def __init__(self, c, groups, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        w = torch.softmax(q @ k.transpose(1, 2) / (C ** 0.5), dim=-1)
        o = self.to_out[0](w @ v).transpose(1, 2).reshape(B, C, H, W)
        return o + x


class _B(nn.Module):
    pass


class Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        ch, G, L = c["block_out_channels"], c["norm_num_groups"], c["layers_per_block"]
        self.conv_in = nn.Conv2d(c["in_channels"], ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, co in enumerate(ch):
            b = _B()
            cin, out = out, co
            b.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else out, out, G) for j in range(L)])
            if i < len(ch) - 1:
                b.downsamplers = nn.ModuleList([Downsample2D(out)])
            self.down_blocks.append(b)
        self.mid_block = _B()
        self.mid_block.resnets = nn.ModuleList([ResnetBlock2D(ch[-1], ch[-1], G), ResnetBlock2D(ch[-1], ch[-1], G)])
        self.mid_block.attentions = nn.ModuleList([Attention(ch[-1], G)])
        self.conv_norm_out = nn.GroupNorm(G, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * c["latent_channels"], 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            for r in b.resnets:
                x = r(x)
            if hasattr(b, "downsamplers"):
                x = b.downsamplers[0](x)
        x = self.mid_block.resnets[0](x)
        x = self.mid_block.attentions[0](x)
        x = self.mid_block.resnets[1](x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLEncoder(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        c = dict(SDVAE)
        c.update(cfg or {})
        self.cfg = c
        self.encoder = Encoder(c)
        self.quant_conv = nn.Conv2d(2 * c["latent_channels"], 2 * c["latent_channels"], 1)

    def encode_sample(self, images, z):
        """images [B,3,H,W] -> latents [B,4,H/8,W/8] = (mean + std * z) * scaling_factor  (latent_dist.sample() with injected z)."""
        moments = self.quant_conv(self.encoder(images))
        mean, logvar = moments.chunk(2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        return (mean + torch.exp(0.5 * logvar) * z) * self.cfg["scaling_factor"]


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Decoder(nn.Module):
    """diffusers vae.Decoder: conv_in -> mid (res, attn, res) -> 4 UpDecoderBlock2D (layers_per_block + 1 resnets each) -> GN/SiLU/conv_out."""

    def __init__(self, c):
        super().__init__()
        ch, G, L = c["block_out_channels"], c["norm_num_groups"], c["layers_per_block"]
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(c["latent_channels"], rev[0], 3, padding=1)
        self.mid_block = _B()
        self.mid_block.resnets = nn.ModuleList([ResnetBlock2D(rev[0], rev[0], G), ResnetBlock2D(rev[0], rev[0], G)])
        self.mid_block.attentions = nn.ModuleList([Attention(rev[0], G)])
        self.up_blocks = nn.ModuleList()
        out = rev[0]
        for i, co in enumerate(rev):
            b = _B()
            cin, out = out, co
            b.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else out, out, G) for j in range(L + 1)])
            if i < len(ch) - 1:
                b.upsamplers = nn.ModuleList([Upsample2D(out)])
            self.up_blocks.append(b)
        self.conv_norm_out = nn.GroupNorm(G, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], c["in_channels"], 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block.resnets[0](x)
        x = self.mid_block.attentions[0](x)
        x = self.mid_block.resnets[1](x)
        for b in self.up_blocks:
            for r in b.resnets:
                x = r(x)
            if hasattr(b, "upsamplers"):
                x = b.upsamplers[0](x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLDecoder(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        c = dict(SDVAE)
        c.update(cfg or {})
        self.cfg = c
        self.decoder = Decoder(c)
        self.post_quant_conv = nn.Conv2d(c["latent_channels"], c["latent_channels"], 1)

    def decode(self, latents):
        """`self.vae.decode(latents / scaling_factor)` (modeling_plugins.py:842)."""
        return self.decoder(self.post_quant_conv(latents / self.cfg["scaling_factor"]))
