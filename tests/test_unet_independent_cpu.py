"""Independent second check of the from-spec UNet / scheduler oracle.

`oracle/unet_oracle.py` restates diffusers 0.24's `UNet2DConditionModel` from SURVEY.md Appendix A; diffusers itself cannot be installed
here (searched: no wheel / sdist / HF cache / vendored copy anywhere in the image), so parity of the UNet rows stays **unpinned vs
diffusers**.  What this file adds is a *second, separately written* interpretation of the same appendix that shares no code with the
oracle: a functional state-dict interpreter (no module classes) built on different primitives —

  * GroupNorm by explicit reshape / mean / var arithmetic (not `nn.GroupNorm`),
  * self-attention through `torch.nn.MultiheadAttention` with the q|k|v weights packed into `in_proj_weight`, cross-attention through
    `F.multi_head_attention_forward` with separate projection weights (`use_separate_proj_weight`),
  * GEGLU through the erf formula, convolutions through `F.unfold` + matmul (im2col) instead of `nn.Conv2d`,
  * the skip-connection bookkeeping as an explicit list of (tensor, channels) walked from Appendix A.1's table,
  * the sinusoidal timestep embedding through complex exponentials,
  * DDIM / DDPM coefficients in closed form from Appendix A.2 in float64.

Two independent readings agreeing to 1e-5 removes transcription slips (wrong skip order, eps, head split, GEGLU half order, time-embedding
flip) as an error source; it cannot remove a shared misreading of diffusers, and the docs say so.
"""
import math

import torch
import torch.nn.functional as F

from oracle import unet_oracle as UO

SMALL = dict(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=96)


# ------------------------------------------------------------------------------------------------ primitives (second implementation)
def group_norm(x, w, b, groups, eps):
    N, C, H, W = x.shape
    g = x.reshape(N, groups, (C // groups) * H * W).double()
    mu = g.mean(-1, keepdim=True)
    var = ((g - mu) ** 2).mean(-1, keepdim=True)
    y = ((g - mu) / torch.sqrt(var + eps)).reshape(N, C, H, W).to(x.dtype)
    return y * w[None, :, None, None] + b[None, :, None, None]


def conv2d_im2col(x, w, b, stride=1, padding=1):
    N, C, H, W = x.shape
    Co, _, kh, kw = w.shape
    cols = F.unfold(x, (kh, kw), padding=padding, stride=stride)                    # [N, C*kh*kw, L]
    Ho, Wo = (H + 2 * padding - kh) // stride + 1, (W + 2 * padding - kw) // stride + 1
    y = torch.einsum("ok,nkl->nol", w.reshape(Co, -1), cols).reshape(N, Co, Ho, Wo)
    return y + b[None, :, None, None]


def silu(x):
    return x * torch.sigmoid(x)


def sinusoid(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float64) / half)
    z = torch.exp(1j * t.double()[:, None] * freqs[None, :])                         # cos + i sin
    return torch.cat([z.real, z.imag], -1).float()                                   # flip_sin_to_cos=True -> [cos | sin]


def self_attention(sd, pre, x, heads):
    C = x.shape[-1]
    mha = torch.nn.MultiheadAttention(C, heads, bias=True, batch_first=True)
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.cat([sd[pre + "to_q.weight"], sd[pre + "to_k.weight"], sd[pre + "to_v.weight"]]))
        mha.in_proj_bias.zero_()                                                     # to_q / to_k / to_v have no bias (Appendix A.1)
        mha.out_proj.weight.copy_(sd[pre + "to_out.0.weight"])
        mha.out_proj.bias.copy_(sd[pre + "to_out.0.bias"])
    return mha(x, x, x, need_weights=False)[0]


def cross_attention(sd, pre, x, ctx, heads):
    C = x.shape[-1]
    out, _ = F.multi_head_attention_forward(
        x.transpose(0, 1), ctx.transpose(0, 1), ctx.transpose(0, 1), C, heads, in_proj_weight=None, in_proj_bias=None, bias_k=None, bias_v=None,
        add_zero_attn=False, dropout_p=0.0, out_proj_weight=sd[pre + "to_out.0.weight"], out_proj_bias=sd[pre + "to_out.0.bias"],
        training=False, need_weights=False, use_separate_proj_weight=True, q_proj_weight=sd[pre + "to_q.weight"],
        k_proj_weight=sd[pre + "to_k.weight"], v_proj_weight=sd[pre + "to_v.weight"])
    return out.transpose(0, 1)


def layer_norm(sd, pre, x):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * sd[pre + "weight"] + sd[pre + "bias"]


def resnet(sd, pre, x, temb, groups):
    h = conv2d_im2col(silu(group_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], groups, 1e-5)), sd[pre + "conv1.weight"],
                      sd[pre + "conv1.bias"])
    h = h + (silu(temb) @ sd[pre + "time_emb_proj.weight"].T + sd[pre + "time_emb_proj.bias"])[:, :, None, None]
    h = conv2d_im2col(silu(group_norm(h, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], groups, 1e-5)), sd[pre + "conv2.weight"],
                      sd[pre + "conv2.bias"])
    if pre + "conv_shortcut.weight" in sd:
        x = conv2d_im2col(x, sd[pre + "conv_shortcut.weight"], sd[pre + "conv_shortcut.bias"], padding=0)
    return x + h


def transformer(sd, pre, x, ctx, heads, groups):
    N, C, H, W = x.shape
    h = group_norm(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"], groups, 1e-6)       # eps 1e-6 here, 1e-5 in the resnets
    h = h.flatten(2).transpose(1, 2)                                                      # [N, HW, C]
    h = h @ sd[pre + "proj_in.weight"].T + sd[pre + "proj_in.bias"]
    b = pre + "transformer_blocks.0."
    h = h + self_attention(sd, b + "attn1.", layer_norm(sd, b + "norm1.", h), heads)
    h = h + cross_attention(sd, b + "attn2.", layer_norm(sd, b + "norm2.", h), ctx, heads)
    f = layer_norm(sd, b + "norm3.", h) @ sd[b + "ff.net.0.proj.weight"].T + sd[b + "ff.net.0.proj.bias"]
    inner = f.shape[-1] // 2
    val, gate = f[..., :inner], f[..., inner:]                                            # chunk(2): first half is the value
    f = val * (0.5 * gate * (1.0 + torch.erf(gate / math.sqrt(2.0))))
    h = h + (f @ sd[b + "ff.net.2.weight"].T + sd[b + "ff.net.2.bias"])
    h = h @ sd[pre + "proj_out.weight"].T + sd[pre + "proj_out.bias"]
    return h.transpose(1, 2).reshape(N, C, H, W) + x


def unet_forward(sd, cfg, sample, t, ctx):
    """Appendix A.1 walked literally: conv_in, 4 down stages (2 resnets each, attention on the first 3, stride-2 conv between), mid
    (res, attn, res), 4 up stages (3 resnets each on cat(x, skip), attention on the last 3, nearest-2x + conv between), GN-SiLU-conv_out."""
    ch, heads, G = cfg["block_out_channels"], cfg["attention_head_dim"], cfg.get("norm_num_groups", 32)
    temb = sinusoid(t, ch[0]) @ sd["time_embedding.linear_1.weight"].T + sd["time_embedding.linear_1.bias"]
    temb = silu(temb) @ sd["time_embedding.linear_2.weight"].T + sd["time_embedding.linear_2.bias"]
    x = conv2d_im2col(sample, sd["conv_in.weight"], sd["conv_in.bias"])
    stack = [x]
    for i in range(4):
        for j in range(2):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}.", x, temb, G)
            if i < 3:
                x = transformer(sd, f"down_blocks.{i}.attentions.{j}.", x, ctx, heads[i], G)
            stack.append(x)
        if i < 3:
            x = conv2d_im2col(x, sd[f"down_blocks.{i}.downsamplers.0.conv.weight"], sd[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
            stack.append(x)
    assert len(stack) == 12                                                               # Appendix A.1 "skip stack"
    x = resnet(sd, "mid_block.resnets.0.", x, temb, G)
    x = transformer(sd, "mid_block.attentions.0.", x, ctx, heads[3], G)
    x = resnet(sd, "mid_block.resnets.1.", x, temb, G)
    for i in range(4):
        for j in range(3):
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}.", torch.cat([x, stack.pop()], 1), temb, G)
            if i > 0:
                x = transformer(sd, f"up_blocks.{i}.attentions.{j}.", x, ctx, heads[3 - i], G)
        if i < 3:
            x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)                 # nearest 2x
            x = conv2d_im2col(x, sd[f"up_blocks.{i}.upsamplers.0.conv.weight"], sd[f"up_blocks.{i}.upsamplers.0.conv.bias"])
    assert not stack
    x = silu(group_norm(x, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], G, 1e-5))
    return conv2d_im2col(x, sd["conv_out.weight"], sd["conv_out.bias"])


# ------------------------------------------------------------------------------------------------ tests
def test_second_interpretation_agrees_with_the_oracle_unet():
    torch.manual_seed(0)
    ref = UO.UNet2DConditionModel(SMALL).eval()
    with torch.no_grad():
        for p in ref.parameters():                      # non-trivial norm weights / biases everywhere
            p.add_(torch.randn_like(p) * 0.05)
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    cfg = dict(UO.SD21)
    cfg.update(SMALL)
    g = torch.Generator().manual_seed(1)
    sample = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(2, 7, 96, generator=g)
    t = torch.tensor([981, 3])
    with torch.no_grad():
        want = ref(sample, t, ctx)
        got = unet_forward(sd, cfg, sample, t, ctx)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)


def test_full_size_block_shapes_follow_appendix_a1():
    """The SD-2.1 table itself: channel plan of every resnet (incl. the concatenated up-path inputs 2560/1920/1280/960/640) and the
    865 910 724 parameter total, recomputed here from the appendix numbers alone."""
    ch, L, temb, ctx = (320, 640, 1280, 1280), 2, 1280, 1024

    def res(cin, cout):
        n = 2 * cin + cin * cout * 9 + cout + temb * cout + cout + 2 * cout + cout * cout * 9 + cout
        return n + (cin * cout + cout if cin != cout else 0)

    def tr(c):
        attn1 = 3 * c * c + c * c + c
        attn2 = c * c + 2 * ctx * c + c * c + c
        ff = c * 8 * c + 8 * c + 4 * c * c + c
        return 2 * c + 2 * (c * c + c) + 3 * 2 * c + attn1 + attn2 + ff
    total = 4 * 320 * 9 + 320 + (320 * temb + temb) + (temb * temb + temb)
    cin = 320
    for i, co in enumerate(ch):
        for j in range(L):
            total += res(cin if j == 0 else co, co) + (tr(co) if i < 3 else 0)
        if i < 3:
            total += co * co * 9 + co
        cin = co
    total += 2 * res(1280, 1280) + tr(1280)
    rev, prev, ups = (1280, 1280, 640, 320), 1280, []
    for i, co in enumerate(rev):
        skip_in = rev[min(i + 1, 3)]
        for j in range(L + 1):
            rin = (prev if j == 0 else co) + (skip_in if j == L else co)
            ups.append(rin)
            total += res(rin, co) + (tr(co) if i > 0 else 0)
        if i < 3:
            total += co * co * 9 + co
        prev = co
    total += 2 * 320 + 320 * 4 * 9 + 4
    assert ups == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]
    assert total == 865_910_724
    with torch.device("meta"):
        assert sum(p.numel() for p in UO.UNet2DConditionModel().parameters()) == total


def test_scheduler_closed_forms_agree_with_the_oracle():
    """Appendix A.2 in float64 closed form vs oracle ddim_step / ddpm_step / add_noise / set_timesteps."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    abar = torch.cumprod(1 - betas, 0)
    torch.testing.assert_close(UO.alphas_cumprod().double(), abar, rtol=1e-6, atol=0)
    N = 50
    ts = UO.set_timesteps(N)
    assert ts.tolist() == [981 - 20 * i for i in range(N)]
    g = torch.Generator().manual_seed(0)
    x, e, z = (torch.randn(2, 4, 8, 8, generator=g).double() for _ in range(3))
    ac = UO.alphas_cumprod()
    for t in (981, 501, 21, 1):
        tp = t - 20
        a_t = abar[t]
        x0 = (x - (1 - a_t).sqrt() * e) / a_t.sqrt()
        a_p = abar[tp] if tp >= 0 else abar[0]                                     # DDIM: set_alpha_to_one = False
        ddim = a_p.sqrt() * x0 + (1 - a_p).sqrt() * e
        torch.testing.assert_close(UO.ddim_step(x.float(), e.float(), t, 20, ac).double(), ddim, rtol=1e-4, atol=1e-5)
        a_p = abar[tp] if tp >= 0 else torch.tensor(1.0, dtype=torch.float64)     # DDPM: alpha_prod_t_prev = one
        a_cur = a_t / a_p
        mu = a_p.sqrt() * (1 - a_cur) / (1 - a_t) * x0 + a_cur.sqrt() * (1 - a_p) / (1 - a_t) * x
        var = torch.clamp((1 - a_p) / (1 - a_t) * (1 - a_cur), min=1e-20)
        ddpm = mu + (var.sqrt() * z if t > 0 else 0)
        torch.testing.assert_close(UO.ddpm_step(x.float(), e.float(), t, 20, ac, z.float()).double(), ddpm, rtol=1e-4, atol=1e-5)
    tt = torch.tensor([0, 999])
    noisy = abar[tt].sqrt()[:, None, None, None] * x + (1 - abar[tt]).sqrt()[:, None, None, None] * e
    torch.testing.assert_close(UO.add_noise(x.float(), e.float(), tt, ac).double(), noisy, rtol=1e-4, atol=1e-5)
