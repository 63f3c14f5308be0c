"""Launch the UNet step's kernels once at the C4 shapes (32 samples, 64x64x320 and 32x32x640 planes) for `ncu --set full`."""
import sys, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g).to(BF)
N, H, W, C = 32, 64, 64, 320
x = r(N, H, W, C); w = r(C); b = r(C)
for _ in range(2):
    y, st = ops.groupnorm(x, w, b, 32, 1e-5, True, return_stats=True)          # gn_partial + finalize + apply
    ops.groupnorm_bwd(y, x, w, b, st, 32, True)
    t2 = x.view(-1, C)
    ln = ops.layernorm_fwd(t2, w, b, 1e-5)
    ops.layernorm_bwd(ln, t2, w, 1e-5)
    f = r(N * H * W, 8 * C)
    gg = ops.geglu(f)
    ops.geglu_bwd(gg, f)
    wk = r(C, 9 * C) * 0.02
    ops.conv3x3(x, wk, bias=b, rowbias=r(N, C), residual=x)                    # implicit-GEMM conv, fused epilogue
    x2 = r(N, 32, 32, 640)
    ops.upsample2x(x2); ops.concat_channels(x, x)
    q = r(N, H * W, 5, 64); kv = r(N, 77, 2, 5, 64)
    ops.attn_fwd_cross(q, kv[:, :, 0], kv[:, :, 1])
    qkv = r(N, H * W, 3, 5, 64)
    ops.attn_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal=False)
    ops.linear(t2, r(2560, C) * 0.02, bias=r(2560))
    wp, bp = ops.geglu_permute(r(2560, C) * 0.02, r(2560))
    ops.linear_geglu(t2, wp, bp)                                                # FF-in GEMM with the GEGLU epilogue
    # fused CFG + DDIM update (sampler_step + step counter) on the 16 x 4 x 64 x 64 latents
    eps = torch.randn(32, 4, 64, 64, device="cuda", generator=g)
    lat = torch.randn(16, 4, 64, 64, device="cuda", generator=g)
    coef = torch.rand(50, 5, device="cuda", generator=g) + 0.1
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.sampler_step_(eps, lat, coef, step, 7.5, True, 0)
torch.cuda.synchronize()
print("done")
