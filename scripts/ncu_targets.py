"""Launch each hot kernel once at C2 shapes (for `ncu --set full` captures; numbers printed here are NOT bench values)."""
import sys, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
BF = torch.bfloat16
T, H, I, nh, d, B, S = 16384, 4096, 11008, 32, 128, 8, 2048
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g).to(BF)
x, w = r(T, H), r(H, H) * 0.02
wgu = r(2 * I, H) * 0.02
for _ in range(2):
    y = ops.linear(x, w)                       # fwd NT
    gu = ops.linear(x, wgu)
    dx = ops.linear_dgrad(gu, wgu)             # dgrad NN
    dw = ops.linear_wgrad(gu, x)               # wgrad TN
    qkv = r(B, S, 3, nh, d)
    out, lse = ops.attn_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2])
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(out, qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
    nw = torch.ones(H, device="cuda", dtype=BF)
    h, rstd, xs = ops.rmsnorm_fwd(x, nw, 1e-6, add=y)
    ops.rmsnorm_bwd(h, xs, nw, rstd, dres=x)
    act = ops.swiglu_fwd(gu, I)
    ops.swiglu_bwd(act, gu, I)
    ops.rope_(qkv.view(T, 3 * H), r(2048, d), r(2048, d), torch.arange(S, device="cuda", dtype=torch.int32).repeat(B), 2 * nh, d)
    ops.add(x, y)
torch.cuda.synchronize()
print("done")
