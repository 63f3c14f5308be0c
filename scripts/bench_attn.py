import sys, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
BF = torch.bfloat16
import os
B, S, nh, d = int(os.environ.get("ATTN_B", 8)), 2048, 32, 128
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, S, 3, nh, d, device="cuda", generator=g).to(BF)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
out, lse = ops.attn_fwd(q, k, v)
do = torch.randn_like(out)
dqkv = torch.empty_like(qkv)
def bwd(): ops.attn_bwd(do, q, k, v, out, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
def fwd(): ops.attn_fwd(q, k, v)
for name, fn, fl in (("fwd", fwd, 4 * B * nh * S * S * d / 2), ("bwd", bwd, 2.5 * 4 * B * nh * S * S * d / 2)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms*1e3:.0f} us  {fl/ms/1e9:.0f} TF/s (algorithmic, causal)")
