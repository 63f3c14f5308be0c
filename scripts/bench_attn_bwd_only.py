import os, sys, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
BF = torch.bfloat16
B, S, nh, d = int(os.environ.get("ATTN_B", 8)), int(os.environ.get("ATTN_S", 2048)), int(os.environ.get("ATTN_NH", 32)), int(os.environ.get("ATTN_D", 128))
causal = os.environ.get("ATTN_CAUSAL", "1") == "1"
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, S, 3, nh, d, device="cuda", generator=g).to(BF)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
out, lse = ops.attn_fwd(q, k, v, causal=causal)
do = torch.randn_like(out)
dqkv = torch.empty_like(qkv)
for _ in range(4):
    ops.attn_fwd(q, k, v, causal=causal)
    ops.attn_bwd(do, q, k, v, out, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], causal=causal)
torch.cuda.synchronize()
