"""Per-iteration timeline of one dK/dV backward CTA (kv tile 0 of head 1 / batch 0: 32 q tiles); needs the -DDLLM_ATTN_TRACE build
(see scripts/attn_trace.py).  Columns: row warp 0 {S^T,dP^T ready, in registers, P^T/dS^T computed, stored + arrive, loop top},
MMA thread {Q/dO(it+1) ready -> S/dP issue, before / after the P^T,dS^T wait}, TMA thread {ring-slot wait begin / end}."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from dreamllm_b200 import _lib, ops  # noqa: E402

_lib.build()
B, S, nh, d = 4, 2048, 32, 128
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, S, 3, nh, d, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
o, lse = ops.attn_fwd(q, k, v, causal=True)
do = torch.randn(B, S, nh * d, device="cuda", generator=g).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    ops.attn_bwd(do, q, k, v, o, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], causal=True)
torch.cuda.synchronize()
L = _lib.lib()
n = 64 * 16 + 16
buf = (ctypes.c_longlong * n)()
assert L.dllm_attn_trace_read(buf, n) == 0
t0 = buf[64 * 16 + 3]
names = ["sdp_ready", "loaded", "-", "computed", "arrived", "mma_qdo_ready", "mma_pre_pds", "mma_pds_seen", "tma_wait0", "tma_wait1", "row_top", "mma_pre_qdo"]
print("it  " + " ".join(f"{x:>13}" for x in names))
for it in range(32):
    print(f"{it:<3} " + " ".join(f"{buf[it * 16 + c] - t0:>13}" for c in range(12)))
