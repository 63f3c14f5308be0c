"""Per-item timeline of persistent forward-attention CTA 0 (-DDLLM_ATTN_TRACE build)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from dreamllm_b200 import _lib, ops  # noqa: E402
_lib.build()
B, S, nh, d = 8, 2048, 32, 128
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, S, 3, nh, d, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
for _ in range(3):
    ops.attn_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal=True)
torch.cuda.synchronize()
L = _lib.lib()
n = 64 * 16 + 16
buf = (ctypes.c_longlong * n)()
assert L.dllm_attn_trace_read(buf, n) == 0
t0 = buf[64 * 16]
names = ["sm_item_top", "S0_ready", "lastP_done", "pv_done", "epi_done", "mma_qfull", "tma_qe_w0", "tma_qe_w1", "n_kv", "tma_item_end", "mma_pre_qfull", "S4_ready", "S8_ready"]
print("item " + " ".join(f"{x:>13}" for x in names))
for it in range(16):
    print(f"{it:<4} " + " ".join(f"{(buf[it * 16 + c] - (0 if c == 8 else t0)):>13}" for c in range(13)))
