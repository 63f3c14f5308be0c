"""Per-tile timeline of one forward-attention CTA (dev tool; needs a -DDLLM_ATTN_TRACE build of the library):

    DLLM_LIB_PATH=$PWD/dreamllm_b200/libdreamllm_sm100_trace.so DLLM_NVCC_EXTRA=-DDLLM_ATTN_TRACE python scripts/attn_trace.py

Prints, for the heaviest q tile of head 1 / batch 0 at the C2 shape, clock64 deltas (relative to the CTA's start) of: softmax warp 0
{S ready, S in registers, max + vote done, exp done, P stored + arrive}, the MMA thread {QK(j) issued, V(j) ready, P(j) seen} and the TMA
thread {V(j) stage wait begin / end}."""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
from dreamllm_b200 import _lib, ops  # noqa: E402

_lib.build()
B, S, nh, d = 4, 2048, 32, 128
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, S, 3, nh, d, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
for _ in range(3):
    o, lse = ops.attn_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal=True)
torch.cuda.synchronize()
L = _lib.lib()
n = 64 * 16 + 16
buf = (ctypes.c_longlong * n)()
rc = L.dllm_attn_trace_read(buf, n)
assert rc == 0, rc
t0 = buf[64 * 16]
rows = []
for j in range(32):
    r = [buf[j * 16 + k] - t0 for k in range(11)]
    rows.append(r)
names = ["s_ready", "s_loaded", "max_done", "exp_done", "p_arrived", "mma_qk_issue", "mma_v_ready", "mma_p_seen", "tma_v_wait0", "tma_v_wait1", "sm_loop_top"]
out = {"names": names, "rows": rows, "end_pv_done": buf[64 * 16 + 1] - t0, "end_stored": buf[64 * 16 + 2] - t0}
print(json.dumps(out))
sys.stderr.write("j   " + " ".join(f"{x:>12}" for x in names) + "\n")
for j, r in enumerate(rows):
    sys.stderr.write(f"{j:<3} " + " ".join(f"{x:>12}" for x in r) + "\n")
sys.stderr.write(f"pv_done(last) {out['end_pv_done']}  stored {out['end_stored']}\n")
