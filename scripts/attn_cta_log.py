"""SM-level picture of the forward-attention grid (dev tool, -DDLLM_ATTN_TRACE build): every CTA logs {start, end, smid, n_kv}; prints per-SM
busy fraction, the per-CTA overhead (duration - n_kv * t_tile fit) and clocks per KV tile as a function of n_kv."""
import ctypes
import json
import sys

import numpy as np
import torch

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
n = 4096 * 4
buf = (ctypes.c_longlong * n)()
assert L.dllm_attn_cta_log_read(buf, n) == 0
a = np.array(buf, dtype=np.int64).reshape(4096, 4)
t0, t1, sm, nkv = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
dur = t1 - t0
out = {}
# linear fit duration = ovh + t_tile * n_kv
A = np.stack([np.ones_like(nkv), nkv], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, dur.astype(np.float64), rcond=None)
out["fit_overhead_clk"], out["fit_clk_per_tile_per_cta"] = float(coef[0]), float(coef[1])
out["by_nkv"] = {int(k): {"mean_clk": float(dur[nkv == k].mean()), "clk_per_tile": float(dur[nkv == k].mean() / k)} for k in sorted(set(nkv.tolist()))}
spans, busy2, busy1, idle = [], [], [], []
for s_ in sorted(set(sm.tolist())):
    m = sm == s_
    ev = sorted([(int(x), 1) for x in t0[m]] + [(int(x), -1) for x in t1[m]])
    lo, hi = ev[0][0], ev[-1][0]
    lvl, last, acc = 0, lo, {0: 0, 1: 0, 2: 0}
    for t, dlt in ev:
        acc[min(lvl, 2)] += t - last
        last = t
        lvl += dlt
    spans.append(hi - lo); busy2.append(acc[2]); busy1.append(acc[1]); idle.append(acc[0])
out["sm_span_clk"] = {"mean": float(np.mean(spans)), "max": int(np.max(spans)), "min": int(np.min(spans))}
out["frac_two_ctas"] = float(np.sum(busy2) / np.sum(spans))
out["frac_one_cta"] = float(np.sum(busy1) / np.sum(spans))
out["frac_idle_inside_span"] = float(np.sum(idle) / np.sum(spans))
out["ctas_per_sm"] = {"mean": 4096 / len(spans), "max": int(max(np.sum(sm == s_) for s_ in set(sm.tolist())))}
out["tiles_per_sm"] = {"mean": float(nkv.sum() / len(spans)), "max": int(max(nkv[sm == s_].sum() for s_ in set(sm.tolist()))),
                       "min": int(min(nkv[sm == s_].sum() for s_ in set(sm.tolist())))}
print(json.dumps(out, indent=1))
