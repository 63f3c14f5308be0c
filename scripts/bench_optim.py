"""Microbench of the optimizer-shard kernels (csrc/optim.cu): achieved HBM GB/s vs MEASURED_PEAKS.json.  Algorithmic bytes/element:
adamw fp32-state 28 B (14 read + 14 written), adamw bf16-state 14 B (8 + 6), sumsq 2 B.  Buffers are far larger than the 126 MB L2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamllm_b200 import ops  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(os.environ.get("N", 1 << 28))
    dev = "cuda"
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    g = (torch.randn(n, device=dev) * 0.01).to(torch.bfloat16)
    p = (torch.randn(n, device=dev) * 0.02).to(torch.bfloat16)
    hp = dict(lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0)
    ss = torch.ones(1, device=dev)
    out = {"n": n, "hbm_peak_gbps": hbm, "kernels": {}}
    m, v, w = torch.zeros(n, device=dev), torch.zeros(n, device=dev), p.float()
    step = [0]

    def f32():
        step[0] += 1
        ops.adamw_step_(g, p, m, v, w, step=step[0], grad_sumsq=ss, max_grad_norm=1.0, **hp)
    ms = timed(f32)
    out["kernels"]["adamw_fp32_state"] = {"ms": ms, "bytes": 28 * n, "gbps": 28 * n / ms / 1e6, "frac": 28 * n / ms / 1e6 / hbm}
    del m, v, w
    mb, vb = torch.zeros(n, device=dev, dtype=torch.bfloat16), torch.zeros(n, device=dev, dtype=torch.bfloat16)

    def b16():
        step[0] += 1
        ops.adamw_step_(g, p, mb, vb, None, step=step[0], grad_sumsq=ss, max_grad_norm=1.0, **hp)
    ms = timed(b16)
    out["kernels"]["adamw_bf16_state"] = {"ms": ms, "bytes": 14 * n, "gbps": 14 * n / ms / 1e6, "frac": 14 * n / ms / 1e6 / hbm}
    acc = torch.zeros(1, device=dev)
    ms = timed(lambda: ops.sumsq_bf16_(g, acc, accumulate=False))
    out["kernels"]["sumsq_bf16"] = {"ms": ms, "bytes": 2 * n, "gbps": 2 * n / ms / 1e6, "frac": 2 * n / ms / 1e6 / hbm}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
