"""Microbench of the HBM-bound kernels at the C2 (decoder, T = 16 384) and C4 (UNet, 32 samples) shapes: achieved GB/s against
MEASURED_PEAKS.json, one JSON line.  Built for the round-2 A/B of the `-DDLLM_VEC128 -DDLLM_GN_GROUP2` build flags:

    python scripts/bench_hbm_kernels.py                                           # shipped library
    DLLM_NVCC_EXTRA="-DDLLM_VEC128 -DDLLM_GN_GROUP2" DLLM_LIB_PATH=/tmp/ab.so python scripts/bench_hbm_kernels.py --build

Bytes = algorithmic traffic (tensor reads + writes x 2 B; fp32 side arrays ignored).  Inputs exceed L2 for the decoder shapes; the UNet
planes (<= 84 MB) partly fit L2, as in the real step.  NOTE: written after round 1's GPU budget was spent — not yet run on hardware."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamllm_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16


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
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    if "--build" in sys.argv:
        _lib.build(force=True)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).to(BF)
    out = {"lib": _lib.LIB_PATH, "nvcc_extra": os.environ.get("DLLM_NVCC_EXTRA", ""), "hbm_peak_gbps": peak, "kernels": {}}

    def rec(name, fn, nbytes):
        t = timed(fn)
        out["kernels"][name] = {"us": t * 1e6, "gbps": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / peak}

    # ---- decoder path, C2: T = 16384, H = 4096, I = 11008
    T, H, I, nh, d = 16384, 4096, 11008, 32, 128
    x, w, add = rnd(T, H), rnd(H), rnd(T, H)
    rec("rmsnorm_fwd", lambda: ops.rmsnorm_fwd(x, w, 1e-6), 2 * T * H * 2)
    rec("rmsnorm_fwd_add", lambda: ops.rmsnorm_fwd(x, w, 1e-6, add=add), 4 * T * H * 2)
    _, rstd, _ = ops.rmsnorm_fwd(x, w, 1e-6)
    dy = rnd(T, H)
    rec("rmsnorm_bwd(dx+dw)", lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dres=add), 6 * T * H * 2)        # dx pass 4 tensors + dw pass 2 reads
    qkv = rnd(T, 3 * H)
    cos = torch.randn(2048, d, device=dev, generator=g).to(BF)
    sin = torch.randn(2048, d, device=dev, generator=g).to(BF)
    pos = torch.arange(2048, device=dev, dtype=torch.int32).repeat(T // 2048)
    rec("rope", lambda: ops.rope_(qkv, cos, sin, pos, 2 * nh, d), 4 * T * H * 2)
    gu = rnd(T, 2 * I)
    rec("swiglu_fwd", lambda: ops.swiglu_fwd(gu, I), 3 * T * I * 2)
    dact = rnd(T, I)
    rec("swiglu_bwd", lambda: ops.swiglu_bwd(dact, gu, I), 5 * T * I * 2)
    del x, add, dy, qkv, gu, dact
    # ---- UNet, C4: 32 samples (bs 16 x CFG)
    for C, HW in ((320, 4096), (640, 1024), (1280, 256)):
        N = 32
        Tn = N * HW
        xl, wl, bl = rnd(Tn, C), rnd(C), rnd(C)
        rec(f"layernorm_fwd[{Tn}x{C}]", lambda: ops.layernorm_fwd(xl, wl, bl, 1e-5), 2 * Tn * C * 2)
        dyl = rnd(Tn, C)
        rec(f"layernorm_bwd[{Tn}x{C}]", lambda: ops.layernorm_bwd(dyl, xl, wl, 1e-5), 3 * Tn * C * 2)
        xg = xl.view(N, HW, C)
        rec(f"groupnorm_silu[{N}x{HW}x{C}]", lambda: ops.groupnorm(xg, wl, bl, 32, 1e-5, True), 3 * Tn * C * 2)   # stats pass + apply pass
        _, stats = ops.groupnorm(xg, wl, bl, 32, 1e-5, True, return_stats=True)
        dyg = dyl.view(N, HW, C)
        rec(f"groupnorm_silu_bwd[{N}x{HW}x{C}]", lambda: ops.groupnorm_bwd(dyg, xg, wl, bl, stats, 32, True), 5 * Tn * C * 2)
        ff = rnd(Tn, 8 * C)
        rec(f"geglu[{Tn}x{4 * C}]", lambda: ops.geglu(ff), 3 * Tn * 4 * C * 2)
        del xl, dyl, ff
    print(json.dumps(out))


if __name__ == "__main__":
    main()
