"""The kernel the reference would call on this box: flash-attn 2.8.x `flash_attn_func` (sm_100 SASS of the mma.sync
algorithm), timed on the C2 attention shape next to ours.  One JSON line.  Library code = the bar to beat, never on our product path."""
import json
import sys

import torch

sys.path.insert(0, ".")
BF = torch.bfloat16
B, S, nh, d = 8, 2048, 32, 128
FWD = 4 * B * nh * S * S * d / 2          # causal-counted


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
    out = {"shape": [B, S, nh, d], "causal": True}
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B, S, 3, nh, d, device="cuda", generator=g).to(BF)
    try:
        from flash_attn import flash_attn_func
        q, k, v = (qkv[:, :, i].contiguous().requires_grad_(True) for i in range(3))
        o = flash_attn_func(q, k, v, causal=True)
        do = torch.randn_like(o)
        t_f = timed(lambda: flash_attn_func(q, k, v, causal=True))

        def fb():
            oo = flash_attn_func(q, k, v, causal=True)
            oo.backward(do)
        t_fb = timed(fb)
        out["flash_attn_2"] = {"fwd_us": t_f * 1e3, "fwd_tflops": FWD / t_f / 1e9, "bwd_us": (t_fb - t_f) * 1e3,
                               "bwd_tflops": 2.5 * FWD / (t_fb - t_f) / 1e9}
    except Exception as ex:  # noqa: BLE001
        out["flash_attn_2"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    from dreamllm_b200 import ops
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attn_fwd(q, k, v)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    t_f = timed(lambda: ops.attn_fwd(q, k, v))
    t_b = timed(lambda: ops.attn_bwd(do, q, k, v, o, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]))
    out["ours"] = {"fwd_us": t_f * 1e3, "fwd_tflops": FWD / t_f / 1e9, "bwd_us": t_b * 1e3, "bwd_tflops": 2.5 * FWD / t_b / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
