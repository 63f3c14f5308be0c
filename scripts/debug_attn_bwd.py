import sys, math, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
sys.path.insert(0, "tests"); from test_attn_gpu import _ref_core
BF = torch.bfloat16
for (B, S, nh, d, causal, seqlens) in [(1, 320, 1, 128, True, None), (1, 256, 1, 64, True, None), (1, 320, 1, 64, True, None),
                                       (2, 320, 2, 64, True, [320, 129]), (1, 192, 1, 128, True, None), (1, 64, 1, 128, True, None)]:
    g = torch.Generator().manual_seed(100 + S)
    qkv = torch.randn(B, S, 3, nh, d, generator=g).to(BF)
    dout = (torch.randn(B, S, nh * d, generator=g) * 0.5).to(BF)
    sl = torch.tensor(seqlens) if seqlens is not None else None
    valid = torch.ones(B, S, dtype=torch.bool) if sl is None else (torch.arange(S)[None] < sl[:, None])
    dout = dout * valid[..., None]
    q32, k32, v32 = (qkv[:, :, i].float().requires_grad_(True) for i in range(3))
    ref = _ref_core(q32, k32, v32, causal, sl)
    ref.backward(dout.float())
    dev = qkv.cuda()
    q, k, v = dev[:, :, 0], dev[:, :, 1], dev[:, :, 2]
    sl_dev = sl.int().cuda() if sl is not None else None
    out, lse = ops.attn_fwd(q, k, v, causal=causal, seqlens=sl_dev)
    dqkv = torch.zeros_like(dev)
    try:
        ops.attn_bwd(dout.cuda(), q, k, v, out, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], causal=causal, seqlens=sl_dev)
        torch.cuda.synchronize()
    except Exception as e:
        print("EXC", e); break
    got = dqkv.cpu().float()
    print(f"== B{B} S{S} nh{nh} d{d} causal{causal} sl{seqlens}")
    for i, (name, rg) in enumerate((("dq", q32.grad), ("dk", k32.grad), ("dv", v32.grad))):
        err = (got[:, :, i] - rg).abs().amax(dim=(2, 3))          # [B, S]
        per = err.view(B, -1, 32).amax(-1) if S % 32 == 0 else err
        print(name, "scale", float(rg.abs().max()), "err per 32-row block:", [[round(float(x), 3) for x in r] for r in per])
