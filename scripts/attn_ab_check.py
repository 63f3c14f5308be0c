"""Bit-level A/B of the attention kernels: run every shape in this process (persistent kernels unless DLLM_ATTN_NONPERSIST=1) and either
save the outputs (`save <file>`) or compare them with a saved file (`cmp <file>`).  Both variants do the same arithmetic in the same order,
so any difference is a synchronisation bug.  Repeats each shape 3x to catch races."""
import sys
import torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops  # noqa: E402

BF = torch.bfloat16
SHAPES = [  # (B, S, Skv, nh, d, causal, seqlens)
    (2, 256, 256, 5, 64, False, None), (2, 64, 64, 10, 64, False, None), (2, 16, 16, 20, 64, False, None), (2, 4, 4, 20, 64, False, None),
    (2, 256, 7, 5, 64, False, None), (2, 64, 7, 10, 64, False, None), (2, 16, 7, 20, 64, False, None), (2, 4, 7, 20, 64, False, None),
    (1, 1024, 64, 5, 64, False, None), (3, 1024, 77, 5, 64, False, None), (2, 1024, 1024, 10, 64, False, None),
    (2, 384, 384, 4, 128, True, None), (2, 200, 200, 2, 128, True, [200, 131]), (3, 577, 577, 16, 64, False, None),
    (4, 2048, 2048, 32, 128, True, None), (1, 100, 100, 32, 128, True, None), (2, 512, 512, 32, 128, True, [512, 300]),
]
mode, path = sys.argv[1], sys.argv[2]
if len(sys.argv) > 3:
    SHAPES = [sh for sh in SHAPES if sh[1] >= int(sys.argv[3])]
res = {}
for (B, S, Skv, nh, d, causal, sl) in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + S + Skv)
    r = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.7).to(BF)  # noqa: E731
    q, k, v, do = r(B, S, nh, d), r(B, Skv, nh, d), r(B, Skv, nh, d), r(B, S, nh * d)
    seql = torch.tensor(sl, device="cuda", dtype=torch.int32) if sl else None
    for rep in range(3):
        if S == Skv:
            o, lse = ops.attn_fwd(q, k, v, causal=causal, seqlens=seql)
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            ops.attn_bwd(do, q, k, v, o, lse, dq, dk, dv, causal=causal, seqlens=seql)
        else:
            o, lse = ops.attn_fwd_cross_lse(q, k, v)
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            ops.attn_bwd_cross(do, q, k, v, o, lse, dq, dk, dv)
        torch.cuda.synchronize()
        res[(B, S, Skv, nh, d, causal, rep)] = [t.cpu() for t in (o, lse, dq, dk, dv)]
if mode == "save":
    torch.save(res, path)
    print("saved", len(res))
else:
    ref = torch.load(path)
    bad = 0
    for key, ts in res.items():
        for name, a, b in zip(("o", "lse", "dq", "dk", "dv"), ts, ref[key]):
            a32, b32 = a.float(), b.float()
            fin = torch.isfinite(b32)
            if not torch.equal(torch.isfinite(a32), fin) or not torch.equal(a32[fin], b32[fin]):
                bad += 1
                diff = (a32[fin] - b32[fin]).abs()
                print("MISMATCH", key, name, "max abs diff", float(diff.max()) if diff.numel() else "nan-pattern", "n", int((diff > 0).sum()))
                if name in ("dk", "dv", "dq") and a32.dim() == 4:
                    B_, S_, nh_, d_ = a32.shape
                    bad_el = (a32 != b32)
                    # boxes: [b, s // 32, h, d // 64]
                    bx = bad_el.view(B_, S_ // 32, 32, nh_, d_ // 64, 64).any(dim=5).any(dim=2)
                    frac = bad_el.view(B_, S_ // 32, 32, nh_, d_ // 64, 64).float().mean(dim=(2, 5))
                    idx = bx.nonzero()
                    print("   bad [32x64] boxes:", idx.shape[0], " fully-bad:", int((frac > 0.9).sum()), " first:", idx[:12].tolist())
                    tiles = bx.view(B_, S_ // 128, 4, nh_, d_ // 64)
                    print("   per-tile box counts (b, kvtile, h):", sorted({(int(i[0]), int(i[1]) // 4, int(i[2])) for i in idx})[:16])
                    print("   wq histogram:", [int(tiles[:, :, w].sum()) for w in range(4)], " chunk histogram:", [int(tiles[..., c].sum()) for c in range(d_ // 64)])
    print("compared", len(res), "bad", bad)
    sys.exit(1 if bad else 0)
