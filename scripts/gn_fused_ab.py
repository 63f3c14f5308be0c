"""Fused (one cooperative launch) vs three-launch GroupNorm forward: outputs and statistics must be bit-identical.  Run twice:
`python scripts/gn_fused_ab.py save f.pt` with DLLM_GN_NO_FUSE=1, then `python scripts/gn_fused_ab.py cmp f.pt` without."""
import sys
import torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops  # noqa: E402

mode, path = sys.argv[1], sys.argv[2]
res = {}
for (N, HW, C, silu) in [(4, 4096, 320, True), (4, 1024, 640, True), (4, 256, 1280, False), (4, 64, 1280, True), (2, 4096, 960, True),
                         (1, 1024, 1920, True), (4, 65536, 128, True), (3, 300, 2560, True), (32, 4096, 320, True), (2, 16, 128, False)]:
    g = torch.Generator(device="cuda").manual_seed(N * 7 + HW + C)
    x = (torch.randn(N, HW, C, device="cuda", generator=g) * 1.3 + 0.2).to(torch.bfloat16)
    w = torch.randn(C, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(C, device="cuda", generator=g).to(torch.bfloat16)
    for rep in range(2):
        y, st = ops.groupnorm(x, w, b, 32, 1e-5, silu, return_stats=True)
        y2 = ops.groupnorm(x, w, b, 32, 1e-5, silu)
        torch.cuda.synchronize()
        assert torch.equal(y, y2)
        res[(N, HW, C, silu, rep)] = (y.cpu(), st.cpu())
    # timing
    for _ in range(3):
        ops.groupnorm(x, w, b, 32, 1e-5, silu)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.groupnorm(x, w, b, 32, 1e-5, silu)
    e1.record()
    torch.cuda.synchronize()
    print(f"N={N} HW={HW} C={C}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
if mode == "save":
    torch.save(res, path)
else:
    ref = torch.load(path)
    bad = 0
    for k, (y, st) in res.items():
        if not torch.equal(y, ref[k][0]) or not torch.equal(st, ref[k][1]):
            bad += 1
            print("MISMATCH", k, float((y.float() - ref[k][0].float()).abs().max()), float((st - ref[k][1]).abs().max()))
    print("compared", len(res), "bad", bad)
    sys.exit(1 if bad else 0)
