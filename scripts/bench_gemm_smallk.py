import sys, json, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
for (M, N, K, bias, res, name) in [(131072, 2560, 320, 1, 0, "unet ff.proj 64x64"), (131072, 320, 1280, 1, 1, "unet ff.out 64x64"), (131072, 960, 320, 0, 0, "unet qkv 64x64"),
                                   (32768, 5120, 640, 1, 0, "unet ff.proj 32x32"), (18464, 3072, 1024, 1, 0, "clip qkv 32 img"), (18464, 4096, 1024, 1, 0, "clip fc1"),
                                   (16384, 12288, 4096, 0, 0, "llm qkv"), (16384, 4096, 11008, 0, 1, "llm down+res")]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bi = torch.randn(N, device="cuda").to(torch.bfloat16) if bias else None
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.gemm(a, b, out=out, bias=bi, residual=r))
    print(f"{name:24s} M{M} N{N} K{K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.0f} TF/s  ({(M*K+N*K+M*N*(2 if res else 1))*2/ms/1e6:6.0f} GB/s algorithmic)")
