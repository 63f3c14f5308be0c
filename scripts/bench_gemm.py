"""Dev microbenchmark: tcgen05 GEMM vs torch.matmul (cuBLAS) on the decoder shapes. Not the judged bench."""
import sys, json, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

T = 16384
res = []
for (M, N, K, a_mn, b_mn, name) in [
    (T, 12288, 4096, 0, 0, "qkv fwd"), (T, 4096, 4096, 0, 0, "o fwd"), (T, 22016, 4096, 0, 0, "gate_up fwd"),
    (T, 4096, 11008, 0, 0, "down fwd"), (T, 4096, 12288, 0, 1, "qkv dgrad"), (T, 4096, 22016, 0, 1, "gate_up dgrad"),
    (T, 11008, 4096, 0, 1, "down dgrad"), (12288, 4096, T, 1, 1, "qkv wgrad"), (22016, 4096, T, 1, 1, "gate_up wgrad"),
    (4096, 11008, T, 1, 1, "down wgrad"), (T, 32008, 4096, 0, 0, "lm_head fwd"), (512, 4096, 4096, 0, 0, "C1 o fwd"),
]:
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    A = a.t() if a_mn else a
    Bm = b if b_mn else b.t()
    row = {"name": name, "M": M, "N": N, "K": K}
    for cp in (0, 1):
        try:
            ms = timeit(lambda: ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, cta_pair=cp))
            row[f"ours_cta{cp+1}_tflops"] = round(2 * M * N * K / ms / 1e9, 1)
        except Exception as ex:
            row[f"ours_cta{cp+1}_err"] = str(ex)[:80]
    ms = timeit(lambda: torch.matmul(A, Bm, out=out))
    row["cublas_tflops"] = round(2 * M * N * K / ms / 1e9, 1)
    print(json.dumps(row), flush=True)
    res.append(row)
json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)
