"""A/B of the GEMM / implicit-conv N-tile width on the UNet (32 samples) and VAE (4 images) shapes: run once per DLLM_GEMM_BN value.
Prints one JSON line {shape: us}."""
import json, os, sys, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.1).to(BF)
def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
out = {"bn": os.environ.get("DLLM_GEMM_BN", "auto")}
for name, (N, H, W, Ci, Co) in {"conv 32x64x64 320->320": (32, 64, 64, 320, 320), "conv 32x32x32 640->640": (32, 32, 32, 640, 640),
                                "conv 32x64x64 640->320": (32, 64, 64, 640, 320), "conv 32x64x64 960->320": (32, 64, 64, 960, 320),
                                "conv 4x512x512 128->128": (4, 512, 512, 128, 128), "conv 4x256x256 256->256": (4, 256, 256, 256, 256),
                                "conv 4x64x64 320->320": (4, 64, 64, 320, 320)}.items():
    x, w, b = r(N, H, W, Ci), r(Co, 9 * Ci), r(Co)
    us = timed(lambda: ops.conv3x3(x, w, bias=b))
    out[name] = {"us": round(us, 1), "tflops": round(2 * N * H * W * Co * 9 * Ci / us / 1e6, 1)}
    del x, w
for name, (M, N, K) in {"linear 131072 x 320 x 320": (131072, 320, 320), "linear 131072 x 960 x 320": (131072, 960, 320),
                        "linear 131072 x 2560 x 320": (131072, 2560, 320), "linear 131072 x 320 x 1280": (131072, 320, 1280),
                        "linear 32768 x 640 x 640": (32768, 640, 640), "linear 32768 x 640 x 2560": (32768, 640, 2560),
                        "linear 16384 x 320 x 320": (16384, 320, 320)}.items():
    x, w, b = r(M, K), r(N, K), r(N)
    us = timed(lambda: ops.linear(x, w, bias=b))
    out[name] = {"us": round(us, 1), "tflops": round(2 * M * N * K / us / 1e6, 1)}
print(json.dumps(out))
