"""C4 (BASELINE.json configs[3]): SD-2.1 UNet, 64x64 latents, 50-step DDIM, 77 dream-query embeddings, bs 16, CUDA-graph loop.
Secondary benchmark (bench.py stays on configs[1]); prints one JSON line."""
import json, sys, time, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
from dreamllm_b200.modeling_plugins import StableDiffusionHead
BF = torch.bfloat16
B, Q, steps = 16, 77, 50
guidance = float(sys.argv[1]) if len(sys.argv) > 1 else 7.5
torch.manual_seed(0)
old = torch.get_default_dtype(); torch.set_default_dtype(BF)
with torch.device("cuda"):
    head = StableDiffusionHead(None)
torch.set_default_dtype(old)
g = torch.Generator(device="cuda").manual_seed(1)
pos = torch.randn(B, Q, 4096, device="cuda", generator=g).to(BF)
neg = torch.randn(B, Q, 4096, device="cuda", generator=g).to(BF)
from dreamllm_b200.unet import DenoiseLoop
cond = head.projector(pos)[-1]
if guidance > 1: cond = torch.cat([head.projector(neg)[-1], cond])
loop = DenoiseLoop(head.unet, cond, steps, guidance, "ddim", height=512, width=512)
loop.run(); torch.cuda.synchronize()          # captures the graph + first full run (warm-up)
ops.LAUNCHES.reset()
times = []
for it in range(3):
    loop.step.zero_(); loop.latents.normal_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop.run(); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = sorted(times)[1]
samples = (2 * B if guidance > 1 else B)
flop = 804.3e9 * samples * steps
print(json.dumps({"workload": f"SD-2.1 UNet 64x64, {steps}-step DDIM, Q={Q}, bs={B}, guidance={guidance} ({samples} UNet samples/step), CUDA-graph loop",
                  "ms_total": ms, "ms_per_step": ms / steps, "images_per_s": B / (ms / 1e3), "pixels_per_s": B * 512 * 512 / (ms / 1e3),
                  "algorithmic_pflop": flop / 1e15, "achieved_tflops": flop / 1e12 / (ms / 1e3), "frac_of_sustained_bf16_peak": flop / 1e12 / (ms / 1e3) / 1442.3,
                  "finite": bool(torch.isfinite(loop.latents).all())}))
