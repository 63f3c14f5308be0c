"""Two eager UNet denoise steps at the C4 shape (32 samples = bs 16 x CFG) for an ncu launch list (`--metrics gpu__time_duration.sum`)."""
import sys, torch
sys.path.insert(0, ".")
from dreamllm_b200.modeling_plugins import StableDiffusionHead
from dreamllm_b200.unet import DenoiseLoop
BF = torch.bfloat16
torch.manual_seed(0)
old = torch.get_default_dtype(); torch.set_default_dtype(BF)
with torch.device("cuda"):
    head = StableDiffusionHead(None)
torch.set_default_dtype(old)
g = torch.Generator(device="cuda").manual_seed(1)
cond = torch.randn(32, 77, 1024, device="cuda", generator=g).to(BF)
loop = DenoiseLoop(head.unet, cond, 50, 7.5, "ddim", height=512, width=512, use_cuda_graph=False)
torch.cuda.synchronize()
print("MARK: model built")
for _ in range(2):
    loop._one_step()
torch.cuda.synchronize()
print("done")
