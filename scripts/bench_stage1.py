"""C5 (BASELINE.json configs[4]): DreamLLM stage-1 *creation* step — Vicuna-7B LLM (frozen) + dream queries (trainable) + SD-2.1 head
(VAE encode + UNet, frozen; projector trainable) — 4 samples / GPU, 512x512 target images, fwd + bwd (dgrad through all 32 LLM
layers and the whole UNet).  Secondary benchmark; run under torchrun for N > 1 (DDP all-reduce of the ~4.5 M trainable params).
Prints one JSON line.  Metric = (sum(attention_mask) + Nd*512*512) / s  (SURVEY.md §8d)."""
import json, os, sys, torch
sys.path.insert(0, ".")
from dreamllm_b200 import ops
from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
from dreamllm_b200.modeling_plugins import DreamEmbedding, StableDiffusionHead
BF = torch.bfloat16
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
B, Q, TXT, steps, warm = 4, 64, 32, int(sys.argv[1]) if len(sys.argv) > 1 else 5, 2
torch.manual_seed(0)
old = torch.get_default_dtype(); torch.set_default_dtype(BF)
with torch.device(dev):
    m = DreamLLMForCausalMLM(DreamLLMConfig.vicuna_7b())
    dream = DreamEmbedding(num_dream_queries=Q, embed_hidden_size=4096)
    m.stable_diffusion_head = StableDiffusionHead(None, embed_hidden_size=4096)
torch.set_default_dtype(old)
m.model.attach_plugins(None, dream, image_start_id=32003, dream_start_id=32006)
for p in m.parameters(): p.requires_grad_(False)
dream.dream_queries.requires_grad_(True)
m.stable_diffusion_head.projector.requires_grad_(True)
m.train()
trainable = [p for p in m.parameters() if p.requires_grad]
reducer = None
if world > 1:
    from dreamllm_b200.ddp import BucketedGradReducer
    reducer = BucketedGradReducer(trainable, bucket_cap_mb=64.0)
g = torch.Generator().manual_seed(1 + rank)
S = 1 + TXT + 1 + Q + 1 + 1
ids = torch.empty(B, S, dtype=torch.long)
for b in range(B):
    ids[b] = torch.tensor([1] + torch.randint(3, 32000, (TXT,), generator=g).tolist() + [32006] + [32002] * Q + [32007, 2])
labels = torch.full((B, S), -100)
imgs = (torch.rand(B, 3, 512, 512, generator=g) * 2 - 1).pin_memory()
ids_pin = ids.pin_memory()

from dreamllm_b200.modeling_plugins import build_splice_plan
plan = build_splice_plan(ids, -1, 32006, 0, Q, 0, B, dev)          # static layout: index maps built once
x_static = torch.empty((B, S), dtype=torch.long, device=dev)
im_static = torch.empty((B, 3, 512, 512), dtype=torch.float32, device=dev)
lab_dev = labels.to(dev)
USE_GRAPH = os.environ.get("DLLM_STAGE1_GRAPH", "1") == "1" and world == 1

def compute():
    for p in trainable: p.grad = None
    out = m(input_ids=x_static, images_dm=im_static.to(BF), labels=lab_dev, attention_mask_has_padding=False, splice_plan=plan)
    out.loss.backward()
    return out

graph = None
static_out = None
def step():
    global graph, static_out
    if reducer: reducer.zero_grad()
    x_static.copy_(ids_pin, non_blocking=True); im_static.copy_(imgs, non_blocking=True)      # H2D every step
    if USE_GRAPH and graph is not None:
        graph.replay()
        out = static_out
    else:
        out = compute()
    if reducer: reducer.finalize()
    return out

for _ in range(warm): step()
torch.cuda.synchronize()
graph_note = "eager launches"
if USE_GRAPH:
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            compute()
        torch.cuda.current_stream().wait_stream(s)
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            static_out = compute()
        graph = g_
        graph_note = "whole fwd+bwd step captured in ONE CUDA graph"
        step(); torch.cuda.synchronize()
    except Exception as ex:
        graph = None; USE_GRAPH = False
        graph_note = f"graph capture failed ({type(ex).__name__}: {str(ex)[:120]}); eager launches"
        torch.cuda.synchronize()
if world > 1: dist.barrier()
ops.LAUNCHES.reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): out = step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(ms)
    toks, pix = B * S * world, B * 512 * 512 * world
    print(json.dumps({"workload": f"stage-1 creation step: Vicuna-7B (frozen) + {Q} dream queries + SD-2.1 VAE-enc/UNet (frozen), {B} samples/GPU x {world} GPU, "
                      f"seq {S}, 512x512 targets, fwd+bwd incl. H2D of ids/images", "ms_per_step": ms, "tokens_per_s": toks / ms * 1e3, "pixels_per_s": pix / ms * 1e3,
                      "tokens_plus_pixels_per_s": (toks + pix) / ms * 1e3, "launches_per_step": ops.LAUNCHES.count / steps,
                      "vm_loss": float(out.additional_log_info["vm_loss"]), "n_gpus": world, "launch_mode": graph_note,
                      "dq_grad_norm": float(dream.dream_queries.grad.float().norm())}))
if world > 1: dist.destroy_process_group()
