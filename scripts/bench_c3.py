"""C3 (BASELINE.json configs[2]): Vicuna-7B + CLIP ViT-L/14-336 tower + linear projector, interleaved 1 image (576 visual tokens) +
1024 text tokens per sample, bs 4 / GPU, fwd + bwd (LLM + projector trainable, CLIP frozen — the stage-2 comprehension setting).
Secondary benchmark (bench.py stays on configs[1]).  The batch comes from the index-map collator (dreamllm_b200/collator.py): pinned host
ids / images -> async H2D -> `model(**batch)`; no `torch.where` / `.cpu()` sync inside the step.  Prints one JSON line; metric =
sum(attention_mask) / s.   NOTE: written after round 1's GPU budget was spent — not yet run on hardware."""
import json
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamllm_b200 import ops  # noqa: E402
from dreamllm_b200.clip_vision import CLIPVisionConfigLite  # noqa: E402
from dreamllm_b200.collator import DataCollatorForDreamLLMDataset, to_device  # noqa: E402
from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM  # noqa: E402
from dreamllm_b200.modeling_plugins import CLIPVisionEmbedding  # noqa: E402

BF = torch.bfloat16
IM_START, IM_PATCH, IM_END = 32003, 32002, 32004


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    warm, B, TXT, R = 3, 4, 1024, 336
    layers = int(os.environ.get("LAYERS", 32))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(BF)
    with torch.device(dev):
        model = DreamLLMForCausalMLM(DreamLLMConfig.vicuna_7b(num_hidden_layers=layers))
        clip = CLIPVisionEmbedding(CLIPVisionConfigLite(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                                                        image_size=R, patch_size=14), projector_type="linear", embed_hidden_size=4096)
    torch.set_default_dtype(old)
    model.model.attach_plugins(clip, None, image_start_id=IM_START, dream_start_id=32006)
    model.train()
    P = clip.embed_len
    assert P == 576
    collate = DataCollatorForDreamLLMDataset(SimpleNamespace(pad_token_id=32000), image_start_id=IM_START, clip_embed_len=P, pin_memory=True)
    g = torch.Generator().manual_seed(1234)
    examples = []
    for _ in range(B):
        text = torch.randint(3, 32000, (TXT,), generator=g).tolist()
        ids = torch.tensor([1, IM_START] + [IM_PATCH] * P + [IM_END] + text + [2])
        labels = ids.clone()
        labels[: P + 3] = -100                                             # image positions carry no LM loss (builder_dreamllm.py:197-200)
        examples.append(dict(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels,
                             images=torch.randn(1, 3, R, R, generator=g).to(BF), images_dm=None))
    host = collate(examples)
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    tokens = host["num_tokens"]
    keep = ("input_ids", "images", "attention_mask", "labels", "input_ids_cpu", "splice_plan", "attention_mask_has_padding", "seqlens",
            "shifted_labels")

    def step():
        for p in model.parameters():
            p.grad = None
        batch = to_device({k: host[k] for k in keep}, dev)                 # H2D every step (ids, images, index maps)
        out = model(**batch)
        out.loss.backward()
        return out.loss

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    ops.LAUNCHES.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = float(step().item())                                        # D2H of the loss every step
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"metric": "interleaved tokens/s (C3: CLIP ViT-L/14-336 + linear projector + Vicuna-7B fwd+bwd)", "value": tokens / ms * 1e3,
                      "unit": "tokens/s", "ms_per_step": ms, "n_gpus": 1, "steps": steps, "warmup": warm, "dtype": "bf16",
                      "config": {"workload": "BASELINE.json configs[2]", "layers": layers, "bs": B, "seq_len": int(host["input_ids"].shape[1]),
                                 "visual_tokens": P, "text_tokens": TXT}, "loss": loss, "gpu_launches": ops.LAUNCHES.count,
                      "data": "synthetic"}))


if __name__ == "__main__":
    main()
