#!/usr/bin/env python
"""bench.py — the judged benchmark (contract in the task statement).

Workload (BASELINE.json configs[1]): Vicuna-7B-shaped DreamLLM text-only causal-LM fwd+bwd, seq 2048, bs 8 per GPU,
bf16, synthetic tokens, random-init weights.  Metric: interleaved tokens+pixels / s (text-only => pixels = 0).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1 under torchrun, weak scaling, DDP)
  python bench.py --impl reference ...                     # the reference algorithm on the host CPU cores (oracle port)

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "interleaved tokens+pixels/sec @ Vicuna-7B+SD2.1 512px, 1/2/4/8 B200"
UNIT = "tokens+pixels/s"

# Vicuna-7B DreamLLM (vocab 32000 + 8 special tokens, projects/dreamllm/train.py:74-89)
H, I, NH, L, V = 4096, 11008, 32, 32, 32008


def flops_per_step(bs, seq, layers=L, hidden=H, inter=I, vocab=V):
    """Algorithmic FLOPs of one fwd+bwd: causal attention counted at half, bwd = 2 x fwd."""
    T = bs * seq
    per_tok_layer = 2 * (4 * hidden * hidden + 3 * hidden * inter) + 2 * seq * hidden  # GEMMs + causal attn (2*S*H)
    fwd = T * (layers * per_tok_layer + 2 * hidden * vocab)
    return 3 * fwd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1442.3), d.get("hbm_gbs", 6569.6), "measured"
    return 1400.0, 6650.0, "fallback"


def gemm_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/*_gemm_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_gemm_traffic.json")))
    if not files:
        return None, "no ncu --set full capture committed"
    d = json.load(open(files[-1]))
    return d["dram_bytes_per_launch"], d["note"]


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(seq, layers_sample=1, iters=1, warm=1):
    """Reference algorithm (oracle port of modeling_dreamllm.py:599-654 + :1452-1470) on the host cores, bf16 (the config's
    dtype; AMX/AVX512-bf16 where present), bounded sample: `layers_sample` decoder layers + lm_head/CE at bs=1, seq tokens,
    fwd+bwd; extrapolated linearly to 32 layers."""
    from oracle import decoder_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(0)
    ps = []
    for li in range(layers_sample):
        p = {k: v.requires_grad_(True) for k, v in O.init_layer_params(H, I, 100 + li, dtype=dt).items()}
        ps.append(p)
    lm_w = (torch.randn(V, H, generator=g) * 0.02).to(dt).requires_grad_(True)
    norm_w = torch.ones(H, dtype=dt, requires_grad=True)
    x = torch.randn(1, seq, H, generator=g).to(dt).requires_grad_(True)
    labels = torch.randint(0, 32000, (1, seq), generator=g)
    cos, sin = O.rope_tables(H // NH, 2048, dtype=dt)
    pos = torch.arange(seq)[None]
    mask = O.causal_additive_mask(1, seq, dt)

    def layers_step():
        h = x
        for p in ps:
            h = O.decoder_layer(h, p, NH, cos, sin, pos, mask)
        h.float().pow(2).mean().backward()

    def head_step():
        hh = O.rmsnorm(x, norm_w)
        logits = torch.nn.functional.linear(hh, lm_w).float()
        O.lm_loss(logits, labels).backward()

    def t(fn):
        for _ in range(warm):
            fn()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        return (time.perf_counter() - t0) / iters

    t_layer = t(layers_step) / layers_sample
    t_head = t(head_step)
    t_full = t_layer * L + t_head
    return {"value": seq / t_full, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (CPU torch bf16, eager attention): {layers_sample} decoder layer(s) + lm_head/CE, bs=1 seq={seq}, "
                      f"fwd+bwd, {iters} iter; layer {t_layer:.2f}s x{L} + head {t_head:.2f}s extrapolated to the 32-layer model",
            "t_layer_s": round(t_layer, 3), "t_head_s": round(t_head, 3)}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cb = cpu_reference(args.seq, layers_sample=1, iters=max(1, min(args.steps, 2)), warm=1 if args.warmup else 0)
    tok_per_step = args.bs * args.seq
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tok_per_step / cb["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"BASELINE.json configs[1]: Vicuna-7B DreamLLM text-only causal-LM fwd+bwd, seq={args.seq} bs={args.bs}/GPU bf16",
            "layers": args.layers, "hidden": H, "vocab": V, "global_batch": args.bs * args.gpus, "seq_len": args.seq,
            "parallelism": f"dp{args.gpus}", "l2": "working set (13.5 GB weights + activations) >> 126 MB L2; no explicit flush",
            "optimizer_step": "none (config is fwd+bwd)"}


# ---------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=L, help="dev only: fewer layers (reported in config; not a valid bench)")
    ap.add_argument("--bs", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from dreamllm_b200 import ops
    from dreamllm_b200.ddp import BucketedGradReducer
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: dreamllm_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from dreamllm_b200.ddp import configure_nccl_env
        configure_nccl_env()
        dist.init_process_group("nccl", device_id=dev)

    cfg = DreamLLMConfig.vicuna_7b(num_hidden_layers=args.layers) if args.layers != L else DreamLLMConfig.vicuna_7b()
    torch.manual_seed(1234)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = DreamLLMForCausalMLM(cfg)
    torch.set_default_dtype(old)
    reducer = BucketedGradReducer(model.parameters(), bucket_cap_mb=512.0) if world > 1 else None

    gen = torch.Generator().manual_seed(1234 + rank)
    B, S = args.bs, args.seq
    host_ids = torch.randint(0, 32000, (B, S), generator=gen).pin_memory()
    dev_ids = host_ids.to(dev)
    tokens_per_step = B * S * world

    def step(ids, labels):
        if reducer is not None:
            reducer.zero_grad()
        else:
            for p in model.parameters():
                p.grad = None
        out = model(input_ids=ids, labels=labels, attention_mask_has_padding=False)
        out.loss.backward()
        if reducer is not None:
            reducer.finalize()
        return out.loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- warm-up
    for _ in range(args.warmup):
        step(dev_ids, dev_ids)
    barrier()

    # ---- device-resident timing (inputs already in HBM) with live per-GEMM event timing
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.PROFILE.reset(enabled=True)
    ops.LAUNCHES.reset()
    total_ms = timed(lambda: step(dev_ids, dev_ids), args.steps)
    launches = ops.LAUNCHES.count
    gemm_stats = ops.PROFILE.summary()
    ops.PROFILE.reset(enabled=False)

    # ---- end-to-end through the public API: pinned host inputs -> H2D, loss -> D2H, every step
    def e2e_step():
        ids = host_ids.to(dev, non_blocking=True)
        loss = step(ids, ids)
        return float(loss.item())

    e2e_ms = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = tokens_per_step / (ms_per_step / 1e3)
    e2e_value = tokens_per_step / (e2e_ms / args.steps / 1e3)
    peak_tf, peak_hbm, how = measured_peaks()
    fl = flops_per_step(B, S, layers=args.layers)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (uniform token ids, random-init N(0,0.02) weights)", "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_ids.numel() * 8), "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "dllm::gemm_kernel<2,*,*,bf16> (tcgen05 GEMM, all fwd/dgrad/wgrad launches)",
                     "achieved": gemm_stats["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": gemm_stats["tflops"] / peak_tf if gemm_stats["tflops"] else None, "traffic": gemm_traffic()[0],
                     "traffic_note": gemm_traffic()[1],
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({how})", "launches_timed": gemm_stats["n"],
                     "gemm_share_of_step": gemm_stats["ms"] / total_ms if total_ms else None},
        "step_roofline": {"algorithmic_tflop_per_step_per_gpu": fl / 1e12, "achieved_tflops_per_gpu": fl / 1e12 / (ms_per_step / 1e3),
                          "frac_of_measured_sustained_peak": fl / 1e12 / (ms_per_step / 1e3) / peak_tf},
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference(S, layers_sample=1, iters=1, warm=0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
