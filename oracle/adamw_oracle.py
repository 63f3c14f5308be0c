"""TEST INFRASTRUCTURE — CPU checker for the sharded AdamW step (dreamllm_b200/zero.py, csrc/optim.cu).  Never imported by the product.

The reference's optimizer is `torch.optim.AdamW` (`optim="adamw_torch"`, projects/dreamllm/configs/stage1/base.py:85, stage2/base.py:95),
stepping a model loaded in bf16 (projects/dreamllm/train.py:68-70, :138), after `clip_grad_norm_` (omni/train/trainer.py:800-807).
torch is installed here AND on the GPU box, so this oracle is **pinned**: tests/test_zero_cpu.py checks the flat restatement below against
`torch.optim.AdamW(foreach=False)` itself, bit for bit, in bf16 and in fp32.

`adamw_flat_` restates torch/optim/adamw.py `_single_tensor_adamw` (same ATen calls, same order) on flat shards:
    p.mul_(1 - lr*wd); m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, value=1-b2)
    denom = (v.sqrt() / sqrt(1-b2^t)).add_(eps); p.addcdiv_(m, denom, value=-lr/(1-b1^t))
With bf16 tensors every one of those ops rounds to bf16 — that is the arithmetic the reference runs.
"""
from __future__ import annotations

import math

import torch


def clip_coef(total_sumsq: torch.Tensor, max_norm: float) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max=1)."""
    return torch.clamp(max_norm / (total_sumsq.float().sqrt() + 1e-6), max=1.0)


def adamw_flat_(g, p, m, v, master, *, lr, beta1, beta2, eps, weight_decay, step, grad_sumsq=None, max_grad_norm=0.0):
    """Signature of ShardedAdamW's `update_fn`.  master None: p/m/v updated in their own dtype (bf16 = reference arithmetic).
    master given: fp32 master/m/v updated, p rewritten as master.to(bf16)."""
    tgt = p if master is None else master
    grad = g.to(tgt.dtype)
    if grad_sumsq is not None and max_grad_norm > 0:
        grad = (g.float() * clip_coef(grad_sumsq, max_grad_norm)).to(tgt.dtype)   # g.mul_(coef): fp32 product, rounds in the gradient dtype
    if weight_decay != 0:
        tgt.mul_(1 - lr * weight_decay)
    m.lerp_(grad, 1 - beta1)
    v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    step_size = lr / bias_correction1
    denom = (v.sqrt() / math.sqrt(bias_correction2)).add_(eps)
    tgt.addcdiv_(m, denom, value=-step_size)
    if master is not None:
        p.copy_(master)


def sumsq_flat(g, out):
    out += g.float().pow(2).sum()
