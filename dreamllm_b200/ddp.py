"""Bucketed gradient all-reduce overlapped with backward (data-parallel hot path).

Replaces what the reference gets implicitly from `DistributedDataParallel` via HF Trainer / accelerate
(omni/train/trainer.py:577-601: `_wrap_model` -> DDP, 25 MB buckets, all-reduce overlapped with backward).
One process per GPU; torch.distributed (NCCL over NVLink/NVSwitch; gloo in the CPU tests) is the transport.

Owning the reducer removes the reference's "dummy forward" hacks (modeling_dreamllm.py:1142-1144, :1443-1445,
modeling_plugins.py:315-329): hooks are registered only on parameters with requires_grad, and a parameter that
received no gradient in a step contributes zeros to its bucket at `finalize()`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


import os as _os

NCCL_CTAS = int(_os.environ.get("DLLM_NCCL_CTAS", "0"))        # SM budget handed to the overlapped all-reduce (0 = leave NCCL alone)
RESERVED_SMS = int(_os.environ.get("DLLM_RESERVED_SMS", "0"))  # CTA pairs kept out of the persistent GEMM grids while grads are in flight


def configure_nccl_env():
    """Optional knobs (off by default — measured on 2 x B200, ms/step: NCCL default + dynamic GEMM scheduler is best;
    capping NCCL to 2 CTAs exposes the all-reduce: 799 ms; reserving SMs for it: 704 ms; see profiles/r01_ddp_n2_variants.md)."""
    import os
    if NCCL_CTAS > 0:
        os.environ.setdefault("NCCL_MAX_CTAS", str(NCCL_CTAS))
        os.environ.setdefault("NCCL_MIN_CTAS", "1")


class BucketedGradReducer:
    def __init__(self, params, bucket_cap_mb: float = 256.0, process_group=None, average: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        self.params = [p for p in params if p.requires_grad]
        # gradients become ready roughly in reverse registration order -> fill buckets in that order
        order = list(reversed(self.params))
        cap = int(bucket_cap_mb * 1024 * 1024)
        self.buckets = []           # dict(flat, views{param: view}, pending:set, params:list, work)
        cur, cur_bytes = [], 0
        for p in order:
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > cap or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._make_bucket(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._make_bucket(cur)
        self._bucket_of = {}
        for b in self.buckets:
            for p in b["params"]:
                self._bucket_of[p] = b
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        backend = dist.get_backend(process_group) if dist.is_initialized() else None
        self._use_avg_op = backend == "nccl"
        self.launched = 0
        if backend == "nccl" and self.world > 1:
            from ._lib import lib
            lib().dllm_set_reserved_sms(RESERVED_SMS)

    def _make_bucket(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=plist[0].dtype, device=plist[0].device)
        views, off = {}, 0
        for p in plist:
            views[p] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append(dict(flat=flat, views=views, params=list(plist), pending=set(plist), work=None))

    # called by autograd right after p.grad has been written for this backward pass
    def _on_grad(self, p):
        b = self._bucket_of[p]
        v = b["views"][p]
        if p.grad.data_ptr() != v.data_ptr():
            v.copy_(p.grad)
            p.grad = v                    # the reduced result lands directly in p.grad (no copy back)
        b["pending"].discard(p)
        if not b["pending"]:
            self._launch(b)

    def _launch(self, b):
        if self.world == 1:
            return
        flat = b["flat"]
        if self._use_avg_op and self.average:
            b["work"] = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            if self.average:
                flat.div_(self.world)
            b["work"] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.launched += 1

    def finalize(self):
        """Call after backward: flush buckets with missing grads (treated as zero) and wait for every all-reduce."""
        for b in self.buckets:
            if b["pending"]:
                for p in b["pending"]:
                    b["views"][p].zero_()
                    p.grad = b["views"][p]
                self._launch(b)
            if b["work"] is not None:
                b["work"].wait()
                b["work"] = None
            b["pending"] = set(b["params"])

    def zero_grad(self):
        """Keep grads as bucket views; autograd then accumulates in place and no copy into the bucket is needed."""
        for b in self.buckets:
            b["flat"].zero_()
            for p in b["params"]:
                p.grad = None

    def remove(self):
        for h in self._hooks:
            h.remove()
