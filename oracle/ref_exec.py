"""TEST INFRASTRUCTURE ONLY — loads the *reference's own* decoder code for golden generation.

Executes lines 69..655 of /root/reference/omni/models/dreamllm/modeling_dreamllm.py
(`_get_unpad_data` .. end of `DreamLLMDecoderLayer`) verbatim in a private namespace,
exactly as SURVEY.md §8(c) describes.  Nothing is copied into this repository: the
source is read from /root/reference at run time, so this module only works inside the
build container (the GPU box has no /root/reference; it uses tests/golden/*.npz).

Only `oracle/gen_golden.py` and the `not gpu` pinning test import this.
"""
from __future__ import annotations

import math
import os
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_FILE = "/root/reference/omni/models/dreamllm/modeling_dreamllm.py"
FIRST_LINE, LAST_LINE = 69, 655


def available() -> bool:
    return os.path.exists(REF_FILE)


def load_reference_namespace() -> dict:
    from transformers.activations import ACT2FN
    from transformers.modeling_attn_mask_utils import _prepare_4d_causal_attention_mask

    with open(REF_FILE) as f:
        lines = f.readlines()
    src = "from __future__ import annotations\n" + "".join(lines[FIRST_LINE - 1 : LAST_LINE])

    class _Logger:
        def __getattr__(self, name):
            return lambda *a, **k: None

    ns = {
        "math": math,
        "torch": torch,
        "nn": nn,
        "F": F,
        "ACT2FN": ACT2FN,
        "ALL_LAYERNORM_LAYERS": [],
        "DreamLLMConfig": object,
        "logger": _Logger(),
        "_prepare_4d_causal_attention_mask": _prepare_4d_causal_attention_mask,
    }
    exec(compile(src, REF_FILE, "exec"), ns)
    return ns


def make_config(hidden_size, intermediate_size, num_heads, max_pos=2048, eps=1e-6, rope_theta=10000.0):
    return types.SimpleNamespace(
        hidden_size=hidden_size,
        intermediate_size=intermediate_size,
        num_attention_heads=num_heads,
        num_key_value_heads=num_heads,
        max_position_embeddings=max_pos,
        rope_theta=rope_theta,
        rope_scaling=None,
        attention_bias=False,
        hidden_act="silu",
        rms_norm_eps=eps,
        pretraining_tp=1,
        _flash_attn_2_enabled=False,
    )


def causal_mask_4d(bsz, seq, dtype, attention_mask_2d=None):
    """What DreamLLMModel._forward builds for the eager path (modeling_dreamllm.py:965-967)."""
    from transformers.modeling_attn_mask_utils import _prepare_4d_causal_attention_mask

    dummy = torch.zeros(bsz, seq, 1, dtype=dtype)
    return _prepare_4d_causal_attention_mask(attention_mask_2d, (bsz, seq), dummy, 0)
