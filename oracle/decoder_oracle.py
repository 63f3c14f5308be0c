"""TEST INFRASTRUCTURE ONLY — CPU oracle for the DreamLLM decoder hot path.

A plain-torch (CPU, fp32 or bf16) restatement of the reference's algorithm, written
functionally so it can travel to the GPU box (which has no /root/reference).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; the product (`dreamllm_b200/`) never does.

Pinning: `tests/test_oracle_pin.py` checks every function here against
  (a) `tests/golden/*.npz`, minted by `oracle/gen_golden.py` from the reference's own code
      executed verbatim (`oracle/ref_exec.py`), and
  (b) the live reference when /root/reference is present (build container only).
The reference itself ships no tests / golden vectors for this path.

All citations are to /root/reference/omni/models/dreamllm/modeling_dreamllm.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- a1
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """DreamLLMRMSNorm.forward, :86-91.  NB the cast to the input dtype happens *before*
    the weight multiply (:91)."""
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return weight * h.to(dt)


# --------------------------------------------------------------------------- a2
def rope_tables(head_dim: int, max_pos: int, base: float = 10000.0, dtype=torch.float32):
    """RotaryEmbedding.__init__/_set_cos_sin_cache, :97-118: inv_freq = base^(-2i/d),
    emb = cat(freqs, freqs); tables are built in fp32 and cast to the model dtype (:126-127)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_pos, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    """:176-180 (half-split, not interleaved)."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2 :]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q, k, cos, sin, position_ids):
    """apply_rotary_pos_emb, :184-209. q,k: [B, nh, S, d]; cos/sin: [max_pos, d]."""
    cos = cos[position_ids].unsqueeze(1)
    sin = sin[position_ids].unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


# --------------------------------------------------------------------------- a3
def causal_additive_mask(bsz, seq, dtype, attention_mask_2d=None):
    """What `_prepare_4d_causal_attention_mask` yields for past_len=0 (:965-967):
    finfo.min above the diagonal and on padded key columns."""
    mn = torch.finfo(dtype).min
    m = torch.full((seq, seq), mn, dtype=dtype).triu(1)[None, None].expand(bsz, 1, seq, seq).clone()
    if attention_mask_2d is not None:
        pad = (attention_mask_2d == 0)[:, None, None, :]
        m = m.masked_fill(pad, mn)
    return m


def attention_eager(x, wq, wk, wv, wo, num_heads, cos, sin, position_ids, mask4d):
    """DreamLLMAttention.forward (eager path), :309-400, pretraining_tp == 1, no kv-cache."""
    bsz, q_len, hidden = x.shape
    d = hidden // num_heads
    q = F.linear(x, wq).view(bsz, q_len, num_heads, d).transpose(1, 2)
    k = F.linear(x, wk).view(bsz, q_len, num_heads, d).transpose(1, 2)
    v = F.linear(x, wv).view(bsz, q_len, num_heads, d).transpose(1, 2)
    q, k = apply_rope(q, k, cos.to(x.dtype), sin.to(x.dtype), position_ids)
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(d)
    if mask4d is not None:
        w = w + mask4d
        w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min, dtype=w.dtype))  # :373-375
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)  # :378
    o = torch.matmul(w, v).transpose(1, 2).contiguous().reshape(bsz, q_len, hidden)
    return F.linear(o, wo)


# --------------------------------------------------------------------------- a6
def mlp(x, w_gate, w_up, w_down):
    """DreamLLMMLP.forward, :237: down(silu(gate(x)) * up(x))."""
    return F.linear(F.silu(F.linear(x, w_gate)) * F.linear(x, w_up), w_down)


# --------------------------------------------------------------------------- a7
LAYER_KEYS = (
    "self_attn.q_proj.weight",
    "self_attn.k_proj.weight",
    "self_attn.v_proj.weight",
    "self_attn.o_proj.weight",
    "mlp.gate_proj.weight",
    "mlp.up_proj.weight",
    "mlp.down_proj.weight",
    "input_layernorm.weight",
    "post_attention_layernorm.weight",
)


def decoder_layer(x, p: dict, num_heads, cos, sin, position_ids, mask4d, eps=1e-6):
    """DreamLLMDecoderLayer.forward, :599-654 (pre-norm residual block). `p` uses the
    reference's state-dict key names (LAYER_KEYS)."""
    res = x
    h = rmsnorm(x, p["input_layernorm.weight"], eps)
    h = attention_eager(
        h,
        p["self_attn.q_proj.weight"],
        p["self_attn.k_proj.weight"],
        p["self_attn.v_proj.weight"],
        p["self_attn.o_proj.weight"],
        num_heads,
        cos,
        sin,
        position_ids,
        mask4d,
    )
    x = res + h
    res = x
    h = rmsnorm(x, p["post_attention_layernorm.weight"], eps)
    h = mlp(h, p["mlp.gate_proj.weight"], p["mlp.up_proj.weight"], p["mlp.down_proj.weight"])
    return res + h


def init_layer_params(hidden, inter, seed, dtype=torch.float32, std=0.02):
    """Synthetic weights as `_init_weights` draws them (:674-683): N(0, 0.02); norm weights 1."""
    g = torch.Generator().manual_seed(seed)
    shapes = {
        "self_attn.q_proj.weight": (hidden, hidden),
        "self_attn.k_proj.weight": (hidden, hidden),
        "self_attn.v_proj.weight": (hidden, hidden),
        "self_attn.o_proj.weight": (hidden, hidden),
        "mlp.gate_proj.weight": (inter, hidden),
        "mlp.up_proj.weight": (inter, hidden),
        "mlp.down_proj.weight": (hidden, inter),
    }
    p = {k: (torch.randn(s, generator=g) * std).to(dtype) for k, s in shapes.items()}
    # non-trivial norm weights so their gradient path is exercised
    p["input_layernorm.weight"] = (1.0 + 0.1 * torch.randn(hidden, generator=g)).to(dtype)
    p["post_attention_layernorm.weight"] = (1.0 + 0.1 * torch.randn(hidden, generator=g)).to(dtype)
    return p


# --------------------------------------------------------------------------- a13 (text-only part)
def lm_loss(logits_fp32, labels):
    """DreamLLMForCausalMLM.forward, :1453-1470: shift, CE(reduction none), masked mean over
    labels != -100 (plain mean if no valid label)."""
    shift_logits = logits_fp32[..., :-1, :].contiguous()
    shift_labels = labels[..., 1:].contiguous()
    V = shift_logits.shape[-1]
    ce = F.cross_entropy(shift_logits.view(-1, V), shift_labels.view(-1), reduction="none")
    valid = (shift_labels.view(-1) != -100)
    if valid.sum() > 0:
        return (ce * valid).sum() / valid.sum()
    return ce.mean()


def causal_lm(input_ids, labels, embed, layers: list, norm_w, lm_head_w, num_heads, eps=1e-6,
              attention_mask=None, max_pos=2048, inputs_embeds=None):
    """Text-only DreamLLMForCausalMLM forward: embed_tokens (:1066) → L × decoder layer (:986-1014)
    → final RMSNorm (:1024) → lm_head (:1452) → fp32 logits (:1453) → lm_loss.
    Returns (loss, logits_fp32, last_hidden)."""
    x = F.embedding(input_ids, embed) if inputs_embeds is None else inputs_embeds
    bsz, seq = x.shape[:2]
    d = x.shape[-1] // num_heads
    cos, sin = rope_tables(d, max_pos, dtype=x.dtype)
    pos = torch.arange(seq)[None].expand(bsz, -1)
    mask = causal_additive_mask(bsz, seq, x.dtype, attention_mask)
    for p in layers:
        x = decoder_layer(x, p, num_heads, cos, sin, pos, mask, eps)
    h = rmsnorm(x, norm_w, eps)
    logits = F.linear(h, lm_head_w).float()
    loss = lm_loss(logits, labels) if labels is not None else None
    return loss, logits, h
