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


def causal_additive_mask_past(bsz, q_len, past_len, dtype, attention_mask_2d=None):
    """What `_prepare_4d_causal_attention_mask(mask, (bsz, q_len), embeds, past_len)` yields (:965-967) with a kv-cache: [B, 1, q_len,
    past_len + q_len], query i sits at absolute position past_len + i (bottom-right aligned causal), finfo.min on padded key columns of
    the FULL-length 2-D mask (HF generate keeps extending it, :1511-1547)."""
    mn = torch.finfo(dtype).min
    total = past_len + q_len
    qpos = torch.arange(past_len, total)[:, None]
    kpos = torch.arange(total)[None, :]
    m = torch.zeros(q_len, total, dtype=dtype).masked_fill(kpos > qpos, mn)[None, None].expand(bsz, 1, q_len, total).clone()
    if attention_mask_2d is not None:
        assert attention_mask_2d.shape == (bsz, total)
        m = m.masked_fill((attention_mask_2d == 0)[:, None, None, :], mn)
    return m


def attention_eager(x, wq, wk, wv, wo, num_heads, cos, sin, position_ids, mask4d, past_kv=None, return_kv=False):
    """DreamLLMAttention.forward (eager path), :309-400, pretraining_tp == 1.  `past_kv` = (k, v) [B, nh, past, d] already rotated:
    the new keys / values are concatenated behind them (:344-355) and `(k, v)` is the present (`use_cache`, :355)."""
    bsz, q_len, hidden = x.shape
    d = hidden // num_heads
    q = F.linear(x, wq).view(bsz, q_len, num_heads, d).transpose(1, 2)
    k = F.linear(x, wk).view(bsz, q_len, num_heads, d).transpose(1, 2)
    v = F.linear(x, wv).view(bsz, q_len, num_heads, d).transpose(1, 2)
    q, k = apply_rope(q, k, cos.to(x.dtype), sin.to(x.dtype), position_ids)
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=2)
        v = torch.cat([past_kv[1], v], dim=2)
    present = (k, v)
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(d)
    if mask4d is not None:
        w = w + mask4d
        w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min, dtype=w.dtype))  # :373-375
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)  # :378
    o = torch.matmul(w, v).transpose(1, 2).contiguous().reshape(bsz, q_len, hidden)
    out = F.linear(o, wo)
    return (out, present) if return_kv else out


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


def decoder_layer(x, p: dict, num_heads, cos, sin, position_ids, mask4d, eps=1e-6, past_kv=None, return_kv=False):
    """DreamLLMDecoderLayer.forward, :599-654 (pre-norm residual block). `p` uses the
    reference's state-dict key names (LAYER_KEYS).  With `return_kv` returns (hidden, present_key_value) as `use_cache=True` does."""
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
        past_kv=past_kv,
        return_kv=return_kv,
    )
    present = None
    if return_kv:
        h, present = h
    x = res + h
    res = x
    h = rmsnorm(x, p["post_attention_layernorm.weight"], eps)
    h = mlp(h, p["mlp.gate_proj.weight"], p["mlp.up_proj.weight"], p["mlp.down_proj.weight"])
    return (res + h, present) if return_kv else res + h


def cached_decode_scenario(hidden=256, inter=512, heads=2, seed=21, lens=(37, 50), steps=3):
    """A LEFT-padded prompt batch pushed through one decoder layer with a kv-cache: prefill, then `steps` single-token decode steps, driven
    the way HF generate drives the reference (full 2-D mask every call, position_ids = cumsum(mask) - 1 with 1 at pads, :1511-1547).
    Returns (params, list of (x [B, S_i, H], attention_mask [B, total_i], position_ids [B, S_i])) — deterministic CPU RNG."""
    p = init_layer_params(hidden, inter, seed)
    g = torch.Generator().manual_seed(seed + 500)
    B, S0 = len(lens), max(lens)
    am = torch.zeros(B, S0, dtype=torch.long)
    for b, n in enumerate(lens):
        am[b, S0 - n:] = 1
    calls = []
    x0 = torch.randn(B, S0, hidden, generator=g)
    pos = (am.cumsum(-1) - 1).masked_fill(am == 0, 1)
    calls.append((x0, am.clone(), pos))
    for _ in range(steps):
        am = torch.cat([am, torch.ones(B, 1, dtype=torch.long)], 1)
        calls.append((torch.randn(B, 1, hidden, generator=g), am.clone(), am.sum(-1, keepdim=True) - 1))
    return p, calls


def run_cached_scenario(p, calls, heads, dtype=torch.float32):
    """Oracle outputs of `cached_decode_scenario`: list of hidden states [B, S_i, H] (pad rows included; compare valid rows only)."""
    pp = {k: v.to(dtype) for k, v in p.items()}
    d = calls[0][0].shape[-1] // heads
    cos, sin = rope_tables(d, 2048, dtype=dtype)
    past, outs = None, []
    for x, am, pos in calls:
        past_len = 0 if past is None else past[0].shape[2]
        mask = causal_additive_mask_past(x.shape[0], x.shape[1], past_len, dtype, am)
        y, past = decoder_layer(x.to(dtype), pp, heads, cos, sin, pos, mask, past_kv=past, return_kv=True)
        outs.append(y)
    return outs


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
