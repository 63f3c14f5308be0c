"""GPU parity of the optimizer-shard kernels (csrc/optim.cu) and of ShardedAdamW's CUDA path against the pinned CPU oracle
(oracle/adamw_oracle.py == torch.optim.AdamW, the reference's `optim="adamw_torch"`)."""
import pytest
import torch

from oracle import adamw_oracle as AO

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
HP = dict(lr=2e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01)


def _rand(n, scale, seed, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, generator=g) * scale).to(dtype)


@pytest.mark.parametrize("n", [8, 1024 + 8, (1 << 20) + 128])
def test_sumsq_bf16(n):
    from dreamllm_b200 import ops
    x = _rand(n, 0.3, n)
    out = torch.full((1,), 5.0, device="cuda", dtype=torch.float32)
    ops.sumsq_bf16_(x.cuda(), out, accumulate=True)
    want = x.double().pow(2).sum() + 5.0
    torch.testing.assert_close(out.cpu().double()[0], want, rtol=2e-6, atol=0)
    ops.sumsq_bf16_(x.cuda(), out, accumulate=False)
    torch.testing.assert_close(out.cpu().double()[0], want - 5.0, rtol=2e-6, atol=0)
    again = torch.zeros(1, device="cuda")
    ops.sumsq_bf16_(x.cuda(), again, accumulate=False)
    assert torch.equal(again, out)                      # deterministic reduction order


@pytest.mark.parametrize("clip", [False, True])
def test_adamw_bf16_state_is_the_reference_arithmetic(clip):
    """bf16 param / exp_avg / exp_avg_sq with per-op rounding: equal to torch.optim.AdamW on bf16 tensors (via the pinned oracle).
    Bit-exact except where the CPU's ATen kernels contract a*b+c into an FMA (<= 1 bf16 ulp, rare)."""
    from dreamllm_b200 import ops
    n = (1 << 18) + 264
    p = _rand(n, 0.05, 1)
    m, v = torch.zeros(n, dtype=BF), torch.zeros(n, dtype=BF)
    dp, dm, dv = p.cuda(), m.cuda(), v.cuda()
    for step in range(1, 4):
        g = _rand(n, 0.1, 10 + step)
        ss = g.float().pow(2).sum().reshape(1)
        kw = dict(grad_sumsq=ss, max_grad_norm=1.0) if clip else {}
        AO.adamw_flat_(g, p, m, v, None, step=step, **HP, **kw)
        dkw = dict(grad_sumsq=ss.cuda(), max_grad_norm=1.0) if clip else {}
        ops.adamw_step_(g.cuda(), dp, dm, dv, None, step=step, **HP, **dkw)
        # absolute slack = one bf16 ulp at each tensor's working scale (update ~ lr, exp_avg ~ 0.1 |g|, exp_avg_sq ~ 1e-3 g^2): results that
        # cancel to ~0 inherit the 1-ulp differences of the few elements that already differed
        for name, got, want, atol in (("param", dp, p, 4e-5), ("exp_avg", dm, m, 2.5e-4), ("exp_avg_sq", dv, v, 1e-7)):
            got = got.cpu()
            exact = (got == want).float().mean().item()
            assert exact > (0.995 if step == 1 else 0.97), f"step {step} {name}: only {exact:.4f} bit-exact"
            torch.testing.assert_close(got.float(), want.float(), rtol=2 ** -6, atol=atol, msg=f"step {step} {name}")


@pytest.mark.parametrize("clip", [False, True])
def test_adamw_fp32_master(clip):
    from dreamllm_b200 import ops
    n = (1 << 18) + 8
    p = _rand(n, 0.05, 2)
    master = p.float()
    m, v = torch.zeros(n), torch.zeros(n)
    dp, dw, dm, dv = p.cuda(), master.cuda(), m.cuda(), v.cuda()
    for step in range(1, 4):
        g = _rand(n, 0.1, 20 + step)
        ss = g.float().pow(2).sum().reshape(1)
        kw = dict(grad_sumsq=ss, max_grad_norm=0.5) if clip else {}
        AO.adamw_flat_(g, p, m, v, master, step=step, **HP, **kw)
        dkw = dict(grad_sumsq=ss.cuda(), max_grad_norm=0.5) if clip else {}
        ops.adamw_step_(g.cuda(), dp, dm, dv, dw, step=step, **HP, **dkw)
        torch.testing.assert_close(dw.cpu(), master, rtol=1e-5, atol=1e-7)     # atol: fp32 ulp at the parameter scale (cancellation)
        torch.testing.assert_close(dm.cpu(), m, rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(dv.cpu(), v, rtol=1e-5, atol=1e-12)
        assert torch.equal(dp, dw.to(BF))                # the bf16 parameter is the rounded master, written by the same kernel


def test_adamw_rejects_bad_shapes():
    from dreamllm_b200 import ops
    x = torch.zeros(12, device="cuda", dtype=BF)
    with pytest.raises(RuntimeError):
        ops.adamw_step_(x, x.clone(), x.clone(), x.clone(), None, step=1, **HP)       # n % 8 != 0
    with pytest.raises(RuntimeError):
        ops.adamw_step_(torch.zeros(16, dtype=BF), torch.zeros(16, dtype=BF), torch.zeros(16, dtype=BF), torch.zeros(16, dtype=BF), None,
                        step=1, **HP)                                                     # CPU tensors: no fallback


@pytest.mark.parametrize("state_dtype", [BF, torch.float32])
def test_sharded_adamw_steps_a_causal_lm_like_the_oracle(state_dtype):
    """ShardedAdamW on its CUDA kernels trains a tiny DreamLLMForCausalMLM (q|k|v and gate|up weights are fused row blocks living in the
    optimizer's flat buckets) while a shadow instance — same host logic, the oracle's arithmetic, the same gradients — tracks it:
    parameters agree to ~1 bf16 ulp step after step, and the loss goes down."""
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    from dreamllm_b200.zero import ShardedAdamW
    cfg = DreamLLMConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2)
    ids = torch.randint(0, 512, (2, 96), generator=torch.Generator().manual_seed(3)).cuda()
    models, opts = [], []
    for use_oracle in (False, True):
        torch.manual_seed(0)
        mdl = DreamLLMForCausalMLM(cfg).to(device="cuda", dtype=BF)
        kw = dict(update_fn=AO.adamw_flat_, sumsq_fn=AO.sumsq_flat) if use_oracle else {}
        opts.append(ShardedAdamW(mdl.parameters(), lr=1e-3, weight_decay=0.0, max_grad_norm=1.0, bucket_cap_mb=0.5,
                                 state_dtype=state_dtype, **kw))
        models.append(mdl)
    assert len(opts[0].buckets) > 2
    losses = []
    for step in range(4):
        opts[0].zero_grad()
        out = models[0](input_ids=ids, labels=ids)
        out.loss.backward()
        losses.append(float(out.loss))
        opts[1].zero_grad()
        for a, b in zip(models[0].parameters(), models[1].parameters()):     # the shadow steps on the very same gradients
            assert a.grad is not None
            b.grad = a.grad.clone()
            opts[1]._on_grad(b)
        n0 = opts[0].step()
        n1 = opts[1].step()
        torch.testing.assert_close(n0, n1, rtol=1e-4, atol=0)
        for (k, a), (_, b) in zip(models[0].named_parameters(), models[1].named_parameters()):
            torch.testing.assert_close(a.detach().float(), b.detach().float(), rtol=2 ** -5, atol=5e-5, msg=f"step {step} {k}")
        with torch.no_grad():                                                 # keep 1-ulp rounding differences from compounding
            for a, b in zip(models[0].parameters(), models[1].parameters()):
                b.copy_(a)
    assert losses[-1] < losses[0] - 0.05, losses
