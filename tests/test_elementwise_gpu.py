"""HBM-bound kernels vs the CPU oracle (oracle/decoder_oracle.py) on the same seeded inputs, through the C ABI."""
import pytest
import torch

from oracle import decoder_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("T,H", [(7, 256), (130, 4096), (33, 1280)])
def test_rmsnorm_fwd_bitexact_vs_oracle(T, H):
    from dreamllm_b200 import ops
    x = torch.randn(T, H, generator=_g(1)).to(BF)
    w = (1 + 0.1 * torch.randn(H, generator=_g(2))).to(BF)
    want = O.rmsnorm(x, w, 1e-6)
    y, rstd, _ = ops.rmsnorm_fwd(x.cuda(), w.cuda(), 1e-6)
    # same rounding points (cast before weight multiply); reduction order may flip the last bf16 bit on rare elements
    diff = (y.cpu().float() - want.float()).abs()
    assert float((diff > 0).float().mean()) < 0.01
    torch.testing.assert_close(y.cpu().float(), want.float(), rtol=8e-3, atol=1e-6)
    # fused residual add
    a = torch.randn(T, H, generator=_g(3)).to(BF)
    y2, _, xs = ops.rmsnorm_fwd(x.cuda(), w.cuda(), 1e-6, add=a.cuda())
    assert torch.equal(xs.cpu(), x + a)
    torch.testing.assert_close(y2.cpu().float(), O.rmsnorm(x + a, w, 1e-6).float(), rtol=8e-3, atol=1e-6)


@pytest.mark.parametrize("T,H", [(64, 512), (300, 4096)])
def test_rmsnorm_bwd_vs_oracle_autograd(T, H):
    from dreamllm_b200 import ops
    x = torch.randn(T, H, generator=_g(4)).to(BF)
    w = (1 + 0.1 * torch.randn(H, generator=_g(5))).to(BF)
    dy = torch.randn(T, H, generator=_g(6)).to(BF)
    dres = torch.randn(T, H, generator=_g(7)).to(BF)
    x32 = x.float().requires_grad_(True)
    w32 = w.float().requires_grad_(True)
    O.rmsnorm(x32, w32, 1e-6).backward(dy.float())
    _, rstd, _ = ops.rmsnorm_fwd(x.cuda(), w.cuda(), 1e-6)
    dx, dw = ops.rmsnorm_bwd(dy.cuda(), x.cuda(), w.cuda(), rstd, dres=dres.cuda())
    torch.testing.assert_close(dx.cpu().float(), x32.grad + dres.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dw.cpu().float(), w32.grad, rtol=2e-2, atol=0.02 * T ** 0.5)


@pytest.mark.parametrize("d", [128, 64])
def test_rope_fwd_bitexact_and_bwd_transpose(d):
    from dreamllm_b200 import ops
    B, S, nh = 2, 50, 3
    T = B * S
    qkv = torch.randn(T, 3 * nh * d, generator=_g(8)).to(BF)
    cos, sin = O.rope_tables(d, 2048, dtype=BF)
    pos = torch.arange(S).repeat(B)
    q = qkv[:, : nh * d].view(B, S, nh, d).transpose(1, 2)
    k = qkv[:, nh * d: 2 * nh * d].view(B, S, nh, d).transpose(1, 2)
    qr, kr = O.apply_rope(q, k, cos, sin, pos.view(B, S))
    buf = qkv.cuda().clone()
    ops.rope_(buf, cos.cuda(), sin.cuda(), pos.int().cuda(), 2 * nh, d)
    got = buf.cpu()
    assert torch.equal(got[:, : nh * d].view(B, S, nh, d).transpose(1, 2), qr)
    assert torch.equal(got[:, nh * d: 2 * nh * d].view(B, S, nh, d).transpose(1, 2), kr)
    assert torch.equal(got[:, 2 * nh * d:], qkv[:, 2 * nh * d:])          # v untouched
    # backward == transpose of forward: <rope(x), g> == <x, rope^T(g)> in fp32-ish
    g = torch.randn(T, 3 * nh * d, generator=_g(9)).to(BF)
    gb = g.cuda().clone()
    ops.rope_(gb, cos.cuda(), sin.cuda(), pos.int().cuda(), 2 * nh, d, backward=True)
    lhs = (got[:, : 2 * nh * d].double() * g[:, : 2 * nh * d].double()).sum()
    rhs = (qkv[:, : 2 * nh * d].double() * gb.cpu()[:, : 2 * nh * d].double()).sum()
    assert abs(float(lhs - rhs)) < 2e-2 * (abs(float(lhs)) + T)


def test_swiglu_fwd_bitexact_bwd_vs_autograd():
    from dreamllm_b200 import ops
    T, I = 77, 1408
    gu = torch.randn(T, 2 * I, generator=_g(10)).to(BF)
    g, u = gu[:, :I], gu[:, I:]
    want = torch.nn.functional.silu(g) * u
    act = ops.swiglu_fwd(gu.cuda(), I)
    assert float((act.cpu() != want).float().mean()) < 2e-3       # expf vs torch's vectorised exp: rare 1-ulp flips
    torch.testing.assert_close(act.cpu().float(), want.float(), rtol=8e-3, atol=1e-6)
    dact = torch.randn(T, I, generator=_g(11)).to(BF)
    g32 = g.float().requires_grad_(True)
    u32 = u.float().requires_grad_(True)
    (torch.nn.functional.silu(g32) * u32).backward(dact.float())
    dgu = ops.swiglu_bwd(dact.cuda(), gu.cuda(), I).cpu().float()
    torch.testing.assert_close(dgu[:, :I], g32.grad, rtol=3e-2, atol=2e-2)
    torch.testing.assert_close(dgu[:, I:], u32.grad, rtol=3e-2, atol=2e-2)


def test_cross_entropy_loss_and_grad_vs_oracle():
    from dreamllm_b200 import ops
    B, S, V = 2, 37, 32008
    logits = (torch.randn(B, S, V, generator=_g(12)) * 2).to(BF)
    labels = torch.randint(0, V, (B, S), generator=_g(13))
    labels[0, 5:9] = -100
    lf = logits.float().requires_grad_(True)
    want = O.lm_loss(lf, labels)
    want.backward()
    shifted = torch.full((B, S), -100, dtype=torch.int64)
    shifted[:, :-1] = labels[:, 1:]
    lg = logits.cuda().view(B * S, V).clone()
    loss = ops.cross_entropy_(lg, shifted.view(-1).cuda())
    assert abs(float(loss) - float(want)) < 2e-4 * abs(float(want))
    torch.testing.assert_close(lg.cpu().float().view(B, S, V), lf.grad, rtol=1e-2, atol=1e-6)
    # no valid label -> loss 0, grads 0 (reference :1468-1469)
    lg2 = logits.cuda().view(B * S, V).clone()
    loss2 = ops.cross_entropy_(lg2, torch.full((B * S,), -100, dtype=torch.int64, device="cuda"))
    assert float(loss2) == 0.0 and float(lg2.abs().max()) == 0.0


def test_embedding_bitexact_and_deterministic_grad():
    from dreamllm_b200 import ops
    V, H, T = 1000, 256, 333
    W = torch.randn(V, H, generator=_g(14)).to(BF)
    ids = torch.randint(0, 40, (3, 111), generator=_g(15))      # many collisions
    out = ops.embedding_fwd(ids.cuda(), W.cuda())
    assert torch.equal(out.cpu(), W[ids])
    dy = torch.randn(T, H, generator=_g(16)).to(BF)
    dW = ops.embedding_bwd(ids.cuda(), dy.cuda(), V)
    ref = torch.zeros(V, H).index_add_(0, ids.view(-1), dy.float())
    torch.testing.assert_close(dW.cpu().float(), ref, rtol=8e-3, atol=1e-2)
    assert torch.equal(dW, ops.embedding_bwd(ids.cuda(), dy.cuda(), V))
