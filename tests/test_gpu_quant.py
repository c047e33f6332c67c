"""Fake-quant CUDA kernels vs the oracle: bit-exact (integer/bit work), through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def lognormal(n, seed=1234, specials=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g) * torch.exp(torch.empty(n).uniform_(-12, 8, generator=g))
    if specials and n >= 16:
        x[:12] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 65504.0, 65520.0, 2.0 ** -25,
                               -2.0 ** -25, 2.0 ** -14, 1 + 2.0 ** -11, -(1 + 2.0 ** -11), 1e-40])
    return x


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("exp,man", [(5, 10), (8, 10), (4, 3), (5, 2), (8, 7), (8, 22), (2, 1)])
@pytest.mark.parametrize("n", [0, 1, 3, 5, 1023, 1 << 20])
def test_float_quantize_nearest_bit_exact(exp, man, n):
    import mv_native as mv
    from oracle import quant_oracle as qo
    x = lognormal(n)
    got = mv.float_quantize(x.cuda(), exp, man).cpu().numpy()
    assert np.array_equal(bits(got), bits(qo.float_quantize(x.numpy(), exp, man)))


@pytest.mark.parametrize("exp,man", [(5, 10), (8, 10), (8, 7), (4, 3), (5, 6)])
def test_float_quantize_stochastic_same_stream_bit_exact(exp, man):
    """man >= 7 consumes the 16-bit stream (eight elements per Philox call), man < 7 the 32-bit one (mv_b200.h)."""
    import mv_native as mv
    from oracle import quant_oracle as qo
    n = 300001                                  # ragged: exercises the scalar tail
    x = lognormal(n)
    half = man >= 7
    for seed, off in ((77, 5), (2 ** 40 + 3, 2 ** 33)):
        r = qo.philox_bits(n, seed, off, half=half)
        assert np.array_equal(mv.philox_bits(n, seed, off, half=half).cpu().numpy().view(np.uint32), r)
        want = qo.float_quantize(x.numpy(), exp, man, "stochastic", r)
        got = mv.float_quantize(x.cuda(), exp, man, "stochastic", seed=seed, offset=off).cpu().numpy()
        assert np.array_equal(bits(got), bits(want))
        xu = torch.cat([torch.zeros(1), x]).cuda()[1:]          # 4-byte aligned only: the scalar kernel, same stream
        got = mv.float_quantize(xu, exp, man, "stochastic", seed=seed, offset=off).cpu().numpy()
        assert np.array_equal(bits(got), bits(want))
        if exp <= 5 and man <= 10:                               # fp16 container: same values
            got = mv.float_quantize(x.cuda(), exp, man, "stochastic", seed=seed, offset=off,
                                    out_dtype=torch.float16).float().cpu().numpy()
            assert np.array_equal(bits(got), bits(want))


def test_unaligned_views_and_fp16_container():
    import mv_native as mv
    from oracle import quant_oracle as qo
    x = lognormal(4099)
    xd = x.cuda()
    view = xd[1:]                                # 4-byte aligned only
    got = mv.float_quantize(view, 5, 10).cpu().numpy()
    assert np.array_equal(bits(got), bits(qo.float_quantize(x[1:].numpy(), 5, 10)))
    h = mv.float_quantize(xd, 5, 10, out_dtype=torch.float16)
    assert np.array_equal(bits(h.float().cpu().numpy()), bits(qo.float_quantize(x.numpy(), 5, 10)))
    with pytest.raises(mv.MvError, match="fp16 container"):
        mv.float_quantize(xd, 8, 10, out_dtype=torch.float16)


@pytest.mark.parametrize("fl", [9, 8, 7])
def test_fixed_point_bit_exact(fl):
    import mv_native as mv
    from oracle import quant_oracle as qo
    n = 200003
    x = (lognormal(n, specials=False) * 1e-2).clamp(-20, 20)
    xd = x.cuda()
    assert np.array_equal(mv.fixed_point_quantize(xd, 11, fl).cpu().numpy(),
                          qo.fixed_point_quantize(x.numpy(), 11, fl))
    assert np.array_equal(mv.fixed_point_quantize(xd, 11, fl, clamp=False, symmetric=True).cpu().numpy(),
                          qo.fixed_point_quantize(x.numpy(), 11, fl, clamp=False, symmetric=True))
    u = qo.philox_uniform(n, 3, 1)
    got = mv.fixed_point_quantize(xd, 11, fl, rounding="stochastic", seed=3, offset=1).cpu().numpy()
    assert np.array_equal(got, qo.fixed_point_quantize(x.numpy(), 11, fl, rounding="stochastic", runif=u))
    o, m = mv.fixed_point_quantize(xd, 11, fl, with_mask=True)
    wo, wm = qo.fixed_point_quantize_mask(x.numpy(), 11, fl)
    assert np.array_equal(o.cpu().numpy(), wo) and np.array_equal(m.cpu().numpy(), wm)


@pytest.mark.parametrize("shape", [(64, 96, 40), (7, 9, 5)])     # 16-byte vector kernels; ragged: the scalar kernels
@pytest.mark.parametrize("dim", [-1, 0, 1, 2])
def test_block_quantize_bit_exact(dim, shape):
    import mv_native as mv
    from oracle import quant_oracle as qo
    x = lognormal(shape[0] * shape[1] * shape[2], specials=False).reshape(shape)
    got = mv.block_quantize(x.cuda(), 8, dim).cpu().numpy()
    assert np.array_equal(bits(got), bits(qo.block_quantize(x.numpy(), 8, dim)))
    for wl in (8, 5):                       # 16-bit stream (wl >= 7) and 32-bit stream
        r = qo.philox_bits(x.numel(), 9, 2, half=wl >= 7)
        got = mv.block_quantize(x.cuda(), wl, dim, "stochastic", seed=9, offset=2).cpu().numpy()
        assert np.array_equal(bits(got), bits(qo.block_quantize(x.numpy(), wl, dim, "stochastic", r)))


def test_weight_quant_and_transpose():
    import mv_native as mv
    from oracle import quant_oracle as qo
    w = lognormal(1152 * 384, specials=False).reshape(1152, 384) * 1e-3
    q, qt = mv.quantize_weight(w.cuda(), 5, 10)
    want = qo.float_quantize(w.numpy(), 5, 10)
    assert np.array_equal(q.float().cpu().numpy(), want)
    assert np.array_equal(qt.float().cpu().numpy(), want.T)


def test_full_size_properties():
    """BASELINE-size input (2^28 elements): idempotence, fp16 representability, odd symmetry."""
    import mv_native as mv
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(n, device="cuda", generator=g) * torch.exp(
        torch.empty(n, device="cuda").uniform_(-12, 8, generator=g))
    q = mv.float_quantize(x, 5, 10)
    assert torch.equal(mv.float_quantize(q, 5, 10), q)
    assert torch.equal(q.half().float(), q)
    assert torch.equal(mv.float_quantize(-x, 5, 10), -q + 0.0)
    ref = x.half().float()                      # IEEE agrees except on ties / saturation
    assert ((q != ref) & (x.abs() < 65504)).float().mean().item() < 1e-3


def test_qtorch_facade_on_gpu():
    import qtorch
    import qtorch.quant as qq
    from oracle import quant_oracle as qo
    x = lognormal(5000)
    got = qq.float_quantize(x.cuda(), 5, 10, "nearest").cpu().numpy()
    assert np.array_equal(bits(got), bits(qo.float_quantize(x.numpy(), 5, 10)))
    quant = qq.Quantizer(qtorch.FloatingPoint(exp=5, man=10), forward_rounding="nearest")
    t = x.cuda().requires_grad_(True)
    y = quant(t)
    y.sum().backward()
    assert torch.equal(t.grad, torch.ones_like(t))           # identity backward (no backward_number)
    qq.manual_seed(5)
    a = qq.float_quantize(x.cuda(), 5, 10, "stochastic")
    qq.manual_seed(5)
    assert torch.equal(a, qq.float_quantize(x.cuda(), 5, 10, "stochastic"))
