"""The C restatement of QPyTorch quant (oracle/quant_oracle.c) against the known-answer
vectors of SURVEY.md Appendix B and source-independent anchors (IEEE fp16 round trip away
from ties, idempotence, monotonicity, odd symmetry).  PARITY UNPINNED by the reference: it
holds no golden vectors for this boundary."""
import numpy as np
import pytest

from oracle import quant_oracle as qo


def bits(h):
    return np.array([int(h, 16)], dtype=np.uint32).view(np.float32)


KAT_5_10 = [
    ("80000000", "00000000"), ("3f801000", "3f802000"), ("3f803000", "3f804000"),
    ("bf801000", "bf802000"), ("477fe000", "477fe000"), ("477fefe6", "477fe000"),
    ("477ff000", "477fe000"), ("49742400", "477fe000"), ("7f800000", "477fe000"),
    ("c9742400", "c77fe000"), ("ff800000", "c77fe000"),
    ("38800000", "38800000"), ("38000000", "38000000"), ("33800000", "33800000"),
    ("33000000", "33800000"), ("b3000000", "b3800000"), ("32800000", "00000000"),
    ("3dcccccd", "3dccc000"), ("40490fdb", "40490000"), ("3a83126f", "3a832000"),
]
KAT_8_10 = [
    ("3f801000", "3f802000"), ("3f800800", "3f800000"), ("40490fdb", "40490000"),
    ("7f800000", "7f7fe000"), ("80000000", "00000000"),
]


@pytest.mark.parametrize("inp,out", KAT_5_10)
def test_kat_half(inp, out):
    r = qo.float_quantize(bits(inp), 5, 10).view(np.uint32)[0]
    assert "%08x" % r == out
    r = qo.float_quantize_nearest_np(bits(inp), 5, 10).view(np.uint32)[0]
    assert "%08x" % r == out


@pytest.mark.parametrize("inp,out", KAT_8_10)
def test_kat_tf32(inp, out):
    r = qo.float_quantize(bits(inp), 8, 10).view(np.uint32)[0]
    assert "%08x" % r == out


def test_tf32_large_and_denormal():
    r = qo.float_quantize(np.array([3.4e38, 1e-40], dtype=np.float32), 8, 10).view(np.uint32)
    assert ["%08x" % v for v in r] == ["7f7fc000", "00012000"]


def _lognormal(n, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) * np.exp(rng.uniform(-12, 8, n))).astype(np.float32)


def test_half_matches_ieee_except_ties_and_saturation():
    x = _lognormal(200000)
    q = qo.float_quantize(x, 5, 10)
    h = x.astype(np.float16).astype(np.float32)
    bad = q != h
    # every disagreement is an exact rounding tie, a saturation or the -0 case
    xb = x[bad]
    tie_or_sat = []
    for v in xb:
        if abs(v) >= 65520.0 or v == 0.0:
            tie_or_sat.append(True)
            continue
        e = max(np.floor(np.log2(abs(np.float64(v)))), -14)
        ulp = 2.0 ** (e - 10)
        frac = (abs(np.float64(v)) / ulp) % 1.0
        # (the subnormal branch rounds twice: fp32 add, then bit rounding -> near-ties count)
        tie_or_sat.append(abs(frac - 0.5) < 2.0 ** -12)
    assert all(tie_or_sat)
    assert bad.sum() < 200


@pytest.mark.parametrize("exp,man", [(5, 10), (8, 10), (4, 3), (5, 2), (8, 7), (6, 9)])
def test_idempotent_monotone_c_equals_numpy(exp, man):
    x = np.sort(_lognormal(50000, seed=exp * 31 + man))
    q = qo.float_quantize(x, exp, man)
    assert np.array_equal(q.view(np.uint32), qo.float_quantize(q, exp, man).view(np.uint32))
    assert np.all(np.diff(q) >= 0)
    assert np.array_equal(q.view(np.uint32),
                          qo.float_quantize_nearest_np(x, exp, man).view(np.uint32))
    # odd symmetry except -0 -> +0
    qn = qo.float_quantize(-x, exp, man)
    nz = q != 0
    assert np.array_equal(qn[nz], -q[nz])
    assert not np.signbit(qn[~nz]).any()


def test_half_outputs_are_fp16_representable():
    x = _lognormal(100000, seed=3)
    q = qo.float_quantize(x, 5, 10)
    assert np.array_equal(q.astype(np.float16).astype(np.float32), q)


def test_philox_16bit_stream_is_the_half_words_of_the_same_generator():
    n = 1000
    w = qo.philox_bits(n, seed=11, offset=2)                 # word (j & 3) of philox(j >> 2): philox(c) = w[4c : 4c + 4]
    h = qo.philox_bits(n, seed=11, offset=2, half=True)      # half-word (i & 7) of philox(i >> 3)
    for i in range(n // 2):
        word = int(w[4 * (i >> 3) + ((i & 7) >> 1)])
        assert int(h[i]) == ((word >> 16) if (i & 1) else (word & 0xFFFF))


def test_philox_known_answer():
    # Random123 KAT for philox4x32-10: counter = 0, key = 0
    r = qo.philox_bits(4, seed=0, offset=0)
    assert ["%08x" % v for v in r] == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]


def test_stochastic_rounding_is_unbiased_and_bounded():
    x = np.full(40000, 1.0 + 2.0 ** -12, dtype=np.float32)        # quarter of an fp16 ulp above 1
    r = qo.philox_bits(x.size, seed=7, offset=3)
    q = qo.float_quantize(x, 5, 10, "stochastic", r)
    assert set(np.unique(q)) <= {np.float32(1.0), np.float32(1.0 + 2.0 ** -10)}
    assert abs((q == np.float32(1.0 + 2.0 ** -10)).mean() - 0.25) < 0.01
    # zero random bits == truncation
    q0 = qo.float_quantize(x, 5, 10, "stochastic", np.zeros(x.size, np.uint32))
    assert np.all(q0 == 1.0)


def test_fixed_point():
    x = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 1023.4, 1e9, -1e9], dtype=np.float32) / 512
    q = qo.fixed_point_quantize(x, 11, 9)                           # CUDA rule: floor(x+0.5)
    assert np.array_equal(q * 512, np.array([1, 2, 3, 0, -1, 1023, 1023, -1024], np.float32))
    qo.set_switch("fixed_nearest_even", 1)
    try:
        qe = qo.fixed_point_quantize(x, 11, 9)
    finally:
        qo.set_switch("fixed_nearest_even", 0)
    assert np.array_equal(qe * 512, np.array([0, 2, 2, 0, -2, 1023, 1023, -1024], np.float32))
    qs = qo.fixed_point_quantize(x, 11, 9, symmetric=True)
    assert qs[-1] * 512 == -1023
    o, m = qo.fixed_point_quantize_mask(x, 11, 9)
    assert list(m) == [0, 0, 0, 0, 0, 0, 1, 1]
    u = qo.philox_uniform(1000, 5)
    assert u.min() >= 0 and u.max() < 1
    xs = np.full(1000, 0.25 / 512, np.float32)
    qst = qo.fixed_point_quantize(xs, 11, 9, rounding="stochastic", runif=u)
    assert set(np.unique(qst * 512)) <= {0.0, 1.0}
    assert abs((qst * 512).mean() - 0.25) < 0.05


def test_block_quantize():
    x = _lognormal(4096, seed=9).reshape(8, 16, 32)
    for dim in (-1, 0, 1, 2):
        q = qo.block_quantize(x, 8, dim)
        # grid spacing is 2^(e_max + 2 - wl) inside each block (base 6*2^e lies in [2^(e+2), 2^(e+3)))
        if dim < 0:
            blocks = [(x.ravel(), q.ravel())]
        else:
            xs, qs = np.moveaxis(x, dim, 0), np.moveaxis(q, dim, 0)
            blocks = [(xs[i].ravel(), qs[i].ravel()) for i in range(xs.shape[0])]
        for xb, qb in blocks:
            e = np.floor(np.log2(np.abs(xb).max()))
            step = 2.0 ** (e + 2 - 8)
            assert np.all(np.abs(qb - xb) <= step / 2 + 1e-30)
            assert np.all(np.abs(qb / step - np.round(qb / step)) == 0)
