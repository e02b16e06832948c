"""Device math routines (csrc/fastmath.cuh) against the host's libm / IEEE arithmetic.

divExact must be bit-identical to IEEE division (grad U / V feeds the bit-exact strain rate); the
table-driven log / exp / pow / asinh must stay within a few ulp, far inside the 1e-11 parity gate."""
import numpy as np
import pytest

from solidstateamfoam_b200 import capi

from helpers import ulp_diff

pytestmark = pytest.mark.gpu

LOG, EXP, POW, ASINH, DIVX, DIV, SELF, ASINH_EXP, DIVFAST = range(9)


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def rng():
    return np.random.default_rng(20261019)


def special():
    return np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308, 1e-300, 1e-150,
                     1e150, 1.7976931348623157e308, 2.0 ** -300, 2.0 ** 300, 2.0 ** -301, 2.0 ** 299, 3.0, 1e-15,
                     0.1, 1.0 - 2 ** -53, 1.0 + 2 ** -52])


def test_exact_division_is_ieee(ctx):
    r = rng()
    n = 2_000_000
    a = r.standard_normal(n) * 10.0 ** r.uniform(-30, 30, n)
    b = r.standard_normal(n) * 10.0 ** r.uniform(-30, 30, n)
    # mantissas of all ones / near powers of two are the classical hard cases for Newton division
    b[:1000] = np.nextafter(2.0 ** r.integers(-50, 50, 1000).astype(float), 0.0)
    a[:500] = np.nextafter(2.0 ** r.integers(-50, 50, 500).astype(float), np.inf)
    a[1000:1100] = 0.0
    with np.errstate(all="ignore"):
        ref = a / b
    got = ctx.debug_math(DIVX, a, b)
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.array_equal(np.signbit(got), np.signbit(ref))
    # special values: every pair
    s = special()
    A, B = np.meshgrid(s, s)
    A, B = A.ravel(), B.ravel()
    with np.errstate(all="ignore"):
        ref = A / B
    got = ctx.debug_math(DIVX, A, B)
    assert np.array_equal(got, ref, equal_nan=True)
    ok = ~np.isnan(ref)
    assert np.array_equal(np.signbit(got[ok]), np.signbit(ref[ok]))
    # and the plain operator, for completeness
    assert np.array_equal(ctx.debug_math(DIV, A, B), ref, equal_nan=True)


def test_fast_division(ctx):
    """divFast: <= 2 ulp inside the guarded range, IEEE semantics for special divisors"""
    r = rng()
    n = 1_000_000
    a = r.standard_normal(n) * 10.0 ** r.uniform(-100, 100, n)
    b = r.standard_normal(n) * 10.0 ** r.uniform(-100, 100, n)
    got = ctx.debug_math(DIVFAST, a, b)
    assert ulp_diff(got, a / b).max() <= 2
    s = special()
    A, B = np.meshgrid(s, s)
    A, B = A.ravel(), B.ravel()
    with np.errstate(all="ignore"):
        ref = A / B
    got = ctx.debug_math(DIVFAST, A, B)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    fin = np.isfinite(ref) & (np.abs(ref) > 1e-300)
    assert np.array_equal(got[~fin], ref[~fin], equal_nan=True) or np.allclose(got[~fin], ref[~fin], equal_nan=True)
    assert ulp_diff(got[fin], ref[fin]).max() <= 2


def test_self_ratio(ctx):
    s = special()
    with np.errstate(all="ignore"):
        ref = s / s
    assert np.array_equal(ctx.debug_math(SELF, s), ref, equal_nan=True)


def test_fast_log(ctx):
    r = rng()
    x = np.concatenate([10.0 ** r.uniform(-300, 300, 500_000), 1.0 + r.uniform(-0.3, 0.3, 500_000),
                        1.0 + r.uniform(-1e-6, 1e-6, 200_000), 2.0 ** r.integers(-1000, 1000, 2000).astype(float)])
    got = ctx.debug_math(LOG, x)
    ref = np.log(x)
    u = ulp_diff(got, ref)
    assert u.max() <= 4, (int(u.max()), x[int(np.argmax(u))])
    s = special()
    with np.errstate(all="ignore"):
        ref = np.log(s)
    got = ctx.debug_math(LOG, s)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~fin & ~np.isnan(ref)], ref[~fin & ~np.isnan(ref)])
    assert ulp_diff(got[fin], ref[fin]).max() <= 4


def test_fast_exp(ctx):
    r = rng()
    x = np.concatenate([r.uniform(-745, 710, 800_000), r.uniform(-1, 1, 300_000), r.uniform(-1e-8, 1e-8, 100_000)])
    got = ctx.debug_math(EXP, x)
    with np.errstate(all="ignore"):
        ref = np.exp(x)
    normal = ref > 2.3e-308
    u = ulp_diff(got[normal], ref[normal])
    assert u.max() <= 4, (int(u.max()), x[normal][int(np.argmax(u))])
    assert np.allclose(got[~normal], ref[~normal], rtol=1e-10, atol=1e-320)
    s = special()
    with np.errstate(all="ignore"):
        ref = np.exp(s)
    got = ctx.debug_math(EXP, s)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.allclose(got[ok], ref[ok], rtol=1e-15, atol=0)


def test_fast_pow(ctx):
    r = rng()
    n = 600_000
    x = 10.0 ** r.uniform(-30, 30, n)
    y = r.uniform(-5, 5, n)
    got = ctx.debug_math(POW, x, y)
    ref = np.power(x, y)
    rel = np.abs(got - ref) / ref
    # relative error ~ |y log x| * 2^-52: up to ~350 * 2.2e-16
    assert rel.max() <= 2e-13, float(rel.max())
    # the exponents of the configured models (1/n = 0.2817, m-1 = -0.8, 1/c2, b2)
    for yy in (1 / 3.55, -0.8, 0.7, 0.4, 2.0, 1.0):
        got = ctx.debug_math(POW, x, np.full(n, yy))
        ref = np.power(x, yy)
        assert (np.abs(got - ref) / ref).max() <= 5e-14
    # reference semantics at the edges: pow(0, y>0) = 0, pow(x<0, non-integer) = NaN
    # (the exponent is always a finite model parameter, so pow(1, NaN) = 1 is not reproduced)
    e = ctx.debug_math(POW, np.array([0.0, -1.0, np.inf, np.nan, 0.0]), np.array([0.3, 0.5, 0.5, 0.5, -0.8]))
    assert e[0] == 0.0 and np.isnan(e[1]) and e[2] == np.inf and np.isnan(e[3]) and e[4] == np.inf


def test_fast_asinh(ctx):
    r = rng()
    x = np.concatenate([10.0 ** r.uniform(-40, 40, 600_000), -(10.0 ** r.uniform(-10, 10, 100_000)),
                        r.uniform(0.0, 0.05, 200_000), 2.0 ** r.uniform(27, 29, 10_000)])
    got = ctx.debug_math(ASINH, x)
    ref = np.arcsinh(x)
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= 1e-14, (float(rel.max()), x[int(np.argmax(rel))])
    s = special()
    got = ctx.debug_math(ASINH, s)
    ref = np.arcsinh(s)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.allclose(got[ok], ref[ok], rtol=1e-14, atol=0)
    # asinh(exp(t)) as used by SheppardWright in the log domain
    t = r.uniform(-200, 200, 300_000)
    got = ctx.debug_math(ASINH_EXP, t)
    import mpmath as mp
    idx = r.integers(0, t.size, 300)
    for i in idx:
        ref_i = float(mp.asinh(mp.exp(mp.mpf(float(t[i])))))
        assert abs(got[i] - ref_i) <= 2e-14 * abs(ref_i), (t[i], got[i], ref_i)
    assert got[np.argmin(t)] > 0
