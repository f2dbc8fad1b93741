"""The integer guards of the strict kernel's fast paths (csrc/e2d_lean.cuh), restated in numpy and checked as
IMPLICATIONS over adversarial doubles: whenever a guard accepts, the property the kernel relies on must hold.
(The GPU tests check the kernel's results bit for bit; these check the reasoning the guards rest on, on the CPU.)

  floor_guard    hi(x) > hi(f), f > 0                   =>  x > f            (so fmax(x, f) == x)
  pfloor_guard   hi(p) - hi(rho) >= gap(smallp)         =>  p > RN(rho * smallp)
  recip_of       (uint)(hi(d) - floor_hi - 1) < span    =>  floor < d < 2^1000
  div_by         |float(hi(q))| > float(kQuotLoHi)      <=> 2^-900 < |q| < 2^1017 (1 + 2^-20) (NaN / inf rejected)
  accepted numerator: |q| > 2^-900 and |d| > 2^-64      =>  |a| >= 2^-969 (nvcc's own acceptance condition)
"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "euler2d_kokkos_b200", "csrc", "e2d_lean.cuh")).read()


def const(name):
    m = re.search(rf"constexpr int {name} = \((\d+) ([+-]) (\d+)\) << 20;", SRC)
    assert m, name
    a, op, b = int(m.group(1)), m.group(2), int(m.group(3))
    return ((a + b) if op == "+" else (a - b)) << 20


K_DEN_LO, K_DEN_HI, K_QUOT_LO = const("kDenLoHi"), const("kDenHiHi"), const("kQuotLoHi")


def hi(x):
    return (np.asarray(x, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)


def adversarial(rng, n, emin=-1070, emax=1023, signed=True):
    mant = rng.uniform(1.0, 2.0, n)
    mant[::7] = 1.0
    mant[1::7] = np.nextafter(2.0, 0.0)
    mant[2::7] = np.nextafter(1.0, 2.0)
    x = np.ldexp(mant, rng.integers(emin, emax, n))
    if signed:
        x *= rng.choice([-1.0, 1.0], n)
    return x


def test_constants_are_the_documented_windows():
    assert K_DEN_LO == (1023 - 64) << 20 and K_DEN_HI == (1023 + 1000) << 20 and K_QUOT_LO == (1023 - 900) << 20
    # 2^-900 * 2^-64 = 2^-964 >= 2^-969: every accepted numerator satisfies nvcc's own condition
    assert -900 - 64 >= -969


def test_floor_guard_implies_x_above_floor():
    rng = np.random.default_rng(1)
    x = adversarial(rng, 200000)
    for f in (1e-10, float(np.float32(1e-10)), 1e-20, 3.7e-15, 1.0):
        ok = hi(x) > hi(f)
        assert (x[ok] > f).all()
        # and the guard is not vacuous: it accepts everything from 2 f upwards
        assert ok[x >= 2 * f].all()
    # values sharing the floor's high word are rejected either way (slow path), never mis-accepted
    f = float(np.float32(1e-10))
    near = np.nextafter(f, [0.0, 1.0])
    assert not (hi(near) > hi(f)).any()


def smallp_gap(smallp):
    es = ((int(hi(smallp)) >> 20) & 0x7ff) - 1023
    return (es + 4) * (1 << 20)


def test_pfloor_guard_implies_p_above_rho_smallp():
    rng = np.random.default_rng(2)
    for smallp in (1e-20 / 1.4, float(np.float32(1e-10)) ** 2 / float(np.float32(1.666)), 1e-7, 2.0 ** -40):
        gap = smallp_gap(smallp)
        rho = adversarial(rng, 300000, -60, 60, signed=False)
        p = adversarial(rng, 300000, -300, 300)
        # adversarial pairs right at the boundary p ~ rho * smallp * 2^k
        p[::5] = rho[::5] * smallp * np.ldexp(1.0, rng.integers(-2, 6, len(p[::5]))) * rng.uniform(0.9, 1.1, len(p[::5]))
        ok = (hi(p) - hi(rho)) >= gap
        assert (p[ok] > rho[ok] * smallp).all()
        assert ok[p > 64 * rho * smallp].all(), "ordinary pressures must pass"
        assert not ok[p <= 0].any()


def test_reciprocal_range_test():
    rng = np.random.default_rng(3)
    d = adversarial(rng, 300000)
    for floor in (float(np.float32(1e-10)), 2.0 ** -64):
        fh = int(hi(floor))
        lhs = ((hi(d) - fh - 1) & 0xFFFFFFFF).astype(np.uint64)  # the kernel's (unsigned)(hi - floor_hi - 1)
        ok = lhs < np.uint64((K_DEN_HI - fh - 1) & 0xFFFFFFFF)
        assert (d[ok] > floor).all() and (d[ok] < 2.0 ** 1000).all()
        assert ok[(d >= 2 * floor) & (d < 2.0 ** 999)].all()
        assert not ok[d <= 0].any()


def test_quotient_window_on_the_high_word_read_as_a_float():
    rng = np.random.default_rng(4)
    q = adversarial(rng, 300000)
    q[::11] = np.inf
    q[1::11] = np.nan
    q[2::11] = 0.0
    q[3::11] = np.ldexp(1.0, rng.integers(1015, 1024, len(q[3::11])))
    qh = hi(q).astype(np.int32).view(np.float32)
    with np.errstate(invalid="ignore"):
        good = np.abs(qh) > np.array([K_QUOT_LO], dtype=np.int32).view(np.float32)[0]
    a = np.abs(q)
    with np.errstate(invalid="ignore"):
        # the upper end: a high word of exactly 0x7f800000 reads as float +inf and passes, anything above reads as NaN and
        # fails — so the window closes within 2^-20 of 2^1017 (any normal quotient is fine for the sequence; the limit
        # that matters up there is the DENOMINATOR's, tested by recip_of)
        assert (a[good] > 2.0 ** -900).all() and (a[good] < 2.0 ** 1017 * (1 + 2.0 ** -19)).all()
        assert good[(a > 2.0 ** -899) & (a < 2.0 ** 1016)].all()
    assert not good[np.isnan(q) | np.isinf(q) | (q == 0)].any()


def test_accepted_numerators_satisfy_nvccs_condition():
    rng = np.random.default_rng(5)
    q = adversarial(rng, 200000, -899, 500)
    d = adversarial(rng, 200000, -63, 500)
    with np.errstate(over="ignore"):
        a = q * d  # the numerator the kernel divided, up to one rounding
    assert (np.abs(a[np.isfinite(a)]) >= 2.0 ** -969).all()
