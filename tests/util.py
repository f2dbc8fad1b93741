"""Shared helpers of the test-suite."""
from __future__ import annotations

import os
import tempfile

import numpy as np

import euler2d_kokkos_b200 as e2d
import oracle
from euler2d_kokkos_b200.decks import deck_text

INNER = (slice(None), slice(2, -2), slice(2, -2))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def both_params(name: str, **overrides):
    """(product HydroParams, oracle Params) for a deck with overrides (section__key=value)."""
    text = deck_text(name, **overrides)
    hp = e2d.HydroParams.from_string(text)
    with tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False) as f:
        f.write(text)
        path = f.name
    try:
        op = oracle.params_from_ini(path)
    finally:
        os.unlink(path)
    return hp, op


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint64)


def assert_bitwise(a: np.ndarray, b: np.ndarray, what: str = ""):
    """Exact comparison; on failure report where and by how much."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if np.array_equal(bits(a), bits(b)):
        return
    # +0 / -0 are equal values with different bits: report them separately
    neq = bits(a) != bits(b)
    val_neq = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    idx = np.argwhere(neq)
    k = tuple(idx[0])
    with np.errstate(all="ignore"):
        rel = np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    raise AssertionError(
        f"{what}: {neq.sum()} of {a.size} values differ bitwise ({val_neq.sum()} by value); first at {k}: "
        f"{a[k]!r} vs {b[k]!r} ({float(a[k]).hex()} vs {float(b[k]).hex()}); max rel diff {rel:.3e}")


def rel_errors(a: np.ndarray, ref: np.ndarray):
    """north_star's tolerance metric: relative L1 and L-inf per conserved variable, momenta normalised by a
    field scale so that fields that vanish by symmetry do not give noise/noise."""
    out = []
    mom_scale = max(np.abs(ref[2]).max(), np.abs(ref[3]).max(), 1e-300)
    for v in range(4):
        d = np.abs(a[v] - ref[v])
        if v < 2:
            l1 = d.sum() / np.abs(ref[v]).sum()
            linf = d.max() / np.abs(ref[v]).max()
        else:
            l1 = d.sum() / max(np.abs(ref[v]).sum(), mom_scale * ref[v].size * 1e-3)
            linf = d.max() / mom_scale
        out.append((l1, linf))
    return out


def random_state(rng, n, lo=0.1, hi=10.0):
    """n primitive states (rho, p, u, v) as in SURVEY.md §8(d): rho,p in U[0.1,10], u,v in U[-2,2]."""
    q = np.empty((n, 4))
    q[:, 0] = rng.uniform(lo, hi, n)
    q[:, 1] = rng.uniform(lo, hi, n)
    q[:, 2] = rng.uniform(-2, 2, n)
    q[:, 3] = rng.uniform(-2, 2, n)
    return q


def random_conservative_field(rng, op, jsize=None, smooth=False):
    """A random but physically valid conservative array [4][jsize][isize]."""
    jsize = jsize or op.jsize
    n = jsize * op.isize
    q = random_state(rng, n)
    rho, p, u, v = q.T
    U = np.empty((4, jsize, op.isize))
    U[0] = rho.reshape(jsize, -1)
    U[2] = (rho * u).reshape(jsize, -1)
    U[3] = (rho * v).reshape(jsize, -1)
    U[1] = (p / (op.gamma0 - 1.0) + 0.5 * rho * (u * u + v * v)).reshape(jsize, -1)
    return U


def gpu_eval(hp, func, rec):
    import ctypes as C

    nout = {"prim": 5, "slope": 8, "trace": 16, "hllc": 4, "approx": 8, "cmpflx": 4, "hll": 4, "hllc_lean": 5,
            "cell_lean": 6, "trace_lean": 25, "div": 5, "sqrt": 3, "fast_div": 2, "fast_sqrt": 2, "fast_hllc": 4,
            "fast_cell": 5, "fast_slope": 8, "fast_trace": 16, "rusanov": 4}[func]
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    n = rec.shape[0]
    out = np.zeros((n, nout))
    dp = C.POINTER(C.c_double)
    e2d.check(e2d.lib().e2d_k_eval_host(C.byref(hp.raw), func.encode(), rec.ctypes.data_as(dp),
                                        out.ctypes.data_as(dp), n), "e2d_k_eval_host")
    return out


# ------------------------------------------------------------------ Sod shock tube (property tests of the solvers)
def sod_exact_density(x, t, gamma=1.4, x0=0.5):
    """Density of the exact solution of Sod's problem (rho, u, p) = (1, 0, 1 | 0.125, 0, 0.1) at time t (Toro, ch. 4:
    left rarefaction, contact, right shock).  p* is found by Newton iteration on the pressure function."""
    rl, pl, rr, pr = 1.0, 1.0, 0.125, 0.1
    cl, cr = np.sqrt(gamma * pl / rl), np.sqrt(gamma * pr / rr)
    g1, g2 = (gamma - 1) / (2 * gamma), (gamma + 1) / (2 * gamma)

    def f(p, pk, rk, ck):
        if p > pk:  # shock
            A, B = 2 / ((gamma + 1) * rk), (gamma - 1) / (gamma + 1) * pk
            return (p - pk) * np.sqrt(A / (p + B)), np.sqrt(A / (p + B)) * (1 - (p - pk) / (2 * (B + p)))
        return 2 * ck / (gamma - 1) * ((p / pk) ** g1 - 1), 1 / (rk * ck) * (p / pk) ** (-g2)

    p = 0.5 * (pl + pr)
    for _ in range(50):
        fl, dfl = f(p, pl, rl, cl)
        fr, dfr = f(p, pr, rr, cr)
        dp = (fl + fr) / (dfl + dfr)
        p -= dp
        if abs(dp) < 1e-14:
            break
    fl, _ = f(p, pl, rl, cl)
    fr, _ = f(p, pr, rr, cr)
    us = 0.5 * (fr - fl)
    rsl = rl * (p / pl) ** (1 / gamma)                                   # behind the rarefaction
    rsr = rr * ((p / pr + (gamma - 1) / (gamma + 1)) / ((gamma - 1) / (gamma + 1) * p / pr + 1))  # behind the shock
    csl = cl * (p / pl) ** g1
    S = cr * np.sqrt(g2 * p / pr + g1)                                   # shock speed (ur = 0)
    xi = (x - x0) / t
    rho = np.where(xi < -cl, rl, 0.0)
    fan = (xi >= -cl) & (xi < us - csl)
    rho = np.where(fan, rl * (2 / (gamma + 1) - (gamma - 1) / ((gamma + 1) * cl) * xi) ** (2 / (gamma - 1)), rho)
    rho = np.where((xi >= us - csl) & (xi < us), rsl, rho)
    rho = np.where((xi >= us) & (xi < S), rsr, rho)
    rho = np.where(xi >= S, rr, rho)
    return rho


def sod_initial_state(isize, jsize, nx, gamma=1.4, x0=0.5):
    """Conservative array [4][jsize][isize] of Sod's problem on [0, 1], uniform in y (ghost cells included)."""
    x = (np.arange(isize) - 2 + 0.5) / nx
    left = x < x0
    U = np.zeros((4, jsize, isize))
    U[0] = np.where(left, 1.0, 0.125)[None, :]
    U[1] = (np.where(left, 1.0, 0.1) / (gamma - 1.0))[None, :]
    return U
