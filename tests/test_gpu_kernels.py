"""GPU parity tests, kernel level: every e2d_k_* entry point of the C ABI against the CPU oracle on the same
seeded inputs.  Tolerance: NONE — the strict build (-fmad=false) must be bit-identical for IEEE doubles."""
import ctypes as C

import numpy as np
import pytest

import euler2d_kokkos_b200 as e2d
import oracle
from util import INNER, assert_bitwise, both_params, gpu_eval, random_conservative_field, random_state

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def ptr(t):
    return C.c_void_p(t.data_ptr())


L = e2d.lib
ck = e2d.check


# ------------------------------------------------------------------ function-level KATs
def kat_params(gamma="1.4"):
    # gamma0 goes through float like every real parameter; the oracle sees the same value
    return both_params("implode", hydro__gamma0=gamma, hydro__slope_type="2")


@pytest.mark.parametrize("func", ["prim", "slope", "trace", "hllc", "approx", "cmpflx"])
@pytest.mark.parametrize("gamma", ["1.4", "1.666"])
def test_device_functions_match_oracle(func, gamma):
    hp, op = kat_params(gamma)
    rng = np.random.default_rng(1234)
    n = 4096
    if func == "prim":
        q = random_state(rng, n)
        rec = np.stack([q[:, 0], q[:, 1] / 0.4 + 0.5 * q[:, 0] * (q[:, 2] ** 2 + q[:, 3] ** 2), q[:, 0] * q[:, 2],
                        q[:, 0] * q[:, 3]], axis=1)
        rec[::97, 0] = 1e-12  # below smallr: exercises the density floor
        qo, co = oracle.compute_primitives(op, rec)
        ref = np.concatenate([qo, co[:, None]], axis=1)
    elif func == "slope":
        rec = np.concatenate([random_state(rng, n) for _ in range(5)], axis=1)
        rec[::5, 4:8] = rec[::5, 0:4]  # flat on one side -> zero slopes
        ref = oracle.slopes(op, rec)
    elif func == "trace":
        rec = np.concatenate([random_state(rng, n), rng.normal(0, 0.3, (n, 8)), rng.uniform(0.05, 0.5, (n, 2))], axis=1)
        ref = oracle.trace(op, rec)
    elif func in ("hllc", "approx"):
        rec = np.concatenate([random_state(rng, n), random_state(rng, n)], axis=1)
        rec[::7, 2] += 8.0   # supersonic to the right on both sides
        rec[::7, 6] += 8.0
        rec[1::7, 2] -= 8.0  # supersonic to the left
        rec[1::7, 6] -= 8.0
        rec[2::7, 4:8] = rec[2::7, 0:4]  # identical states
        ref = oracle.riemann_hllc(op, rec) if func == "hllc" else oracle.riemann_approx(op, rec)
    else:
        rec = random_state(rng, n)
        ref = oracle.cmpflx(op, rec)
    out = gpu_eval(hp, func, rec)
    assert_bitwise(out, ref, f"{func} gamma={gamma}")


def test_known_answers_of_the_reference():
    """values produced by the reference's own HydroBaseFunctor methods
    (those literals were computed with gamma0 = 1.4 as a DOUBLE, which cannot be set through the float-parsing .ini —
    1.4f differs from 1.4 by 2.4e-8 relative — hence rtol 2e-7 here and only here.  The bit-exact versions of this
    check are test_device_functions_match_oracle above, against golden vectors produced by the reference's own
    sources with the same float-parsed gamma, and test_oracle_pins.py on the CPU)."""
    hp, _ = kat_params("1.4")
    ql, qr = [1, 1, 0.1, 0.2], [0.125, 0.1, -0.1, 0.3]
    flux = gpu_eval(hp, "hllc", np.array([ql + qr]))[0]
    np.testing.assert_allclose(flux, [0.4679818632378312, 1.3156618140470782, 0.53779980134648486,
                                      0.093596372647566248], rtol=2e-7)
    out = gpu_eval(hp, "approx", np.array([ql + qr]))[0]
    np.testing.assert_allclose(out[4:], [0.45346810650719527, 1.3382030475873425, 0.76925064462513149,
                                         0.090693621301439056], rtol=2e-7)
    np.testing.assert_allclose(out[:4], [0.48103826852028286, 0.34177254760851022, 0.94268613576650373, 0.2], rtol=2e-7)
    q = [1, 1, .1, .2]
    rec = np.array([q + [.9, .8, .15, .1] + [1.2, 1.3, 0, .25] + [1.1, .9, .12, .3] + [.7, 1.05, .05, .15]])
    dq = gpu_eval(hp, "slope", rec)[0]
    np.testing.assert_allclose(dq, [-0.15, -0.25, 0.075, -0.075, 0.2, -0.075, 0.035, 0.075], rtol=1e-14)
    tr = gpu_eval(hp, "trace", np.array([q + list(dq) + [0.3, 0.2]]))[0]
    np.testing.assert_allclose(tr[:4], [1.04875, 1.0995, 0.097825, 0.247625], rtol=2e-7)


def test_hll_extension_is_consistent():
    """HLL is not in the reference (parity unpinned): F(q,q) must equal the physical flux, and for states
    supersonic to the right it must equal HLLC's upwind flux."""
    hp, op = kat_params("1.4")
    rng = np.random.default_rng(7)
    q = random_state(rng, 512)
    f = gpu_eval(hp, "hll", np.concatenate([q, q], axis=1))
    phys = oracle.cmpflx(op, q)
    np.testing.assert_allclose(f, phys, rtol=1e-12, atol=1e-12)
    ql, qr = random_state(rng, 512), random_state(rng, 512)
    ql[:, 2] += 12
    qr[:, 2] += 12
    rec = np.concatenate([ql, qr], axis=1)
    np.testing.assert_allclose(gpu_eval(hp, "hll", rec), gpu_eval(hp, "hllc", rec), rtol=1e-13)


# ------------------------------------------------------------------ the fused kernel's lean-but-exact math
def adversarial_doubles(rng, n):
    """Doubles spread over the whole exponent range, both signs, with zeros, subnormals and hard mantissas."""
    mant = rng.uniform(1.0, 2.0, n)
    mant[::11] = np.nextafter(2.0, 0.0)   # all-ones mantissa
    mant[1::11] = 1.0
    mant[2::11] = np.nextafter(1.0, 2.0)
    expo = rng.integers(-1070, 1023, n)
    x = np.ldexp(mant, expo) * rng.choice([-1.0, 1.0], n)
    x[3::17] = 0.0
    x[4::17] = -0.0
    return x


def test_shared_reciprocal_division_is_correctly_rounded():
    """div_by (csrc/e2d_lean.cuh): wherever its guard accepts the fast path the quotient is bit-identical to
    the IEEE `/` (and to numpy's), over the whole exponent range, signed zeros included."""
    hp, _ = kat_params("1.4")
    rng = np.random.default_rng(99)
    n = 1 << 18
    a, d = adversarial_doubles(rng, n), adversarial_doubles(rng, n)
    # a moderate-range block like the solver sees (most of these must stay on the fast path)
    a[: n // 2] = rng.normal(0, 3, n // 2) * 10.0 ** rng.integers(-12, 6, n // 2)
    d[: n // 2] = np.abs(rng.normal(0, 3, n // 2)) * 10.0 ** rng.integers(-9, 6, n // 2) + 1e-300
    a[5 : n // 2 : 13] = 0.0
    a[6 : n // 2 : 13] = -0.0
    out = gpu_eval(hp, "div", np.stack([a, d], axis=1))
    with np.errstate(all="ignore"):
        ref = a / d
    assert_bitwise(out[:, 2], ref, "device `/` vs numpy")  # sanity of the comparison itself
    for col, gcol, what in ((0, 1, "zero-ok"), (3, 4, "plain")):
        ok = out[:, gcol] == 1.0
        assert_bitwise(out[ok, col], ref[ok], f"shared-reciprocal quotient ({what})")
    okz, okp = out[:, 1] == 1.0, out[:, 4] == 1.0
    assert okz[: n // 2].mean() > 0.99, "moderate-range quotients must stay on the fast path"
    zero_num = (a == 0) & (d >= 2.3e-308) & np.isfinite(d)
    assert okz[zero_num].all(), "zero numerators over positive denominators stay on the fast path"
    assert not okp[(a == 0) & np.isfinite(d) & (d != 0)].any(), "without ZERO_OK a zero numerator must be rejected"
    assert not okz[(d <= 0)].any(), "ZERO_OK requires a positive denominator"


def test_fast_path_sqrt_is_correctly_rounded():
    hp, _ = kat_params("1.4")
    rng = np.random.default_rng(5)
    n = 1 << 18
    x = np.abs(adversarial_doubles(rng, n))
    x[: n // 2] = rng.uniform(0, 4, n // 2) * 10.0 ** rng.integers(-20, 20, n // 2)
    # squares of random doubles +- 1ulp: results next to rounding boundaries
    r = rng.uniform(1, 2, n // 4)
    x[n // 2 : n // 2 + n // 4] = np.nextafter(r * r, rng.choice([0.0, 10.0], n // 4))
    out = gpu_eval(hp, "sqrt", x[:, None])
    ref = np.sqrt(x)
    assert_bitwise(out[:, 2], ref, "device sqrt() vs numpy")
    ok = out[:, 1] == 1.0
    assert_bitwise(out[ok, 0], ref[ok], "fast-path sqrt")
    assert ok[: n // 2][x[: n // 2] > 1e-250].all()


@pytest.mark.parametrize("gamma", ["1.4", "1.666"])
def test_lean_device_functions_match_oracle(gamma):
    """hllc_lean / prim_lean / cfl_lean / slope_lean / trace_sources_lean == the reference formulas, bit for bit,
    including states at rest (zero numerators), identical states and supersonic states."""
    hp, op = kat_params(gamma)
    rng = np.random.default_rng(4321)
    n = 8192
    rec = np.concatenate([random_state(rng, n), random_state(rng, n)], axis=1)
    rec[::7, 2] += 8.0
    rec[::7, 6] += 8.0
    rec[1::7, 2] -= 8.0
    rec[1::7, 6] -= 8.0
    rec[2::7, 4:8] = rec[2::7, 0:4]      # identical states
    rec[3::7, 2:4] = 0.0                 # gas at rest on both sides, equal pressure: ustar numerator is exactly 0
    rec[3::7, 6:8] = 0.0
    rec[3::7, 5] = rec[3::7, 1]
    rec[4::7, 2] = -0.0                  # signed zero normal velocity
    rec[5::70, 0] = 1e-300               # absurd states: must fall back, still exact
    rec[6::70, 1] = 1e-280
    out = gpu_eval(hp, "hllc_lean", rec)
    assert_bitwise(out[:, :4], oracle.riemann_hllc(op, rec), f"hllc_lean gamma={gamma}")
    assert (out[:, 4] == 1.0).mean() > 0.95, "ordinary states must stay on the fast path"

    q = random_state(rng, n)
    u = np.stack([q[:, 0], q[:, 1] / 0.4 + 0.5 * q[:, 0] * (q[:, 2] ** 2 + q[:, 3] ** 2), q[:, 0] * q[:, 2],
                  q[:, 0] * q[:, 3]], axis=1)
    u[::5, 2] = 0.0
    u[1::5, 3] = -0.0
    u[::97, 0] = 1e-12
    u[7::500, 2] = 1e-300
    out = gpu_eval(hp, "cell_lean", u)
    qo, co = oracle.compute_primitives(op, u)
    assert_bitwise(out[:, :4], qo, f"prim_lean gamma={gamma}")
    inv = (co + np.abs(qo[:, 2])) / op.dx + (co + np.abs(qo[:, 3])) / op.dy
    assert_bitwise(out[:, 4], inv, f"cfl_lean gamma={gamma}")
    # the fast path is taken by every ordinary record; a density below the floor (smallr) and a momentum so small that
    # the velocity leaves the guard window (|q| > 2^-900) are recomputed with the plain operators — exact either way
    ordinary = np.ones(n, dtype=bool)
    ordinary[::97] = False
    ordinary[7::500] = False
    assert (out[ordinary, 5] == 1.0).all()
    assert (out[::97, 5] == 0.0).all(), "a density below smallr must leave the fast path (the floor is part of the guard)"

    st = np.concatenate([random_state(rng, n) for _ in range(5)], axis=1)
    st[::5, 4:8] = st[::5, 0:4]          # flat on one side -> zero slopes -> zero numerators in the trace
    st[1::5, 4:20] = np.tile(st[1::5, 0:4], 4)   # completely flat
    st[2::50, 2] = -0.0
    dts = rng.uniform(0.05, 0.5, (n, 2))
    out = gpu_eval(hp, "trace_lean", np.concatenate([st, dts], axis=1))
    dq = oracle.slopes(op, st)
    assert_bitwise(out[:, :8], dq, f"slope_lean gamma={gamma}")
    ref = oracle.trace(op, np.concatenate([st[:, :4], dq, dts], axis=1))
    assert_bitwise(out[:, 8:24], ref, f"trace_lean gamma={gamma}")
    assert (out[:, 24] == 1.0).mean() > 0.99


# ------------------------------------------------------------------ boundary fill: bit-exact
BC_CASES = [(1, 1, 1, 1), (2, 2, 2, 2), (3, 3, 3, 3), (3, 3, 2, 1), (1, 2, 3, 3), (2, 1, 1, 2)]


@pytest.mark.parametrize("bcs", BC_CASES)
@pytest.mark.parametrize("shape", [(37, 23), (128, 5), (7, 64)])
def test_make_boundaries_bit_exact(bcs, shape):
    nx, ny = shape
    hp, op = both_params("implode", mesh__nx=nx, mesh__ny=ny, mesh__boundary_type_xmin=bcs[0],
                         mesh__boundary_type_xmax=bcs[1], mesh__boundary_type_ymin=bcs[2],
                         mesh__boundary_type_ymax=bcs[3])
    rng = np.random.default_rng(nx * 100 + ny)
    U = rng.normal(size=(4, op.jsize, op.isize))
    U[2:, 3:6, 3:6] = 0.0  # zeros pick up a sign at reflecting walls (-0.0), must match too
    ref = U.copy()
    oracle.make_boundaries(op, ref)
    d = dev(U)
    ck(L().e2d_k_make_boundaries(C.byref(hp.raw), ptr(d), op.jsize, e2d.FACES_ALL, None))
    assert_bitwise(host(d), ref, f"make_boundaries {bcs}")


@pytest.mark.parametrize("faces,do_ymin,do_ymax", [(e2d.FACES_X, 0, 0), (e2d.FACES_X | e2d.FACES_YMIN, 1, 0),
                                                   (e2d.FACES_X | e2d.FACES_YMAX, 0, 1)])
def test_make_boundaries_slab_masks(faces, do_ymin, do_ymax):
    hp, op = both_params("four_quadrant", mesh__nx=40, mesh__ny=64)
    rng = np.random.default_rng(5)
    jsize_loc = 21
    U = rng.normal(size=(4, jsize_loc, op.isize))
    ref = U.copy()
    oracle.make_boundaries(op, ref, bool(do_ymin), bool(do_ymax))
    d = dev(U)
    ck(L().e2d_k_make_boundaries(C.byref(hp.raw), ptr(d), jsize_loc, faces, None))
    assert_bitwise(host(d), ref, "slab boundaries")


# ------------------------------------------------------------------ array kernels
@pytest.fixture(scope="module")
def field():
    hp, op = both_params("blast", mesh__nx=150, mesh__ny=70)
    rng = np.random.default_rng(99)
    U = random_conservative_field(rng, op)
    oracle.make_boundaries(op, U)
    return hp, op, U


def test_reduce_invdt_exact(field):
    hp, op, U = field
    d = dev(U)
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    ck(L().e2d_k_reduce_invdt(C.byref(hp.raw), ptr(d), op.jsize, ptr(acc), None))
    assert host(acc)[0] == oracle.compute_invdt(op, U)


def test_convert_to_primitives_exact(field):
    hp, op, U = field
    d, q = dev(U), dev(np.zeros_like(U))
    ck(L().e2d_k_convert_to_primitives(C.byref(hp.raw), ptr(d), ptr(q), op.jsize, None))
    assert_bitwise(host(q), oracle.convert_to_primitives(op, U), "Q")


def test_unfused_pipeline_exact(field):
    hp, op, U = field
    dt = op.cfl / oracle.compute_invdt(op, U)
    dtdx, dtdy = dt / op.dx, dt / op.dy
    Q = oracle.convert_to_primitives(op, U)
    Fx, Fy = oracle.compute_and_store_fluxes(op, Q, dtdx, dtdy)
    dQ, dFx, dFy = dev(Q), dev(np.zeros_like(U)), dev(np.zeros_like(U))
    ck(L().e2d_k_compute_and_store_fluxes(C.byref(hp.raw), ptr(dQ), ptr(dFx), ptr(dFy), dtdx, dtdy, op.jsize, None))
    assert_bitwise(host(dFx), Fx, "Fx")
    assert_bitwise(host(dFy), Fy, "Fy")
    ref = U.copy()
    oracle.update(op, ref, Fx, Fy)
    dU = dev(U)
    ck(L().e2d_k_update(C.byref(hp.raw), ptr(dU), ptr(dFx), ptr(dFy), op.jsize, None))
    assert_bitwise(host(dU), ref, "update")


def test_implementation1_trio_equals_implementation0(field):
    hp, op, U = field
    dt = op.cfl / oracle.compute_invdt(op, U)
    dtdx, dtdy = dt / op.dx, dt / op.dy
    ref = oracle.godunov(op, U, dt)
    Q = oracle.convert_to_primitives(op, U)
    dQ, dSx, dSy, dF, dU = dev(Q), dev(np.zeros_like(U)), dev(np.zeros_like(U)), dev(np.zeros_like(U)), dev(U)
    p = C.byref(hp.raw)
    ck(L().e2d_k_compute_slopes(p, ptr(dQ), ptr(dSx), ptr(dSy), op.jsize, None))
    for direction in (1, 2):
        ck(L().e2d_k_compute_trace_and_fluxes(p, ptr(dQ), ptr(dSx), ptr(dSy), ptr(dF), dtdx, dtdy, direction,
                                              op.jsize, None))
        ck(L().e2d_k_update_dir(p, ptr(dU), ptr(dF), direction, op.jsize, None))
    assert_bitwise(host(dU)[INNER], ref[INNER], "impl 1")


@pytest.mark.parametrize("deck,nx,ny", [("blast", 150, 70), ("implode", 124, 33), ("four_quadrant", 125, 200),
                                        ("shocked_bubble", 380, 11), ("implode", 2, 2), ("implode", 3, 700)])
def test_fused_step_exact(deck, nx, ny):
    hp, op = both_params(deck, mesh__nx=nx, mesh__ny=ny)
    rng = np.random.default_rng(nx + ny)
    U = random_conservative_field(rng, op)
    oracle.make_boundaries(op, U)
    dt = op.cfl / oracle.compute_invdt(op, U)
    ref = oracle.godunov(op, U, dt)
    dU, dO = dev(U), dev(np.full_like(U, np.nan))
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    ck(L().e2d_k_fused_step(C.byref(hp.raw), ptr(dU), ptr(dO), op.jsize, dt, None, ptr(acc), None, None))
    out = host(dO)
    assert_bitwise(out[INNER], ref[INNER], f"fused step {deck} {nx}x{ny}")
    assert host(acc)[0] == oracle.compute_invdt(op, ref), "fused CFL reduction"
    # the kernel must not touch ghost cells of the output
    assert np.isnan(out[:, :2]).all() and np.isnan(out[:, -2:]).all()
    assert np.isnan(out[:, :, :2]).all() and np.isnan(out[:, :, -2:]).all()
    # device-resident dt gives the same result
    ddt = dev(np.array([dt]))
    dO2 = dev(np.full_like(U, np.nan))
    ck(L().e2d_k_fused_step(C.byref(hp.raw), ptr(dU), ptr(dO2), op.jsize, 0.0, ptr(ddt), None, None, None))
    assert_bitwise(host(dO2)[INNER], ref[INNER], "fused step, dt from device memory")
    # *d_skip != 0 turns the launch into a no-op
    skip = torch.ones(1, dtype=torch.int32, device="cuda")
    dO3 = dev(np.full_like(U, np.nan))
    ck(L().e2d_k_fused_step(C.byref(hp.raw), ptr(dU), ptr(dO3), op.jsize, dt, None, None, ptr(skip), None))
    assert np.isnan(host(dO3)).all()


def test_fused_step_rejects_in_place(field):
    hp, op, U = field
    d = dev(U)
    rc = L().e2d_k_fused_step(C.byref(hp.raw), ptr(d), ptr(d), op.jsize, 1e-3, None, None, None, None)
    assert rc == 1 and b"out of place" in L().e2d_last_error()


@pytest.mark.parametrize("deck", ["implode", "blast", "four_quadrant", "discontinuity", "shocked_bubble"])
def test_init_problem_exact(deck):
    hp, op = both_params(deck, mesh__nx=61, mesh__ny=47)
    d = dev(np.zeros((4, op.jsize, op.isize)))
    ck(L().e2d_k_init_problem(C.byref(hp.raw), ptr(d), op.jsize, 0, None))
    assert_bitwise(host(d), oracle.init_slab(op), f"init {deck}")
    # a slab of the same problem: rows [j_off, j_off+20)
    d2 = dev(np.zeros((4, 20, op.isize)))
    ck(L().e2d_k_init_problem(C.byref(hp.raw), ptr(d2), 20, 13, None))
    assert_bitwise(host(d2), oracle.init_slab(op, 20, 13), f"init slab {deck}")


def test_sedov_energy_renormalised_init():
    hp, op = both_params("sedov_blast_2d", mesh__nx=128, mesh__ny=128)
    d = dev(np.zeros((4, op.jsize, op.isize)))
    ck(L().e2d_k_init_problem(C.byref(hp.raw), ptr(d), op.jsize, 0, None))
    assert_bitwise(host(d), oracle.init_slab(op), "sedov init")


def test_sedov_init_of_a_slab_is_refused_by_the_stateless_entry_point():
    """the disc energy needs a count over the whole grid: handles do it (e2d_blast_*), e2d_k_init_problem cannot"""
    hp, op = both_params("sedov_blast_2d", mesh__nx=64, mesh__ny=64)
    d = dev(np.zeros((4, 20, op.isize)))
    assert L().e2d_k_init_problem(C.byref(hp.raw), ptr(d), 20, 13, None) == 5  # E2D_ERR_UNSUPPORTED
