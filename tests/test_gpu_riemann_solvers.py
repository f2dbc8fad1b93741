"""the opt-in `riemann=` switch (`[other] honourRiemannSolver=yes`).

  approx   riemann_approx + cmpflx — the reference's own (dead) functions, pinned function by function against its
           sources (tests/golden/kat.npz): whole runs must equal the oracle's run with the same solver BIT FOR BIT
  hll, rusanov   not in the reference (parity unpinned): bit-identical to the oracle's CPU restatement of the same
           published formulas, and checked on the properties SURVEY §8(c) prescribes — consistency, left/right mirror
           antisymmetry, and L1 convergence on Sod's shock tube against the exact solution
"""
import numpy as np
import pytest

import oracle
from euler2d_kokkos_b200 import HydroRun
from util import (INNER, assert_bitwise, both_params, gpu_eval, random_state, sod_exact_density, sod_initial_state)

pytestmark = pytest.mark.gpu
pytest.importorskip("torch")

SOLVERS = {"approx": oracle.RIEMANN_APPROX, "hll": oracle.RIEMANN_HLL, "rusanov": oracle.RIEMANN_RUSANOV,
           "hllc": oracle.RIEMANN_HLLC}
DECKS = [("implode", dict(mesh__nx=96, mesh__ny=64)), ("blast", dict(mesh__nx=64, mesh__ny=96)),
         ("four_quadrant", dict(mesh__nx=80, mesh__ny=80)), ("discontinuity", dict(mesh__nx=72, mesh__ny=56)),
         ("shocked_bubble", dict(mesh__nx=178, mesh__ny=37))]


@pytest.mark.parametrize("solver", ["hll", "rusanov"])
@pytest.mark.parametrize("gamma", ["1.4", "1.666"])
def test_extension_solvers_equal_their_cpu_restatement(solver, gamma):
    hp, op = both_params("implode", hydro__gamma0=gamma)
    rng = np.random.default_rng(99)
    n = 4096
    rec = np.concatenate([random_state(rng, n), random_state(rng, n)], axis=1)
    rec[::7, 2] += 8.0       # supersonic to the right / to the left
    rec[::7, 6] += 8.0
    rec[1::7, 2] -= 8.0
    rec[1::7, 6] -= 8.0
    rec[2::7, 4:8] = rec[2::7, 0:4]
    rec[3::7, 2] = 0.0
    rec[3::7, 6] = -0.0
    ref = {"hll": oracle.riemann_hll, "rusanov": oracle.riemann_rusanov}[solver](op, rec)
    assert_bitwise(gpu_eval(hp, solver, rec), ref, f"{solver} gamma={gamma}")


@pytest.mark.parametrize("unfused", ["no", "yes"])
@pytest.mark.parametrize("solver", ["approx", "hll", "rusanov"])
@pytest.mark.parametrize("deck,ov", DECKS)
def test_opt_in_solver_runs_equal_the_oracle_bitwise(deck, ov, solver, unfused):
    """five decks x three solvers x (fused kernel | the literal kernel sequence), host-driven like main.cpp:100-143"""
    hp, op = both_params(deck, hydro__riemann=solver, other__honourRiemannSolver="yes", other__unfusedKernels=unfused,
                         run__nOutput=-1, **ov)
    steps = 40
    with oracle.flux_solver(SOLVERS[solver]):
        U_ref, dts_ref, n_ref, t_ref = oracle.run(op, steps)
    with HydroRun(hp) as hydro:
        hydro.make_boundaries(HydroRun.U)
        hydro.make_boundaries(HydroRun.U2)
        t, dts = 0.0, []
        for n in range(steps):
            dt = hydro.compute_dt(n % 2)
            if t + dt > hp.tEnd:
                dt = hp.tEnd - t
            hydro.godunov_unsplit(n, dt)
            t += dt
            dts.append(dt)
        U = hydro.download(HydroRun.U if steps % 2 == 0 else HydroRun.U2)
    assert_bitwise(np.array(dts), dts_ref[1:], f"{deck}/{solver}: dt history")
    assert_bitwise(U[INNER], U_ref[INNER], f"{deck}/{solver} unfused={unfused}")


@pytest.mark.parametrize("solver", ["approx", "hll", "rusanov"])
def test_opt_in_solver_device_resident_loop_bitwise(solver):
    hp, op = both_params("implode", mesh__nx=128, mesh__ny=64, hydro__riemann=solver,
                         other__honourRiemannSolver="yes", run__nOutput=-1)
    with oracle.flux_solver(SOLVERS[solver]):
        U_ref, dts_ref, n_ref, t_ref = oracle.run(op, 80)
    with HydroRun(hp) as hydro:
        st = hydro.run(80)
        U = hydro.download(HydroRun.U if st.nStep % 2 == 0 else HydroRun.U2)
        dts = hydro.dt_history()
    assert st.nStep == n_ref and st.t == t_ref
    assert_bitwise(dts, dts_ref[1:], "dt history")
    assert_bitwise(U[INNER], U_ref[INNER], solver)


def test_switch_is_dead_without_the_opt_in():
    """reference behaviour: `riemann=` is parsed and never read (every kernel solves HLLC)"""
    out = {}
    for solver in ("approx", "hll", "rusanov", "hllc"):
        hp, _ = both_params("implode", mesh__nx=64, mesh__ny=32, hydro__riemann=solver, run__nOutput=-1)
        with HydroRun(hp) as hydro:
            st = hydro.run(20)
            out[solver] = hydro.download(HydroRun.U if st.nStep % 2 == 0 else HydroRun.U2)
    for solver in ("approx", "hll", "rusanov"):
        assert_bitwise(out[solver], out["hllc"], solver)


# ------------------------------------------------------------------ properties (SURVEY §8c for the unpinned solvers)
@pytest.mark.parametrize("solver", ["hll", "rusanov", "hllc", "approx"])
def test_consistency_with_the_physical_flux(solver):
    hp, op = both_params("implode", hydro__gamma0="1.4")
    q = random_state(np.random.default_rng(7), 1024)
    f = gpu_eval(hp, solver, np.concatenate([q, q], axis=1))[:, -4:]  # approx: qgdnv then flux
    np.testing.assert_allclose(f, oracle.cmpflx(op, q), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("solver", ["hll", "rusanov", "hllc"])
def test_left_right_mirror_antisymmetry(solver):
    """Mirroring the Riemann problem (swap the states, negate the normal velocities) negates the mass, energy and
    transverse-momentum fluxes and keeps the normal-momentum flux — exactly, since IEEE arithmetic is sign symmetric."""
    hp, _ = both_params("implode", hydro__gamma0="1.4")
    rng = np.random.default_rng(11)
    ql, qr = random_state(rng, 2048), random_state(rng, 2048)
    ql[::5, 2] += 6.0
    qr[1::5, 2] -= 6.0
    f = gpu_eval(hp, solver, np.concatenate([ql, qr], axis=1))
    ml, mr = qr.copy(), ql.copy()
    ml[:, 2] *= -1.0
    mr[:, 2] *= -1.0
    g = gpu_eval(hp, solver, np.concatenate([ml, mr], axis=1))
    # record layout (rho, E|p, normal, transverse)
    assert np.array_equal(g[:, 0], -f[:, 0]) and np.array_equal(g[:, 1], -f[:, 1])
    assert np.array_equal(g[:, 2], f[:, 2]) and np.array_equal(g[:, 3], -f[:, 3])


def sod_l1_error(solver, nx, t_end=0.2):
    hp, _ = both_params("four_quadrant", mesh__nx=nx, mesh__ny=4, mesh__xmin=0.0, mesh__xmax=1.0, mesh__ymin=0.0,
                        mesh__ymax=4.0 / nx, hydro__gamma0="1.4", hydro__riemann=solver,
                        other__honourRiemannSolver="yes", run__tEnd=t_end, run__nStepmax=100000, run__nOutput=-1)
    with HydroRun(hp) as hydro:
        U0 = sod_initial_state(hp.isize, hp.jsize, nx)
        hydro.upload(HydroRun.U, U0)
        hydro.upload(HydroRun.U2, U0)
        st = hydro.run()
        assert st.t == hp.tEnd
        U = hydro.download(HydroRun.U if st.nStep % 2 == 0 else HydroRun.U2)
    x = (np.arange(nx) + 0.5) / nx
    rho = U[0, 2:-2, 2:-2]
    assert np.abs(rho - rho[0]).max() < 1e-13, "the tube stays uniform in y"
    return np.abs(rho[0] - sod_exact_density(x, st.t, float(np.float32(1.4)))).mean()


@pytest.mark.parametrize("solver", ["hll", "rusanov", "hllc", "approx"])
def test_sod_shock_tube_converges(solver):
    e = [sod_l1_error(solver, nx) for nx in (100, 200, 400)]
    orders = [np.log2(e[k] / e[k + 1]) for k in range(2)]
    assert e[2] < e[1] < e[0] < 0.02, e
    assert min(orders) >= 0.8, (e, orders)  # a second-order scheme on a solution with discontinuities: order ~1 in L1
