"""CPU tests of the oracle's opt-in flux solvers and of the oracle chain itself.

approx is the reference's own riemann_approx + cmpflx (pinned per function in test_oracle_pins.py); HLL and Rusanov do
not exist in the reference, so their restatements are checked on properties: consistency, mirror antisymmetry and
convergence on Sod's shock tube.  The GPU tests (test_gpu_riemann_solvers.py) then pin the product to these bit for bit.
"""
import os
import sys
import tempfile

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle
from euler2d_kokkos_b200.decks import deck_text, write_deck
from util import random_state, sod_exact_density, sod_initial_state


def oracle_params(name, **ov):
    with tempfile.TemporaryDirectory() as td:
        return oracle.params_from_ini(write_deck(os.path.join(td, "d.ini"), name, **ov))


FLUX = {"hll": oracle.riemann_hll, "rusanov": oracle.riemann_rusanov, "hllc": oracle.riemann_hllc}
SOLVER_ID = {"approx": oracle.RIEMANN_APPROX, "hll": oracle.RIEMANN_HLL, "rusanov": oracle.RIEMANN_RUSANOV,
             "hllc": oracle.RIEMANN_HLLC}


@pytest.mark.parametrize("solver", ["hll", "rusanov", "hllc"])
def test_consistency(solver):
    op = oracle_params("implode", hydro__gamma0="1.4")
    q = random_state(np.random.default_rng(3), 256)
    np.testing.assert_allclose(FLUX[solver](op, np.concatenate([q, q], axis=1)), oracle.cmpflx(op, q), rtol=1e-12,
                               atol=1e-12)


@pytest.mark.parametrize("solver", ["hll", "rusanov", "hllc"])
def test_mirror_antisymmetry(solver):
    op = oracle_params("implode", hydro__gamma0="1.4")
    rng = np.random.default_rng(5)
    ql, qr = random_state(rng, 512), random_state(rng, 512)
    ql[::5, 2] += 6.0
    qr[1::5, 2] -= 6.0
    f = FLUX[solver](op, np.concatenate([ql, qr], axis=1))
    ml, mr = qr.copy(), ql.copy()
    ml[:, 2] *= -1.0
    mr[:, 2] *= -1.0
    g = FLUX[solver](op, np.concatenate([ml, mr], axis=1))
    assert np.array_equal(g[:, 0], -f[:, 0]) and np.array_equal(g[:, 1], -f[:, 1])
    assert np.array_equal(g[:, 2], f[:, 2]) and np.array_equal(g[:, 3], -f[:, 3])


def test_rusanov_is_the_most_diffusive_and_upwinds_supersonic_hll():
    op = oracle_params("implode", hydro__gamma0="1.4")
    rng = np.random.default_rng(9)
    ql, qr = random_state(rng, 256), random_state(rng, 256)
    ql[:, 2] += 12.0
    qr[:, 2] += 12.0
    rec = np.concatenate([ql, qr], axis=1)
    np.testing.assert_allclose(oracle.riemann_hll(op, rec), oracle.cmpflx(op, ql), rtol=1e-13)  # SL >= 0: upwind flux
    np.testing.assert_allclose(oracle.riemann_hllc(op, rec), oracle.cmpflx(op, ql), rtol=1e-13)
    # Rusanov adds -smax/2 (UR - UL) even there: its mass flux differs from the upwind one by exactly that term
    f = oracle.riemann_rusanov(op, rec)
    assert not np.allclose(f[:, 0], oracle.cmpflx(op, ql)[:, 0], rtol=1e-6)


def sod_l1_error_oracle(solver, nx, t_end=0.2):
    op = oracle_params("four_quadrant", mesh__nx=nx, mesh__ny=4, mesh__xmin=0.0, mesh__xmax=1.0, mesh__ymin=0.0,
                       mesh__ymax=4.0 / nx, hydro__gamma0="1.4", run__tEnd=t_end, run__nOutput=-1)
    U = sod_initial_state(op.isize, op.jsize, nx, op.gamma0)
    t = 0.0
    with oracle.flux_solver(SOLVER_ID[solver]):
        while t < t_end:
            oracle.make_boundaries(op, U)
            dt = op.cfl / oracle.compute_invdt(op, U)
            if t + dt > t_end:
                dt = t_end - t
            U = oracle.godunov(op, U, dt)
            t += dt
    x = (np.arange(nx) + 0.5) / nx
    return np.abs(U[0, 2, 2:-2] - sod_exact_density(x, t, op.gamma0)).mean()


@pytest.mark.parametrize("solver", ["approx", "hll", "rusanov", "hllc"])
def test_sod_shock_tube_converges(solver):
    e = [sod_l1_error_oracle(solver, nx) for nx in (100, 200, 400)]
    orders = [np.log2(e[k] / e[k + 1]) for k in range(2)]
    assert e[2] < e[1] < e[0] < 0.02, e
    assert min(orders) >= 0.8, (e, orders)


def test_flux_solver_switch_restores_the_reference_default():
    op = oracle_params("implode", mesh__nx=32, mesh__ny=24, run__nOutput=-1)
    base = oracle.run(op, 10)[0]
    with oracle.flux_solver(oracle.RIEMANN_RUSANOV):
        other = oracle.run(op, 10)[0]
    again = oracle.run(op, 10)[0]
    assert np.array_equal(base, again) and not np.array_equal(base, other)


@pytest.mark.skipif(not os.path.exists(os.path.join(oracle.REF_DIR, "ref_dump_kokkos")) or
                    not os.path.exists(os.path.join(oracle.REF_DIR, "ref_dump")),
                    reason="needs both builds of the reference driver (oracle/Makefile and baseline/build_ref_omp.sh)")
@pytest.mark.parametrize("deck,ov", [("implode", dict(mesh__nx=96, mesh__ny=64)), ("shocked_bubble", {}),
                                     ("four_quadrant", dict(mesh__nx=80, mesh__ny=80))])
def test_shim_loop_runner_equals_real_kokkos_openmp(deck, ov):
    """The reference's sources give the same bits on the test shim's loop runner (oracle/kokkos_shim) and on the real
    Kokkos 5.1.0 / OpenMP runtime: the golden fixtures were generated with the former."""
    with tempfile.TemporaryDirectory() as td:
        ini = write_deck(os.path.join(td, "d.ini"), deck, run__nOutput=-1, **ov)
        a = oracle.ref_run(ini, nstep=40, binary=os.path.join(oracle.REF_DIR, "ref_dump"), threads=4)
        b = oracle.ref_run(ini, nstep=40, binary=os.path.join(oracle.REF_DIR, "ref_dump_kokkos"), threads=4)
    assert np.array_equal(a["U"].view(np.uint64), b["U"].view(np.uint64))
    assert np.array_equal(a["dts"], b["dts"])
