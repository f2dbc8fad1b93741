"""The product's kernel bodies on the host.  tests/host_emulation/march_emul.cpp compiles
euler2d_kokkos_b200/csrc/e2d_march.cuh + e2d_math.cuh with g++ (-ffp-contract=off) and runs the fused
marching kernel one emulated thread at a time; this checks the kernel's pipelining / indexing logic and its
formulas against the oracle without a GPU (the GPU run of the same code is tests/test_gpu_kernels.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from util import GOLDEN, INNER, assert_bitwise, both_params, random_conservative_field

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emulation", "march_emul.cpp")
SO = os.path.join(HERE, "host_emulation", "libmarch_emul.so")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC] + [os.path.join(HERE, "..", "euler2d_kokkos_b200", "csrc", f) for f in ("e2d_march.cuh", "e2d_math.cuh", "e2d_lean.cuh", "e2d_fast.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-Wno-unknown-pragmas", "-o", SO, SRC])
    return C.CDLL(SO)


def fused(emul, hp, Uin, dt, seg, bx, solver=2):
    Uout = np.full_like(Uin, np.nan)
    inv = C.c_double()
    rc = emul.emul_fused_step(C.byref(hp.raw), Uin.ctypes.data_as(C.c_void_p), Uout.ctypes.data_as(C.c_void_p),
                              C.c_int(Uin.shape[1]), C.c_double(dt), C.c_int(seg), C.c_int(bx), C.c_int(solver),
                              C.byref(inv))
    assert rc == 0
    return Uout, inv.value


@pytest.mark.parametrize("deck,nx,ny", [("implode", 70, 41), ("blast", 33, 64), ("shocked_bubble", 130, 9),
                                        ("four_quadrant", 28, 28), ("implode", 2, 2)])
@pytest.mark.parametrize("bx,seg", [(32, 7), (128, 16), (16, 1), (32, 1000)])
def test_marching_kernel_logic_is_bit_exact(emul, deck, nx, ny, bx, seg):
    hp, op = both_params(deck, mesh__nx=nx, mesh__ny=ny)
    rng = np.random.default_rng(nx * 7 + ny)
    U = random_conservative_field(rng, op)
    oracle.make_boundaries(op, U)
    dt = op.cfl / oracle.compute_invdt(op, U)
    ref = oracle.godunov(op, U, dt)
    out, inv = fused(emul, hp, U, dt, seg, bx)
    assert_bitwise(out[INNER], ref[INNER], f"{deck} {nx}x{ny} bx={bx} seg={seg}")
    assert inv == oracle.compute_invdt(op, ref)
    mask = np.ones(out.shape, bool)
    mask[INNER] = False
    assert np.isnan(out[mask]).all(), "the kernel wrote outside the interior"


@pytest.mark.parametrize("deck,nx,ny", [("implode", 70, 41), ("blast", 33, 64), ("shocked_bubble", 130, 9),
                                        ("four_quadrant", 28, 28), ("implode", 2, 2)])
@pytest.mark.parametrize("bx,seg", [(32, 7), (128, 16), (16, 1), (32, 2), (32, 1000)])
def test_peeled_march_is_bit_exact(emul, deck, nx, ny, bx, seg):
    """The peeled march of short segments (k_fused_step<.., PEEL>: the first phase B of a segment only advances the ring,
    the last one solves the south face only) gives the same bits, including the fused CFL reduction, and writes
    nothing outside the interior — with poisoned shared memory, so nothing it skips is read later."""
    hp, op = both_params(deck, mesh__nx=nx, mesh__ny=ny)
    rng = np.random.default_rng(nx * 11 + ny)
    U = random_conservative_field(rng, op)
    oracle.make_boundaries(op, U)
    dt = op.cfl / oracle.compute_invdt(op, U)
    ref = oracle.godunov(op, U, dt)
    emul.emul_set_peel(1)
    try:
        out, inv = fused(emul, hp, U, dt, seg, bx)
    finally:
        emul.emul_set_peel(0)
    assert_bitwise(out[INNER], ref[INNER], f"peeled {deck} {nx}x{ny} bx={bx} seg={seg}")
    assert inv == oracle.compute_invdt(op, ref)
    mask = np.ones(out.shape, bool)
    mask[INNER] = False
    assert np.isnan(out[mask]).all(), "the kernel wrote outside the interior"


@pytest.mark.parametrize("deck,nx,ny", [("implode", 70, 41), ("blast", 33, 64), ("shocked_bubble", 130, 9),
                                        ("four_quadrant", 28, 28)])
@pytest.mark.parametrize("bx,seg", [(32, 7), (128, 16)])
def test_fast_arithmetic_marching_logic_and_formulas(emul, deck, nx, ny, bx, seg):
    """`arithmetic=fast` (e2d_fast.cuh) through the same state machine: the formulas are the reference's up to
    rounding (1e-13 after one step of a random field; the GPU tests hold the 1e-12 bar over hundreds of steps)."""
    from euler2d_kokkos_b200.parity import state_deviation

    hp, op = both_params(deck, mesh__nx=nx, mesh__ny=ny, other__arithmetic="fast")
    assert hp.arithmetic == 1
    rng = np.random.default_rng(nx * 7 + ny)
    U = random_conservative_field(rng, op)
    oracle.make_boundaries(op, U)
    dt = op.cfl / oracle.compute_invdt(op, U)
    ref = oracle.godunov(op, U, dt)
    out, inv = fused(emul, hp, U, dt, seg, bx)
    for name, l1, linf in state_deviation(out[INNER], ref[INNER]):
        assert l1 <= 1e-13 and linf <= 1e-13, (name, l1, linf)
    assert abs(inv - oracle.compute_invdt(op, ref)) <= 1e-13 * inv
    mask = np.ones(out.shape, bool)
    mask[INNER] = False
    assert np.isnan(out[mask]).all(), "the kernel wrote outside the interior"


def test_marching_kernel_on_a_developed_flow(emul):
    hp, op = both_params("implode", mesh__nx=96, mesh__ny=48)
    U, _, n, _ = oracle.run(op, 60)
    oracle.make_boundaries(op, U)
    dt = op.cfl / oracle.compute_invdt(op, U)
    ref = oracle.godunov(op, U, dt)
    out, _ = fused(emul, hp, U, dt, 13, 32)
    assert_bitwise(out[INNER], ref[INNER], "developed implode")


@pytest.mark.parametrize("func", ["prim", "slope", "trace", "hllc", "approx", "cmpflx"])
def test_product_formulas_match_reference_kats(emul, func):
    """e2d_math.cuh (compiled for the host) against outputs of the reference's own HydroBaseFunctor methods."""
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))
    nout = {"prim": 5, "slope": 8, "trace": 16, "hllc": 4, "approx": 8, "cmpflx": 4}[func]
    for gname, deck in (("g1666", "implode"), ("g12", "shocked_bubble")):
        hp, _ = both_params(deck)
        rec = np.ascontiguousarray(kat[f"{gname}__{func}__in"])
        out = np.zeros((len(rec), nout))
        dp = C.POINTER(C.c_double)
        assert emul.emul_eval(C.byref(hp.raw), func.encode(), rec.ctypes.data_as(dp), out.ctypes.data_as(dp),
                              C.c_long(len(rec))) == 0
        assert_bitwise(out, kat[f"{gname}__{func}__out"], f"{func} {gname}")
