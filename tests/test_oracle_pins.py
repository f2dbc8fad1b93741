"""Pins the CPU oracle (oracle/euler2d_oracle.c) before anything trusts it:
  1. against the committed golden fixtures generated from the COMPILED REFERENCE (tests/golden/make_golden.py),
  2. against the known answers SURVEY.md Appendix B recorded from the reference's Kokkos/OpenMP build,
  3. when oracle/_ref is present (it is built from /root/reference in the build container and travels to the
     GPU box), against the reference binary run live — bit for bit, ghosts included."""
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

import oracle
from euler2d_kokkos_b200.decks import write_deck
from util import GOLDEN, assert_bitwise, both_params

from golden.make_golden import RADIAL_CASES, SMALL_CASES, serial_sum

SMALL = np.load(os.path.join(GOLDEN, "small_cases.npz"))
KAT = np.load(os.path.join(GOLDEN, "kat.npz"))
STOCK = json.load(open(os.path.join(GOLDEN, "stock_decks.json")))


@pytest.mark.parametrize("name", list(SMALL_CASES))
def test_oracle_matches_golden_small_cases(name):
    deck, ov, steps = SMALL_CASES[name]
    _, op = both_params(deck, run__nOutput=-1, **ov)
    U, dts, n, t = oracle.run(op, steps)
    nstep_ref, t_ref = SMALL[name + "__meta"]
    assert n == int(nstep_ref) and t == t_ref
    assert_bitwise(dts, SMALL[name + "__dts"], "dt sequence")
    assert_bitwise(U, SMALL[name + "__U"], name)  # every cell, ghost cells included


@pytest.mark.parametrize("deck", ["implode", "blast", "four_quadrant", "discontinuity", "shocked_bubble"])
def test_oracle_matches_golden_stock_decks(deck):
    g = STOCK[deck]
    _, op = both_params(deck, run__nOutput=-1)
    U, dts, n, t = oracle.run(op, 100)
    assert n == g["nstep"] and t == float.fromhex(g["t_hex"])
    assert hashlib.sha256(U.tobytes()).hexdigest() == g["sha256_U"]
    assert [float(x).hex() for x in dts[[0, 1, 2, 3, 51]]] == g["dts_hex"]
    for k, v in g["params_hex"].items():
        name = {"tend": "tEnd"}.get(k[:-4], k[:-4])
        assert getattr(op, name) == float.fromhex(v), k


# SURVEY.md Appendix B, recorded from the reference's real Kokkos 5.1.0 / OpenMP build
APPENDIX_B = {
    "implode": dict(dt0="0x1.2bdb427efff03p-9", dt1="0x1.fbc82470a44b9p-10", dt2="0x1.d6e6a58c9813bp-10",
                    dt50="0x1.49e459c8c89p-10", t100="0x1.0bc50350205d8p-3",
                    sums=["0x1.d67b000000009p+13", "0x1.68565d7c46669p+14", "-0x1.5b0ff583cf0fp+10",
                          "-0x1.f53a7c0581638p+9"]),
    "blast": dict(dt0="0x1.916799fc1f455p-11", dt1="0x1.7b5900b73ae9ap-11", dt2="0x1.650a53c556be1p-11",
                  dt50="0x1.18f141186d616p-10", t100="0x1.87d76642006f9p-4",
                  sums=["0x1.cc699acb3fff2p+14", "0x1.59d55543ed44cp+12"]),
    "four_quadrant": dict(dt0="0x1.c7b1ea9e1cfa4p-11", dt1="0x1.a6105b5e595dbp-11", dt2="0x1.a9ea080237478p-11",
                          dt50="0x1.5ce1865fa13b1p-11", t100="0x1.1cbd003e6113ap-4",
                          sums=["0x1.6ba96e342c886p+14", "0x1.2e8955bab84e2p+15", "0x1.ab289cd811eb5p+13",
                                "0x1.ab289cd812fb6p+13"]),
    "discontinuity": dict(dt0="0x1.3d569919d5802p-11", dt1="0x1.3d569919d5802p-11", dt2="0x1.3d569919d5802p-11",
                          dt50=None,  # the survey wrote "dt == const" loosely: by step 50 round-off moved the last bit
                          t100="0x1.efd74f385d98p-5",
                          sums=["0x1.69956acp+15", "0x1.806266b0bbfffp+16"]),
    "shocked_bubble": dict(dt0="0x1.55633da118c1cp-21", dt1="0x1.55633da118c1cp-21", dt2="0x1.5550247c7ac09p-21",
                           dt50="0x1.53c06045dec6ep-21", t100="0x1.0976b41ade20cp-14",
                           sums=["0x1.e23743d67dd2dp+15", "0x1.77a30592814c5p+34", "0x1.980683e79d5f7p+21", "0x0p+0"]),
}


@pytest.mark.parametrize("deck", list(APPENDIX_B))
def test_oracle_matches_survey_appendix_b(deck):
    b = APPENDIX_B[deck]
    _, op = both_params(deck, run__nOutput=-1)
    U, dts, n, t = oracle.run(op, 100)
    # dts[0] is the dt of main.cpp:87, dts[k+1] the dt used by step k
    assert dts[1] == float.fromhex(b["dt0"]) and dts[2] == float.fromhex(b["dt1"])
    assert dts[3] == float.fromhex(b["dt2"])
    if b["dt50"]:
        assert dts[51] == float.fromhex(b["dt50"])
    assert t == float.fromhex(b["t100"])
    for v, s in enumerate(b["sums"]):
        assert serial_sum(U[v][2:-2, 2:-2]) == float.fromhex(s), (deck, v)


def test_parameter_values_of_appendix_b():
    _, op = both_params("implode")
    assert op.gamma0 == float.fromhex("0x1.aa7efap+0") and op.cfl == float.fromhex("0x1.99999ap-1")
    assert op.smallr == op.smallc == float.fromhex("0x1.b7cdfep-34")
    assert op.smallp == float.fromhex("0x1.c587529038bb1p-68") and op.smallpp == float.fromhex("0x1.8593fef75958ap-101")
    assert op.gamma6 == float.fromhex("0x1.99a955b0f383dp-1") and op.dx == op.dy == 2.0 ** -7 and op.tEnd == 10.0
    _, op = both_params("four_quadrant")
    assert op.dx == 2.0 ** -8 and op.tEnd == float.fromhex("0x1.333334p-1")
    _, op = both_params("shocked_bubble")
    assert op.gamma0 == float.fromhex("0x1.333334p+0") and op.cfl == 0.5
    assert op.dx == float.fromhex("0x1.0624dce869da1p-10") and op.dy == float.fromhex("0x1.0624dd7baf75fp-10")
    assert op.gamma6 == float.fromhex("0x1.d55554c71c722p-1") and op.smallp == float.fromhex("0x1.3ad30db81bf88p-67")


@pytest.mark.parametrize("gname,deck", [("g1666", "implode"), ("g12", "shocked_bubble")])
@pytest.mark.parametrize("func", ["prim", "slope", "trace", "hllc", "approx", "cmpflx"])
def test_oracle_functions_match_golden_kats(gname, deck, func):
    _, op = both_params(deck)
    rec, ref = KAT[f"{gname}__{func}__in"], KAT[f"{gname}__{func}__out"]
    if func == "prim":
        q, c = oracle.compute_primitives(op, rec)
        out = np.concatenate([q, c[:, None]], axis=1)
    else:
        out = {"slope": oracle.slopes, "trace": oracle.trace, "hllc": oracle.riemann_hllc,
               "approx": oracle.riemann_approx, "cmpflx": oracle.cmpflx}[func](op, rec)
    assert_bitwise(out, ref, f"{func} {gname}")


def test_full_length_step_counts():
    """'identical step count': four_quadrant stops on tEnd after 1029 steps, discontinuity after 992."""
    for deck in ("four_quadrant", "discontinuity"):
        g = STOCK[deck]
        _, op = both_params(deck, run__nOutput=-1)
        U, dts, n, t = oracle.run(op)
        assert n == g["full_nstep"] and t == float.fromhex(g["full_t_hex"]) == op.tEnd
        assert hashlib.sha256(U.tobytes()).hexdigest() == g["full_sha256_U"]
    assert STOCK["four_quadrant"]["full_nstep"] == 1029 and STOCK["discontinuity"]["full_nstep"] == 992


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("deck,ov,steps", [("implode", dict(mesh__nx=70, mesh__ny=50), 30),
                                           ("shocked_bubble", dict(mesh__nx=120, mesh__ny=30), 30),
                                           ("four_quadrant", dict(mesh__nx=64, mesh__ny=64,
                                                                  other__implementationVersion=1), 30)])
def test_oracle_matches_compiled_reference_live(deck, ov, steps):
    with tempfile.TemporaryDirectory() as td:
        ini = write_deck(os.path.join(td, "d.ini"), deck, run__nOutput=-1, **ov)
        r = oracle.ref_run(ini, nstep=steps, threads=2)
        op = oracle.params_from_ini(ini)
    U, dts, n, t = oracle.run(op, steps)
    assert n == r["meta"]["nstep"] and t == r["meta"]["t"]
    assert_bitwise(dts, r["dts"], "dts")
    assert_bitwise(U, r["U"], deck)


# ---- Sedov post-processing (ComputeRadialProfileFunctor.h), fixtures written by the compiled reference itself
RADIAL = np.load(os.path.join(GOLDEN, "radial_profile.npz"))


@pytest.mark.parametrize("name", list(RADIAL_CASES))
def test_oracle_radial_profile_matches_reference_npy(name):
    ov, steps = RADIAL_CASES[name]
    _, op = both_params("sedov_blast_2d", run__nOutput=-1, **ov)
    U, _, n, _ = oracle.run(op, steps)
    assert n == steps
    assert_bitwise(U, RADIAL[name + "__Ufinal"], "sedov state")
    # main.cpp:178 hands hydro->U to the functor whatever the parity of nStep
    dist, sums, counts = oracle.radial_profile(op, RADIAL[name + "__U"])
    assert len(dist) == op.blast_nbins == len(RADIAL[name + "__profile"])
    assert_bitwise(dist, RADIAL[name + "__distances"], "radial distances")
    with np.errstate(invalid="ignore", divide="ignore"):
        prof = sums / counts  # :130, NaN for an empty bin
    ref = RADIAL[name + "__profile"]
    assert np.array_equal(np.isnan(prof), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert_bitwise(prof[ok], ref[ok], "density profile (serial order = the reference on one thread)")
    # every cell of the array is counted except the corner ghost cells the reference bins out of bounds
    assert counts.sum() <= op.isize * op.jsize and counts.sum() >= op.nx * op.ny


def test_oracle_radial_profile_slab_partition():
    _, op = both_params("sedov_blast_2d", mesh__nx=40, mesh__ny=33, blast__radius=0.06, run__nOutput=-1)
    U, _, _, _ = oracle.run(op, 10)
    d0, s0, c0 = oracle.radial_profile(op, U)
    # rows split into three pieces, addressed as slabs with a global row offset
    cuts = [0, 9, 20, op.jsize]
    c = np.zeros_like(c0)
    s = np.zeros_like(s0)
    for a, b in zip(cuts[:-1], cuts[1:]):
        slab = np.ascontiguousarray(U[:, a:b, :])
        _, sk, ck = oracle.radial_profile(op, slab, j_off=a)
        c += ck
        s += sk
    assert np.array_equal(c, c0)
    np.testing.assert_allclose(s, s0, rtol=1e-13)
