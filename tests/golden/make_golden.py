"""Generates the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref/ref_dump, ref_kat =
the reference's own sources built by oracle/Makefile).  Run in the build container, where /root/reference
exists:   python tests/golden/make_golden.py
The fixtures pin the C oracle (tests/test_oracle_pins.py) and, through it, the CUDA path."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from euler2d_kokkos_b200.decks import DECKS, write_deck  # noqa: E402

SMALL_CASES = {
    # name: (deck, overrides, steps)
    "implode_48x32": ("implode", dict(mesh__nx=48, mesh__ny=32), 40),
    "blast_32x48": ("blast", dict(mesh__nx=32, mesh__ny=48), 40),
    "four_quadrant_40x40": ("four_quadrant", dict(mesh__nx=40, mesh__ny=40), 40),
    "discontinuity_36x36": ("discontinuity", dict(mesh__nx=36, mesh__ny=36), 40),
    "shocked_bubble_89x18": ("shocked_bubble", dict(mesh__nx=89, mesh__ny=18), 40),
    "periodic_all_40x24": ("four_quadrant", dict(mesh__nx=40, mesh__ny=24, mesh__boundary_type_xmin=3,
                                                 mesh__boundary_type_xmax=3, mesh__boundary_type_ymin=3,
                                                 mesh__boundary_type_ymax=3), 40),
    "mixed_bc_40x24": ("four_quadrant", dict(mesh__nx=40, mesh__ny=24, mesh__boundary_type_xmin=3,
                                             mesh__boundary_type_xmax=3, mesh__boundary_type_ymin=2,
                                             mesh__boundary_type_ymax=1), 40),
    "slope_type1_40x24": ("implode", dict(mesh__nx=40, mesh__ny=24, hydro__slope_type=1), 40),
    "slope_type0_40x24": ("implode", dict(mesh__nx=40, mesh__ny=24, hydro__slope_type=0), 40),
    "tend_hit_24x24": ("four_quadrant", dict(mesh__nx=24, mesh__ny=24, run__tEnd=0.05), 1200),
    "sedov_64x64": ("sedov_blast_2d", dict(mesh__nx=64, mesh__ny=64, blast__radius=0.05), 30),
}


RADIAL_CASES = {
    # Sedov post-processing (ComputeRadialProfileFunctor): even and odd step counts — main.cpp:178 always passes
    # hydro->U, so after an odd number of steps the profile is taken from the older array
    "sedov_80x64_30": (dict(mesh__nx=80, mesh__ny=64, blast__radius=0.04), 30),
    "sedov_80x64_41": (dict(mesh__nx=80, mesh__ny=64, blast__radius=0.04, blast__nbins=37), 41),
}


def make_radial():
    binary = oracle.ref_binary(prefer_kokkos=False)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (ov, steps) in RADIAL_CASES.items():
            ini = write_deck(os.path.join(td, name + ".ini"), "sedov_blast_2d", run__nOutput=-1, **ov)
            r = oracle.ref_run(ini, nstep=steps, binary=binary, radial=True, threads=1)  # 1 thread: serial sums
            out[name + "__U"] = r["radial_U"]
            out[name + "__Ufinal"] = r["U"]
            out[name + "__distances"] = r["radial_distances"]
            out[name + "__profile"] = r["radial_profile"]
            print(name, r["meta"]["nstep"], r["radial_profile"][:4])
    np.savez_compressed(os.path.join(HERE, "radial_profile.npz"), **out)


def serial_sum(a):
    return float(np.cumsum(a.ravel())[-1])


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    if "--radial-only" in sys.argv:
        return make_radial()
    make_radial()
    binary = oracle.ref_binary(prefer_kokkos=False)
    small = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (deck, ov, steps) in SMALL_CASES.items():
            ini = write_deck(os.path.join(td, name + ".ini"), deck, run__nOutput=-1, **ov)
            r = oracle.ref_run(ini, nstep=steps, binary=binary)
            small[name + "__U"] = r["U"]
            small[name + "__dts"] = r["dts"]
            small[name + "__meta"] = np.array([r["meta"]["nstep"], r["meta"]["t"]])
            print(name, r["meta"]["nstep"], r["meta"]["t_hex"])
        np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **small)

        stock = {}
        for deck in ("implode", "blast", "four_quadrant", "discontinuity", "shocked_bubble"):
            ini = write_deck(os.path.join(td, deck + ".ini"), deck, run__nOutput=-1)
            r = oracle.ref_run(ini, nstep=100, binary=binary)
            U = r["U"]
            stock[deck] = {
                "nstep": r["meta"]["nstep"], "t_hex": r["meta"]["t_hex"],
                "sha256_U": hashlib.sha256(U.tobytes()).hexdigest(),
                "dts_hex": [float(x).hex() for x in r["dts"][[0, 1, 2, 3, 51]]],
                "sums_hex": [float(serial_sum(U[v][2:-2, 2:-2])).hex() for v in range(4)],
                "params_hex": {k: r["meta"][k] for k in r["meta"] if k.endswith("_hex") and k != "t_hex"},
            }
            print(deck, stock[deck]["t_hex"], stock[deck]["sums_hex"][:2])
        # full-length step counts ("identical step count")
        for deck in ("four_quadrant", "discontinuity"):
            ini = write_deck(os.path.join(td, deck + "_full.ini"), deck, run__nOutput=-1)
            r = oracle.ref_run(ini, binary=binary)
            stock[deck]["full_nstep"] = r["meta"]["nstep"]
            stock[deck]["full_t_hex"] = r["meta"]["t_hex"]
            stock[deck]["full_sha256_U"] = hashlib.sha256(r["U"].tobytes()).hexdigest()
            print(deck, "full", r["meta"]["nstep"], r["meta"]["t_hex"])
        json.dump(stock, open(os.path.join(HERE, "stock_decks.json"), "w"), indent=1)

        # function-level known answers from the reference's HydroBaseFunctor methods
        rng = np.random.default_rng(20261017)
        n = 160

        def states(k):
            q = np.empty((k, 4))
            q[:, 0] = rng.uniform(0.1, 10, k)
            q[:, 1] = rng.uniform(0.1, 10, k)
            q[:, 2] = rng.uniform(-2, 2, k)
            q[:, 3] = rng.uniform(-2, 2, k)
            return q

        kat = {}
        for gname, deck in (("g1666", "implode"), ("g12", "shocked_bubble")):
            ini = write_deck(os.path.join(td, gname + ".ini"), deck)
            q = states(n)
            g = 1.666 if gname == "g1666" else 1.2
            recs = {
                "prim": np.stack([q[:, 0], q[:, 1] / (g - 1) + 0.5 * q[:, 0] * (q[:, 2] ** 2 + q[:, 3] ** 2),
                                  q[:, 0] * q[:, 2], q[:, 0] * q[:, 3]], axis=1),
                "slope": np.concatenate([states(n) for _ in range(5)], axis=1),
                "trace": np.concatenate([states(n), rng.normal(0, 0.3, (n, 8)), rng.uniform(0.05, 0.5, (n, 2))], axis=1),
                "hllc": np.concatenate([states(n), states(n)], axis=1),
                "approx": np.concatenate([states(n), states(n)], axis=1),
                "cmpflx": states(n),
            }
            recs["prim"][::50, 0] = 1e-12
            recs["hllc"][::7, 2] += 8.0
            recs["hllc"][::7, 6] += 8.0
            recs["hllc"][1::7, 2] -= 8.0
            recs["hllc"][1::7, 6] -= 8.0
            recs["approx"][::5, 4:8] = recs["approx"][::5, 0:4]
            for func, rec in recs.items():
                kat[f"{gname}__{func}__in"] = rec
                kat[f"{gname}__{func}__out"] = oracle.ref_kat(ini, func, rec)
        np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
