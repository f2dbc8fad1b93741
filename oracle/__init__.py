"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the CPU oracle (oracle/euler2d_oracle.c) and
helpers to run the compiled reference (oracle/_ref/).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product package ``euler2d_kokkos_b200`` never does.

Parity status: PINNED (see oracle/README.md): the C restatement is checked bit-for-bit against
the reference's own sources compiled in oracle/_ref/ and against tests/golden/ fixtures that were
generated from them.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"  # only present in the build container, never on the GPU box


class Params(C.Structure):
    """Mirror of ``e2do_params`` (oracle/euler2d_oracle.h)."""

    _fields_ = (
        [("nStepmax", C.c_int), ("tEnd", C.c_double), ("nOutput", C.c_int), ("enableOutput", C.c_int)]
        + [(n, C.c_int) for n in ("nx", "ny", "ghostWidth", "imin", "imax", "jmin", "jmax", "isize", "jsize")]
        + [(n, C.c_double) for n in ("xmin", "xmax", "ymin", "ymax", "dx", "dy")]
        + [(n, C.c_int) for n in ("boundary_type_xmin", "boundary_type_xmax", "boundary_type_ymin",
                                  "boundary_type_ymax", "ioVTK", "ioHDF5")]
        + [(n, C.c_double) for n in ("gamma0", "gamma6", "cfl", "slope_type", "smallr", "smallc", "smallp",
                                     "smallpp")]
        + [(n, C.c_int) for n in ("niter_riemann", "riemannSolverType", "problemType")]
        + [(n, C.c_double) for n in ("blast_radius", "blast_center_x", "blast_center_y", "blast_density_in",
                                     "blast_density_out", "blast_pressure_in", "blast_pressure_out",
                                     "blast_total_energy_inside")]
        + [("blast_nbins", C.c_int)]
        + [(n, C.c_double) for n in ("bubble_radius", "bubble_center_x", "bubble_center_y", "bubble_density",
                                     "bubble_pressure", "preshock_density", "preshock_pressure",
                                     "postshock_density", "postshock_pressure", "postshock_velocity",
                                     "shock_loc")]
        + [("implementationVersion", C.c_int)]
    )

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None

# The checker runs small grids: on a many-core GPU host an OpenMP team of 100+ threads spends its time in barriers
# (the GPU test-suite took 30x longer on a 2-GPU box than on a 1-GPU one).  Cap the team unless the caller chose.
os.environ.setdefault("OMP_NUM_THREADS", str(min(os.cpu_count() or 1, 8)))


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    if force or not os.path.exists(LIB_PATH) or (
        os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "euler2d_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        pp = C.POINTER(Params)
        L.e2do_params_from_ini.argtypes = [C.c_char_p, pp]
        L.e2do_params_from_ini.restype = C.c_int
        L.e2do_compute_primitives.argtypes = [pp, dp, dp, dp]
        L.e2do_slope_unsplit_hydro_2d.argtypes = [pp, dp, dp, dp, dp, dp, dp, dp]
        L.e2do_trace_unsplit_2d_along_dir.argtypes = [pp, dp, dp, dp, C.c_double, C.c_double, C.c_int, dp]
        L.e2do_riemann_hllc.argtypes = [pp, dp, dp, dp]
        L.e2do_riemann_approx.argtypes = [pp, dp, dp, dp, dp]
        L.e2do_cmpflx.argtypes = [pp, dp, dp]
        L.e2do_init_slab.argtypes = [pp, dp, C.c_int, C.c_int]
        L.e2do_make_boundaries_slab.argtypes = [pp, dp, C.c_int, C.c_int, C.c_int]
        L.e2do_compute_invdt_slab.argtypes = [pp, dp, C.c_int]
        L.e2do_compute_invdt_slab.restype = C.c_double
        L.e2do_convert_to_primitives_slab.argtypes = [pp, dp, dp, C.c_int]
        L.e2do_compute_and_store_fluxes_slab.argtypes = [pp, dp, dp, dp, C.c_double, C.c_double, C.c_int]
        L.e2do_update_slab.argtypes = [pp, dp, dp, dp, C.c_int]
        L.e2do_godunov_slab.argtypes = [pp, dp, dp, dp, C.c_double, C.c_int]
        L.e2do_radial_profile_slab.argtypes = [pp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp,
                                               C.POINTER(C.c_int)]
        L.e2do_run.argtypes = [pp, dp, dp, C.c_long, dp, C.c_long, dp]
        L.e2do_run.restype = C.c_int
        L.e2do_riemann_hll.argtypes = [pp, dp, dp, dp]
        L.e2do_riemann_rusanov.argtypes = [pp, dp, dp, dp]
        L.e2do_set_flux_solver.argtypes = [C.c_int]
        L.e2do_get_flux_solver.restype = C.c_int
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def params_from_ini(path: str) -> Params:
    p = Params()
    rc = lib().e2do_params_from_ini(os.fsencode(path), C.byref(p))
    if rc != 0:
        raise FileNotFoundError(path)
    return p


# ------------------------------------------------------------------ per-cell functions (vectorised over records)
def compute_primitives(p: Params, u: np.ndarray):
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 4)
    q = np.empty_like(u)
    c = np.empty(len(u))
    cc = C.c_double()
    for k in range(len(u)):
        lib().e2do_compute_primitives(C.byref(p), _dp(u[k]), C.byref(cc), _dp(q[k]))
        c[k] = cc.value
    return q, c


def slopes(p: Params, rec: np.ndarray):
    rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 20)
    out = np.zeros((len(rec), 8))
    for k in range(len(rec)):
        r = rec[k]
        lib().e2do_slope_unsplit_hydro_2d(C.byref(p), _dp(r[0:4]), _dp(r[4:8]), _dp(r[8:12]), _dp(r[12:16]),
                                           _dp(r[16:20]), _dp(out[k, 0:4]), _dp(out[k, 4:8]))
    return out


def trace(p: Params, rec: np.ndarray):
    rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 14)
    out = np.zeros((len(rec), 16))
    for k in range(len(rec)):
        r = rec[k]
        for face in range(4):
            lib().e2do_trace_unsplit_2d_along_dir(C.byref(p), _dp(r[0:4]), _dp(r[4:8]), _dp(r[8:12]),
                                                   r[12], r[13], face, _dp(out[k, 4 * face:4 * face + 4]))
    return out


def riemann_hllc(p: Params, rec: np.ndarray):
    rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 8)
    out = np.zeros((len(rec), 4))
    for k in range(len(rec)):
        lib().e2do_riemann_hllc(C.byref(p), _dp(rec[k, 0:4]), _dp(rec[k, 4:8]), _dp(out[k]))
    return out


RIEMANN_APPROX, RIEMANN_HLL, RIEMANN_HLLC, RIEMANN_RUSANOV = 0, 1, 2, 3


class flux_solver:
    """``with oracle.flux_solver(oracle.RIEMANN_APPROX): oracle.run(p)`` — the array operators and ``run`` use that
    solver instead of the reference's hard-wired riemann_hllc (the product's opt-in ``honourRiemannSolver``)."""

    def __init__(self, solver: int):
        self.solver = solver

    def __enter__(self):
        self.prev = lib().e2do_get_flux_solver()
        lib().e2do_set_flux_solver(self.solver)
        return self

    def __exit__(self, *exc):
        lib().e2do_set_flux_solver(self.prev)


def _two_state(fn, p: Params, rec: np.ndarray):
    rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 8)
    out = np.zeros((len(rec), 4))
    for k in range(len(rec)):
        fn(C.byref(p), _dp(rec[k, 0:4]), _dp(rec[k, 4:8]), _dp(out[k]))
    return out


def riemann_hll(p: Params, rec: np.ndarray):
    """extension solver (not in the reference): flux of the HLL two-wave solver"""
    return _two_state(lib().e2do_riemann_hll, p, rec)


def riemann_rusanov(p: Params, rec: np.ndarray):
    """extension solver (not in the reference): Rusanov / local Lax-Friedrichs flux"""
    return _two_state(lib().e2do_riemann_rusanov, p, rec)


def riemann_approx(p: Params, rec: np.ndarray):
    rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 8)
    out = np.zeros((len(rec), 8))
    for k in range(len(rec)):
        lib().e2do_riemann_approx(C.byref(p), _dp(rec[k, 0:4]), _dp(rec[k, 4:8]), _dp(out[k, 0:4]),
                                  _dp(out[k, 4:8]))
    return out


def cmpflx(p: Params, rec: np.ndarray):
    rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 4)
    out = np.zeros((len(rec), 4))
    for k in range(len(rec)):
        lib().e2do_cmpflx(C.byref(p), _dp(rec[k]), _dp(out[k]))
    return out


# ------------------------------------------------------------------ array-level operators ([var][j][i] arrays)
def alloc(p: Params, jsize: int | None = None) -> np.ndarray:
    return np.zeros((4, jsize if jsize is not None else p.jsize, p.isize))


def init_slab(p: Params, jsize: int | None = None, j_off: int = 0) -> np.ndarray:
    U = alloc(p, jsize)
    lib().e2do_init_slab(C.byref(p), _dp(U), U.shape[1], j_off)
    return U


def make_boundaries(p: Params, U: np.ndarray, do_ymin: bool = True, do_ymax: bool = True) -> None:
    lib().e2do_make_boundaries_slab(C.byref(p), _dp(U), U.shape[1], int(do_ymin), int(do_ymax))


def compute_invdt(p: Params, U: np.ndarray) -> float:
    return lib().e2do_compute_invdt_slab(C.byref(p), _dp(U), U.shape[1])


def convert_to_primitives(p: Params, U: np.ndarray) -> np.ndarray:
    Q = np.empty_like(U)
    lib().e2do_convert_to_primitives_slab(C.byref(p), _dp(U), _dp(Q), U.shape[1])
    return Q


def compute_and_store_fluxes(p: Params, Q: np.ndarray, dtdx: float, dtdy: float):
    Fx = np.zeros_like(Q)
    Fy = np.zeros_like(Q)
    lib().e2do_compute_and_store_fluxes_slab(C.byref(p), _dp(Q), _dp(Fx), _dp(Fy), dtdx, dtdy, Q.shape[1])
    return Fx, Fy


def update(p: Params, U: np.ndarray, Fx: np.ndarray, Fy: np.ndarray) -> None:
    lib().e2do_update_slab(C.byref(p), _dp(U), _dp(Fx), _dp(Fy), U.shape[1])


def godunov(p: Params, Uin: np.ndarray, dt: float) -> np.ndarray:
    """out-of-place step on an array whose ghost cells are already filled."""
    Uout = np.empty_like(Uin)
    work = np.zeros((3,) + Uin.shape)
    lib().e2do_godunov_slab(C.byref(p), _dp(Uin), _dp(Uout), _dp(work), dt, Uin.shape[1])
    return Uout


def radial_profile(p: Params, U: np.ndarray, nbins: int | None = None, j_off: int = 0, j_lo: int = 0,
                   j_hi: int | None = None):
    """ComputeRadialProfileFunctor (src/ComputeRadialProfileFunctor.h). Returns (distances, sums, counts); the
    profile the reference saves is sums / counts."""
    nbins = p.blast_nbins if nbins is None else nbins
    j_hi = U.shape[1] if j_hi is None else j_hi
    dist, sums, counts = np.zeros(nbins), np.zeros(nbins), np.zeros(nbins, dtype=np.int32)
    lib().e2do_radial_profile_slab(C.byref(p), _dp(U), U.shape[1], j_off, j_lo, j_hi, nbins, _dp(dist), _dp(sums),
                                   counts.ctypes.data_as(C.POINTER(C.c_int)))
    return dist, sums, counts


def run(p: Params, max_steps: int = -1):
    """Whole-domain driver (main.cpp:86-143). Returns (U_final, dt_seq, nstep, t)."""
    U = alloc(p)
    U2 = alloc(p)
    cap = (p.nStepmax if max_steps < 0 else max_steps) + 2
    dts = np.zeros(cap)
    t = C.c_double()
    n = lib().e2do_run(C.byref(p), _dp(U), _dp(U2), max_steps, _dp(dts), cap, C.byref(t))
    return (U if n % 2 == 0 else U2), dts[: n + 1].copy(), n, t.value


# ------------------------------------------------------------------ compiled reference (oracle/_ref)
def ref_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_dump"))


def ref_binary(prefer_kokkos: bool = True) -> str:
    k = os.path.join(REF_DIR, "ref_dump_kokkos")
    if prefer_kokkos and os.path.exists(k):
        return k
    return os.path.join(REF_DIR, "ref_dump")


def _omp_env(threads: int | None):
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
    env.setdefault("OMP_PROC_BIND", "spread")
    env.setdefault("OMP_PLACES", "threads")
    return env


def ref_run(ini: str, nstep: int | None = None, dump: bool = True, threads: int | None = None,
            states_every: int = 0, binary: str | None = None, radial: bool = False):
    """Run the compiled reference on an .ini. Returns dict(meta, U, dts[, states][, radial_*])."""
    exe = binary or ref_binary()
    with tempfile.TemporaryDirectory() as td:
        cmd = [exe, ini]
        prefix = os.path.join(td, "o")
        if dump:
            cmd += ["--out", prefix]
        if nstep is not None:
            cmd += ["--nstep", str(nstep)]
        if states_every:
            cmd += ["--states", str(states_every)]
        if radial:
            cmd += ["--radial"]  # the reference writes its two .npy files into the cwd
        out = subprocess.run(cmd, check=True, capture_output=True, text=True, env=_omp_env(threads), cwd=td).stdout
        meta = json.loads(out.strip().splitlines()[-1])
        res = {"meta": meta}
        if radial:
            res["radial_distances"] = np.load(os.path.join(td, "sedov_blast_radial_distances.npy"))
            res["radial_profile"] = np.load(os.path.join(td, "sedov_blast_density_profile.npy"))
            if dump:
                res["radial_U"] = np.fromfile(prefix + ".Uradial.bin", dtype=np.float64).reshape(
                    (4, meta["jsize"], meta["isize"]))
        if dump:
            shape = (4, meta["jsize"], meta["isize"])
            res["U"] = np.fromfile(prefix + ".U.bin", dtype=np.float64).reshape(shape)
            res["dts"] = np.fromfile(prefix + ".dt.bin", dtype=np.float64)
            if states_every:
                res["states"] = {}
                for f in os.listdir(td):
                    if f.startswith("o.s") and f.endswith(".bin"):
                        k = int(f[3:-4])
                        res["states"][k] = np.fromfile(os.path.join(td, f), dtype=np.float64).reshape(shape)
        return res


def ref_kat(ini: str, func: str, records: np.ndarray) -> np.ndarray:
    nout = {"prim": 5, "slope": 8, "trace": 16, "hllc": 4, "approx": 8, "cmpflx": 4}[func]
    records = np.ascontiguousarray(records, dtype=np.float64)
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        records.tofile(fi)
        subprocess.run([os.path.join(REF_DIR, "ref_kat"), ini, func, fi, fo], check=True, capture_output=True)
        return np.fromfile(fo, dtype=np.float64).reshape(-1, nout)


def write_ini(path: str, base_ini: str | None = None, **overrides) -> str:
    """Write an .ini: ``base_ini`` text followed by override sections (later assignments win in inih).
    overrides: section__key=value, e.g. mesh__nx=64."""
    text = open(base_ini).read() if base_ini else ""
    by_sec: dict[str, list[str]] = {}
    for k, v in overrides.items():
        sec, key = k.split("__", 1)
        by_sec.setdefault(sec, []).append(f"{key}={v}")
    for sec, lines in by_sec.items():
        text += f"\n[{sec}]\n" + "\n".join(lines) + "\n"
    with open(path, "w") as f:
        f.write(text)
    return path
