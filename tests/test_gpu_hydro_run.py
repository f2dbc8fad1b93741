"""GPU parity tests at the HydroRun level: the reference's driver loop (src/main.cpp:86-143) through the C ABI
against the oracle — identical step count, identical dt sequence, identical bits in every interior cell."""
import ctypes as C
import os

import numpy as np
import pytest

import euler2d_kokkos_b200 as e2d
import oracle
from euler2d_kokkos_b200 import HydroRun
from util import INNER, assert_bitwise, both_params, rel_errors

pytestmark = pytest.mark.gpu

SMALL = {"implode": (96, 48), "blast": (64, 96), "four_quadrant": (80, 80), "discontinuity": (72, 72),
         "shocked_bubble": (178, 36)}


def host_loop(hydro, params, max_steps):
    """main.cpp:86-143 driven from the host, one compute_dt + godunov_unsplit per step."""
    t, n, dts = 0.0, 0, []
    dts.append(hydro.compute_dt(0))
    hydro.make_boundaries(HydroRun.U)
    hydro.make_boundaries(HydroRun.U2)
    while t < params.tEnd and n < max_steps:
        dt = hydro.compute_dt(n % 2)
        if t + dt > params.tEnd:
            dt = params.tEnd - t
        hydro.godunov_unsplit(n, dt)
        n += 1
        t += dt
        dts.append(dt)
    return n, t, np.array(dts)


@pytest.mark.parametrize("unfused", ["no", "yes"])
@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("deck", list(SMALL))
def test_host_driven_loop_bit_exact(deck, impl, unfused):
    """`unfused=yes`: implementations 0 / 1 as the reference's literal kernel sequence; `no` (default): the fused
    step + ghost-frame copy — the same array, ghost cells included, either way."""
    nx, ny = SMALL[deck]
    hp, op = both_params(deck, mesh__nx=nx, mesh__ny=ny, other__implementationVersion=impl,
                         other__unfusedKernels=unfused)
    assert hp.unfusedKernels == (1 if unfused == "yes" else 0)
    steps = 40
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, steps)
    with HydroRun(hp) as hydro:
        n, t, dts = host_loop(hydro, hp, steps)
        U = hydro.download(HydroRun.U if n % 2 == 0 else HydroRun.U2)
    assert n == n_ref and t == t_ref
    assert_bitwise(dts, dts_ref, "dt sequence")
    assert_bitwise(U[INNER], U_ref[INNER], f"{deck} impl {impl}")
    if impl != 2:  # implementations 0/1 also reproduce the ghost cells (out = deep_copy(in), HydroRun.h:302)
        assert_bitwise(U, U_ref, f"{deck} impl {impl} incl. ghosts")


@pytest.mark.parametrize("deck", ["implode", "blast", "four_quadrant", "discontinuity", "shocked_bubble"])
def test_device_resident_run_on_stock_decks(deck):
    """The decks as shipped by the reference, 100 steps (SURVEY.md Appendix B pins these runs)."""
    hp, op = both_params(deck, run__nOutput=-1)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, 100)
    with HydroRun(hp) as hydro:
        hydro.make_boundaries(HydroRun.U)
        hydro.make_boundaries(HydroRun.U2)
        st = hydro.run(100)
        U = hydro.download(HydroRun.U if st.nStep % 2 == 0 else HydroRun.U2)
        dts = hydro.dt_history()
    assert st.nStep == n_ref == 100
    assert st.t == t_ref
    assert_bitwise(dts, dts_ref[1:], "dt history")
    assert_bitwise(U[INNER], U_ref[INNER], deck)
    for l1, linf in rel_errors(U[INNER], U_ref[INNER]):  # north_star's stated tolerance, trivially met
        assert l1 <= 1e-12 and linf <= 1e-12


def test_run_until_tend_identical_step_count():
    """four_quadrant stops on tEnd after 1029 steps at 256x256; here a smaller grid,
    same mechanism: the clamp dt = tEnd - t (main.cpp:131-134) happens on the device."""
    hp, op = both_params("four_quadrant", mesh__nx=64, mesh__ny=64, run__nOutput=-1)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op)
    assert t_ref == op.tEnd and n_ref < op.nStepmax
    with HydroRun(hp) as hydro:
        st = hydro.run()
        U = hydro.download(HydroRun.U if st.nStep % 2 == 0 else HydroRun.U2)
        dts = hydro.dt_history()
    assert st.nStep == n_ref and st.t == t_ref == hp.tEnd
    assert_bitwise(dts, dts_ref[1:], "dt history")
    assert_bitwise(U[INNER], U_ref[INNER], "four_quadrant to tEnd")


def test_run_resumes_and_mixes_with_host_api():
    hp, op = both_params("implode", mesh__nx=64, mesh__ny=40, run__nOutput=-1)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, 31)
    with HydroRun(hp) as hydro:
        st = hydro.run(10)
        assert st.nStep == 10
        st = hydro.run(31)  # continue to a total of 31 (odd) steps
        assert st.nStep == 31 and st.t == t_ref
        U = hydro.download(HydroRun.U2)
    assert_bitwise(U[INNER], U_ref[INNER], "resumed run")


@pytest.mark.parametrize("bcs", [(3, 3, 3, 3), (3, 3, 2, 1), (2, 1, 3, 3)])
@pytest.mark.parametrize("slope_type", [0, 1, 2])
def test_boundary_and_slope_variants(bcs, slope_type):
    hp, op = both_params("four_quadrant", mesh__nx=50, mesh__ny=70, hydro__slope_type=slope_type,
                         mesh__boundary_type_xmin=bcs[0], mesh__boundary_type_xmax=bcs[1],
                         mesh__boundary_type_ymin=bcs[2], mesh__boundary_type_ymax=bcs[3], run__nOutput=-1)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, 60)
    with HydroRun(hp) as hydro:
        st = hydro.run(60)
        U = hydro.download(HydroRun.U if st.nStep % 2 == 0 else HydroRun.U2)
    assert st.nStep == n_ref and st.t == t_ref
    assert_bitwise(U[INNER], U_ref[INNER], f"bc {bcs} slope {slope_type}")


def test_step_host_round_trip():
    hp, op = both_params("blast", mesh__nx=90, mesh__ny=60)
    U0 = oracle.init_slab(op)
    ref_in = U0.copy()
    oracle.make_boundaries(op, ref_in)
    dt_ref = op.cfl / oracle.compute_invdt(op, ref_in)
    ref = oracle.godunov(op, ref_in, dt_ref)
    out = np.empty_like(U0)
    with HydroRun(hp) as hydro:
        dt = hydro.step_host(U0, out)
    assert dt == dt_ref
    assert_bitwise(out[INNER], ref[INNER], "step_host")


@pytest.mark.parametrize("chunk_rows", [0, 16, 23, 1000])
@pytest.mark.parametrize("bcs", [(1, 1, 1, 1), (2, 2, 2, 2), (3, 3, 3, 3), (3, 3, 2, 1), (1, 2, 3, 3)])
def test_step_host_streamed_march(bcs, chunk_rows):
    """A time march whose state lives in host memory, through the streamed step (chunked H2D / step / D2H overlap):
    every chunking gives the oracle's dt sequence and bits; dt is threaded from call to call (dt_next), the first
    one is computed from the uploaded state."""
    hp, op = both_params("four_quadrant", mesh__nx=70, mesh__ny=115, mesh__boundary_type_xmin=bcs[0],
                         mesh__boundary_type_xmax=bcs[1], mesh__boundary_type_ymin=bcs[2],
                         mesh__boundary_type_ymax=bcs[3], run__nOutput=-1)
    nsteps = 12
    U_ref, dts_ref, n_ref, _ = oracle.run(op, nsteps)
    a = oracle.init_slab(op)
    a[:, :2, :] = np.nan  # ghost cells of the host input are never read: the device fills them
    a[:, -2:, :] = np.nan
    a[:, :, :2] = np.nan
    a[:, :, -2:] = np.nan
    b = np.full_like(a, np.nan)
    dts, dt = [], 0.0
    with HydroRun(hp) as hydro:
        for _ in range(nsteps):
            used, dt = hydro.step_host_streamed(a, b, dt, chunk_rows)
            dts.append(used)
            a, b = b, a
    assert n_ref == nsteps
    assert np.array_equal(np.array(dts), dts_ref[1:]), "dt sequence differs from the oracle"
    assert_bitwise(a[INNER], U_ref[INNER], f"streamed march bc {bcs} chunk {chunk_rows}")
    # ghost cells of the result = make_boundaries of the new state
    filled = a.copy()
    oracle.make_boundaries(op, filled)
    assert_bitwise(a, filled, "ghost cells of the streamed result")


def test_step_host_streamed_equals_step_host():
    hp, op = both_params("blast", mesh__nx=200, mesh__ny=333)
    U0 = oracle.init_slab(op)
    out1, out2 = np.empty_like(U0), np.empty_like(U0)
    with HydroRun(hp) as hydro:
        dt1 = hydro.step_host(U0, out1)
        used, nxt = hydro.step_host_streamed(U0, out2, 0.0, 32)
        assert used == dt1
        assert nxt == hydro.step_host(out1, np.empty_like(U0))  # dt_next is compute_dt of the new state
        used2, _ = hydro.step_host_streamed(U0, out2, dt1, 50)   # dt given by the caller
        assert used2 == dt1
    assert_bitwise(out2, out1, "streamed vs plain host step")


def test_upload_download_layouts():
    hp, op = both_params("implode", mesh__nx=20, mesh__ny=12)
    rng = np.random.default_rng(3)
    A = rng.normal(size=(4, op.jsize, op.isize))
    with HydroRun(hp) as hydro:
        hydro.upload(HydroRun.U, A)
        assert_bitwise(hydro.download(HydroRun.U), A)
        K = hydro.download(HydroRun.U, e2d.LAYOUT_KOKKOS_OMP)  # (i, j, var) like the reference's OpenMP views
        assert_bitwise(K, np.ascontiguousarray(A.transpose(2, 1, 0)))
        hydro.upload(HydroRun.U2, K, e2d.LAYOUT_KOKKOS_OMP)
        assert_bitwise(hydro.download(HydroRun.U2), A)


def test_save_vtk_matches_reference_format(tmp_path):
    """Layout of the file and the 6-significant-digit values (the stream's default precision).  The byte-for-byte
    comparison with files written by the reference program itself is in test_gpu_refmain.py."""
    hp, op = both_params("implode", mesh__nx=16, mesh__ny=8, output__outputDir=str(tmp_path),
                         output__outputPrefix="vt")
    with HydroRun(hp) as hydro:
        hydro.saveData(HydroRun.U, 30, "U")
        U = hydro.download(HydroRun.U)
    text = open(os.path.join(tmp_path, "vt_0000030.vti")).read().splitlines()
    assert text[0] == '<?xml version="1.0"?>'
    assert text[1] == '<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian">'
    assert text[2] == '  <ImageData WholeExtent="0 16 0 8 0 0" Origin="-1 0 0" Spacing="0.125 0.125 0">'
    assert text[3] == '  <Piece Extent="0 16 0 8 0 0 ">'
    names = [l for l in text if "DataArray type" in l]
    assert [n.split('Name="')[1].split('"')[0] for n in names] == ["rho", "E", "mx", "my"]
    rho = np.array(text[text.index(names[0]) + 1].split(), dtype=float)
    np.testing.assert_allclose(rho, U[0][2:-2, 2:-2].ravel(), rtol=1e-5)  # 6 significant digits, like the reference


@pytest.mark.parametrize("unfused", ["no", "yes"])
def test_timers_accumulate(unfused):
    hp, _ = both_params("implode", mesh__nx=64, mesh__ny=64, other__unfusedKernels=unfused)
    with HydroRun(hp) as hydro:
        hydro.enable_timers(True)
        for n in range(4):
            hydro.godunov_unsplit(n, 1e-4)
        tm = hydro.timers()
    assert tm["godunov"] > 0 and tm["fluxes"] > 0 and tm["boundaries"] > 0 and tm["godunov"] >= tm["fluxes"]


# the opt-in `riemann=` switch (approx / hll / rusanov) is tested in test_gpu_riemann_solvers.py: whole runs bit for bit
# against the oracle with the same solver, plus the property tests of the two solvers the reference does not have


def test_errors_are_status_codes():
    L = e2d.lib()
    assert L.e2d_compute_dt(None, 0, None, None) == 1
    p = e2d.Params()
    assert L.e2d_params_from_ini(b"/nonexistent/file.ini", C.byref(p)) == 2
    assert (p.nx, p.ny, p.nStepmax) == (2, 2, 1000)  # the reference silently runs these defaults
    bad = e2d.HydroParams.from_string("[mesh]\nnx=1\nny=1\n")
    with pytest.raises(e2d.E2dError):
        HydroRun(bad)


# ------------------------------------------------------------------ BASELINE.json sizes: size-independent properties
@pytest.mark.parametrize("deck,nx,ny,steps", [("blast", 1024, 1536, 20), ("implode", 2048, 2048, 6)])
def test_large_grid_against_oracle(deck, nx, ny, steps):
    """configs[1] (blast 1024x1536, HLLC) and the reference's own 2048^2 deck: still small enough for the
    oracle to follow for a few steps — bit-exact."""
    hp, op = both_params(deck, mesh__nx=nx, mesh__ny=ny, run__nOutput=-1)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, steps)
    with HydroRun(hp) as hydro:
        st = hydro.run(steps)
        U = hydro.download(HydroRun.U if st.nStep % 2 == 0 else HydroRun.U2)
    assert st.t == t_ref
    assert_bitwise(U[INNER], U_ref[INNER], f"{deck} {nx}x{ny}")


def test_full_size_8192_fused_loop_equals_unfused_host_driven_pipeline():
    """configs[2]: four_quadrant 8192^2.  Too big for the oracle in seconds, so check what must hold at any
    size: (1) the fused device-resident run equals the unfused implementation-0 pipeline driven from the host,
    bit for bit (two independent code paths, the second one checked against the oracle above), (2) the fused
    CFL reduction equals the stand-alone ComputeDt kernel."""
    hp, _ = both_params("four_quadrant", mesh__nx=8192, mesh__ny=8192, run__nOutput=-1)
    with HydroRun(hp) as fused:
        st = fused.run(3)
        Uf = fused.download(HydroRun.U2)
        dts = fused.dt_history()
        dt_next = fused.compute_dt(1)
    hp0, _ = both_params("four_quadrant", mesh__nx=8192, mesh__ny=8192, run__nOutput=-1, other__unfusedKernels="yes")
    with HydroRun(hp0) as unfused:
        n, t, dts0 = host_loop(unfused, hp0, 3)
        U0 = unfused.download(HydroRun.U2)
        assert unfused.compute_dt(1) == dt_next
    assert t == st.t
    assert_bitwise(dts, dts0[1:], "dt")
    assert_bitwise(Uf[INNER], U0[INNER], "fused vs unfused at 8192^2")


def test_mass_and_energy_conservation_with_reflecting_walls():
    """implode 2048^2 (the reference's big deck): closed box => total mass and energy are conserved to
    round-off by the flux-form update, at any size."""
    hp, _ = both_params("implode_big", run__nOutput=-1)
    with HydroRun(hp) as hydro:
        U0 = hydro.download(HydroRun.U)
        st = hydro.run(20)
        U1 = hydro.download(HydroRun.U)
    for v in (0, 1):
        a, b = U0[v][2:-2, 2:-2].sum(), U1[v][2:-2, 2:-2].sum()
        assert abs(a - b) / a < 1e-12


# ------------------------------------------------------------------ the peer-memory slab loop (csrc/e2d_slab.cu)
def run_peer_slabs(hp, nslabs, max_steps, devices=None):
    """nslabs slab handles in THIS process (on `devices`, default all on the current device), connected with
    e2d_peer_connect_local and driven by one host thread each — e2d_run blocks while the ranks wait for each other
    on the device.  Returns (global interior [4][ny][nx], stats of rank 0, dt history of rank 0)."""
    import threading

    from euler2d_kokkos_b200 import Slab
    from euler2d_kokkos_b200.distributed import partition_rows

    import torch

    counts, starts = partition_rows(hp.ny, nslabs)
    devices = devices or [torch.cuda.current_device()] * nslabs
    runs = []
    for r in range(nslabs):
        torch.cuda.set_device(devices[r])
        runs.append(HydroRun(hp, slab=Slab(r, nslabs, counts[r], starts[r])))
    hs = (C.c_void_p * nslabs)(*[h._h for h in runs])
    e2d.check(e2d.lib().e2d_peer_connect_local(hs, nslabs), "e2d_peer_connect_local")
    stats, errs = [None] * nslabs, []

    def work(r):
        try:
            torch.cuda.set_device(devices[r])
            stats[r] = runs[r].run(max_steps)
        except Exception as ex:  # noqa: BLE001 - reported below
            errs.append((r, ex))

    th = [threading.Thread(target=work, args=(r,)) for r in range(nslabs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    n = stats[0].nStep
    assert all(s.nStep == n and s.t == stats[0].t for s in stats)
    parts = [runs[r].download(HydroRun.U if n % 2 == 0 else HydroRun.U2)[:, 2:-2, 2:-2] for r in range(nslabs)]
    dts = runs[0].dt_history()
    for h in runs:
        h.close()
    return np.concatenate(parts, axis=1), stats[0], dts


PERIODIC = dict(mesh__boundary_type_xmin=3, mesh__boundary_type_xmax=3, mesh__boundary_type_ymin=3,
                mesh__boundary_type_ymax=3)


@pytest.mark.parametrize("nslabs", [2, 3])
@pytest.mark.parametrize("deck,ov", [("implode", {}), ("four_quadrant", {}), ("implode", PERIODIC),
                                     ("shocked_bubble", {})])
def test_peer_slab_loop_matches_oracle_bitwise(deck, ov, nslabs):
    """y-slabs exchanging halo rows and CFL partials through peer stores + flags reproduce the single-domain run
    bit for bit (uneven split, reflecting / absorbing / periodic-with-wrap boundaries).  All slabs live on one
    device here, so the protocol is exercised on a single-GPU box; test_peer_slab_loop_on_two_gpus uses two."""
    nx, ny = (96, 50) if deck != "shocked_bubble" else (178, 37)
    hp, op = both_params(deck, mesh__nx=nx, mesh__ny=ny, run__nOutput=-1, **ov)
    steps = 60
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, steps)
    U, st, dts = run_peer_slabs(hp, nslabs, steps)
    assert st.nStep == n_ref and st.t == t_ref
    assert_bitwise(dts, dts_ref[1:], "dt history")
    assert_bitwise(U, U_ref[INNER], f"{deck} {nslabs} slabs")


@pytest.mark.parametrize("nslabs", [2, 3])
def test_sedov_renormalised_init_on_slabs(nslabs):
    """problem=blast with total_energy_inside: the energy inside the disc is E_tot / (volume of all disc cells), a
    reduction over the whole grid in the reference (src/HydroRunFunctors.h:1445-1463).  Slabs count their own rows and
    complete the initialisation with the integer sum over the ranks: bit-identical to the single domain, disc cut by the
    slab interfaces included."""
    hp, op = both_params("sedov_blast_2d", mesh__nx=96, mesh__ny=90, blast__radius=0.2, run__nOutput=-1)
    steps = 30
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, steps)
    U, st, dts = run_peer_slabs(hp, nslabs, steps)
    assert st.nStep == n_ref and st.t == t_ref
    assert_bitwise(dts, dts_ref[1:], "dt history")
    assert_bitwise(U, U_ref[INNER], f"sedov on {nslabs} slabs")


def test_sedov_slab_is_refused_until_the_global_count_is_known():
    import ctypes as C

    from euler2d_kokkos_b200 import Slab

    hp, _ = both_params("sedov_blast_2d", mesh__nx=64, mesh__ny=64, blast__radius=0.2, run__nOutput=-1)
    with HydroRun(hp, slab=Slab(0, 2, 32, 0)) as h:
        n, pending = C.c_ulonglong(), C.c_int()
        e2d.check(e2d.lib().e2d_blast_inside_count(h._h, C.byref(n), C.byref(pending)))
        assert pending.value == 1 and n.value > 0
        with pytest.raises(e2d.E2dError, match="Sedov"):
            h.compute_dt(0)
        e2d.check(e2d.lib().e2d_blast_renormalise(h._h, 2 * n.value))
        assert h.compute_dt(0) > 0


def test_peer_slab_loop_stops_on_tend_on_every_rank():
    hp, op = both_params("four_quadrant", mesh__nx=48, mesh__ny=48, run__nOutput=-1)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op)
    assert t_ref == op.tEnd
    U, st, _ = run_peer_slabs(hp, 2, -1)
    assert st.nStep == n_ref and st.t == t_ref
    assert_bitwise(U, U_ref[INNER], "four_quadrant to tEnd on 2 slabs")


def test_peer_slab_loop_on_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    hp, op = both_params("implode", mesh__nx=256, mesh__ny=128, run__nOutput=-1)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, 100)
    U, st, dts = run_peer_slabs(hp, 2, 100, devices=[0, 1])
    torch.cuda.set_device(0)
    assert st.nStep == n_ref and st.t == t_ref
    assert_bitwise(dts, dts_ref[1:], "dt history")
    assert_bitwise(U, U_ref[INNER], "implode on 2 GPUs")


def test_slab_run_without_peers_is_refused():
    from euler2d_kokkos_b200 import Slab

    hp, _ = both_params("implode", mesh__nx=32, mesh__ny=32)
    with HydroRun(hp, slab=Slab(0, 2, 16, 0)) as h:
        with pytest.raises(e2d.E2dError, match="peers"):
            h.run(3)


# ---- Sedov post-processing and fast output ----
def _golden_radial():
    from golden.make_golden import RADIAL_CASES
    from util import GOLDEN
    return RADIAL_CASES, np.load(os.path.join(GOLDEN, "radial_profile.npz"))


@pytest.mark.parametrize("name", ["sedov_80x64_30", "sedov_80x64_41"])
def test_radial_profile_against_reference_npy(name, tmp_path):
    """The Sedov run + ComputeRadialProfileFunctor against the .npy files the compiled reference wrote: state
    bit-exact, bin counts and distances exact, density profile to 1e-12 (the reference sums with atomics; ours is
    a deterministic tree)."""
    cases, G = _golden_radial()
    ov, steps = cases[name]
    hp, op = both_params("sedov_blast_2d", run__nOutput=-1, **ov)
    with HydroRun(hp) as hydro:
        st = hydro.run(steps)
        assert st.nStep == steps
        U = hydro.download(HydroRun.U)  # main.cpp:178 always hands U to the functor
        assert_bitwise(U[INNER], G[name + "__U"][INNER], "array the profile is taken from")
        # ghost cells take part in the profile: give U the reference's ghosts (ours may be fresher, Appendix D)
        hydro.upload(HydroRun.U, G[name + "__U"])
        dist, sums, counts = hydro.radial_profile(HydroRun.U)
        d2, s2, c2 = hydro.radial_profile(HydroRun.U)
        hydro.save_radial_profile(HydroRun.U, str(tmp_path))
    o_dist, o_sums, o_counts = oracle.radial_profile(op, G[name + "__U"])
    assert np.array_equal(counts, o_counts)
    assert_bitwise(dist, G[name + "__distances"], "distances")
    assert_bitwise(s2, sums, "deterministic sums")
    np.testing.assert_allclose(sums, o_sums, rtol=1e-12, atol=0)
    ref = G[name + "__profile"]
    with np.errstate(invalid="ignore", divide="ignore"):
        prof = sums / counts
    assert np.array_equal(np.isnan(prof), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(prof[ok], ref[ok], rtol=1e-12, atol=0)
    # the files apply() writes
    assert_bitwise(np.load(tmp_path / "sedov_blast_radial_distances.npy"), G[name + "__distances"])
    saved = np.load(tmp_path / "sedov_blast_density_profile.npy")
    assert saved.dtype == np.float64 and saved.shape == ref.shape
    np.testing.assert_allclose(saved[ok], ref[ok], rtol=1e-12, atol=0)


def test_radial_profile_larger_grid_and_bins():
    hp, op = both_params("sedov_blast_2d", mesh__nx=300, mesh__ny=170, mesh__ymax=0.6, blast__radius=0.03,
                         blast__nbins=333, run__nOutput=-1)
    with HydroRun(hp) as hydro:
        hydro.run(20)
        U = hydro.download(HydroRun.U)
        dist, sums, counts = hydro.radial_profile(HydroRun.U)
    o_dist, o_sums, o_counts = oracle.radial_profile(op, U)
    assert np.array_equal(counts, o_counts) and counts.sum() <= op.isize * op.jsize
    assert_bitwise(dist, o_dist)
    np.testing.assert_allclose(sums, o_sums, rtol=1e-12, atol=0)


def _parse_vti_appended(path):
    raw = open(path, "rb").read()
    k = raw.index(b'<AppendedData encoding="raw">')
    start = raw.index(b"_", k) + 1
    head = raw[:k].decode()
    import re
    arrays = re.findall(r'<DataArray type="Float64" Name="(\w+)" format="appended" offset="(\d+)"', head)
    out = {}
    for nm, off in arrays:
        o = start + int(off)
        nbytes = int(np.frombuffer(raw[o:o + 8], dtype=np.uint64)[0])
        out[nm] = np.frombuffer(raw[o + 8:o + 8 + nbytes], dtype=np.float64)
    return head, out, raw


@pytest.mark.parametrize("shape", [(70, 45), (1500, 1100)])
def test_fast_output_vti_appended_and_raw(shape, tmp_path):
    """Fast output: same names / extents / cell order as HydroRun::saveVTK, values bit-exact (appended raw binary);
    several row chunks at the larger size."""
    nx, ny = shape
    hp, op = both_params("four_quadrant", mesh__nx=nx, mesh__ny=ny, run__nOutput=-1,
                         output__outputDir=str(tmp_path), output__outputPrefix="fast", output__vtk_appended="yes")
    assert hp.vtkAppended == 1
    with HydroRun(hp) as hydro:
        hydro.run(3)
        U = hydro.download(HydroRun.U2)
        hydro.saveData(HydroRun.U2, 3)  # dispatches to the appended writer
        hydro.save_raw(HydroRun.U2, str(tmp_path / "snap.raw"))
    head, arrs, raw = _parse_vti_appended(tmp_path / "fast_0000003.vti")
    assert f'WholeExtent="0 {nx} 0 {ny} 0 0"' in head and 'header_type="UInt64"' in head
    assert list(arrs) == ["rho", "E", "mx", "my"]
    for v, nm in enumerate(arrs):
        assert_bitwise(arrs[nm].reshape(ny, nx), U[v, 2:-2, 2:-2], nm)
    assert raw.rstrip().endswith(b"</VTKFile>")
    snap = np.fromfile(tmp_path / "snap.raw", dtype=np.float64).reshape(4, ny, nx)
    assert_bitwise(snap, U[INNER], "raw snapshot")


def test_driver_executable_matches_reference_report(tmp_path):
    """The euler2d_b200 driver (csrc/main.cpp = the reference's main.cpp over the C++ HydroRun mirror) on a Sedov
    deck: same final step count / time as the oracle, output files written (appended .vti at the reference's
    cadence, the two radial-profile .npy files of main.cpp:175-179)."""
    import subprocess
    from euler2d_kokkos_b200.decks import write_deck
    exe = os.path.join(os.path.dirname(e2d.__file__), "euler2d_b200")
    if not os.path.exists(exe):
        pytest.skip("driver executable not built")
    ov = dict(mesh__nx=64, mesh__ny=48, blast__radius=0.05, run__nStepmax=25, run__nOutput=10,
              output__outputDir=str(tmp_path), output__vtk_appended="yes")
    ini = write_deck(str(tmp_path / "sedov.ini"), "sedov_blast_2d", **ov)
    r = subprocess.run([exe, ini], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    hp, op = both_params("sedov_blast_2d", **ov)
    U_ref, _, n_ref, t_ref = oracle.run(op, 25)
    assert f"final: nStep={n_ref} t={float(t_ref).hex()}" in r.stdout.replace("0x1.", "0x1.")
    names = sorted(f for f in os.listdir(tmp_path) if f.endswith(".vti"))
    assert names == [f"sedov_blast_2d_{k:07d}.vti" for k in (0, 10, 20, 25)]
    _, arrs, _ = _parse_vti_appended(tmp_path / names[-1])
    assert_bitwise(arrs["rho"].reshape(48, 64), U_ref[0, 2:-2, 2:-2], "rho in the last snapshot")
    prof = np.load(tmp_path / "sedov_blast_density_profile.npy")
    assert prof.shape == (op.blast_nbins,)


def test_compute_dt_cache_after_fused_step_is_exact_and_invalidated():
    """implementationVersion 2: godunov_unsplit folds the next compute_dt into its kernel (e2d_handle::d_cfl).  The
    cached value must be the bits a fresh reduction gives, and every other writer of the array must drop it."""
    hp, op = both_params("implode", mesh__nx=96, mesh__ny=64, other__implementationVersion=2)
    U_ref, dts_ref, n_ref, _ = oracle.run(op, 12)
    with HydroRun(hp) as hydro:
        n, t, dts = host_loop(hydro, hp, 12)          # compute_dt served from the cache from step 1 on
        assert_bitwise(dts, dts_ref, "dt sequence through the cache")
        cached = hydro.compute_dt(n % 2)
        U = hydro.download(HydroRun.U if n % 2 == 0 else HydroRun.U2)
        assert_bitwise(U[INNER], U_ref[INNER], "state")
        # overwrite the array through the API: the cache must not survive
        V = U.copy()
        V[1] *= 1.5
        hydro.upload(HydroRun.U if n % 2 == 0 else HydroRun.U2, V)
        fresh = hydro.compute_dt(n % 2)
        assert fresh != cached
        op_dt = op.cfl / oracle.compute_invdt(op, V)
        assert fresh == op_dt
        # a pointer handed out disables the shortcut for good
        hydro.godunov_unsplit(n, fresh)
        hydro.device_ptr(HydroRun.U)
        W = hydro.download(HydroRun.U if (n + 1) % 2 == 0 else HydroRun.U2)
        assert hydro.compute_dt((n + 1) % 2) == op.cfl / oracle.compute_invdt(op, W)


def test_save_vtk_refuses_slab_handles(tmp_path):
    """every rank of a slab run would write its piece under the same name: refused with a status code"""
    from euler2d_kokkos_b200 import Slab

    hp, _ = both_params("implode", mesh__nx=32, mesh__ny=32, output__outputDir=str(tmp_path), run__nOutput=-1)
    with HydroRun(hp, slab=Slab(0, 2, 16, 0)) as h:
        for fn in (e2d.lib().e2d_save_vtk, e2d.lib().e2d_save_vtk_appended):
            assert fn(h._h, 0, 0) == 5  # E2D_ERR_UNSUPPORTED
    assert not list(tmp_path.iterdir())


def test_tall_grid_beyond_gridDim_y_limit():
    """ny + 4 > 65535 rows: the operator-level kernels spill the row index into gridDim.z"""
    hp, op = both_params("implode", mesh__nx=8, mesh__ny=70000, mesh__ymax=8750.0, run__nOutput=-1,
                         other__unfusedKernels="yes")
    with HydroRun(hp) as hydro:
        U0 = hydro.download(HydroRun.U)
        assert_bitwise(U0, oracle.init_slab(op), "init of a 70004-row slab")
        n, t, dts = host_loop(hydro, hp, 2)
        U = hydro.download(HydroRun.U)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, 2)
    assert_bitwise(np.array(dts), dts_ref, "dt")
    assert_bitwise(U[INNER], U_ref[INNER], "tall grid, literal kernel sequence")


@pytest.mark.parametrize("deck,ov,chunk_rows,nsteps", [
    ("implode", dict(mesh__nx=96, mesh__ny=200), 16, 7),
    ("four_quadrant", dict(mesh__nx=130, mesh__ny=257), 40, 10),
    ("shocked_bubble", dict(mesh__nx=178, mesh__ny=111), 0, 5),
    ("blast", dict(mesh__nx=64, mesh__ny=96), 500, 4),          # one chunk
    ("sedov_blast_2d", dict(mesh__nx=64, mesh__ny=64, run__tEnd=1e9), 16, 6),  # (tEnd is not looked at by the march)
])
def test_march_host_pipelined_is_bit_identical(deck, ov, chunk_rows, nsteps):
    """e2d_march_host: the state lives in two pinned host buffers, every step goes through the device chunk by chunk and
    the steps are pipelined (the upload of step s+1 chases the way back of step s).  Interior, ghost cells of the final
    state, dt of every step and the time must equal the oracle's bit for bit."""
    import torch

    hp, op = both_params(deck, run__nOutput=-1, **ov)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, nsteps)
    with HydroRun(hp) as hydro:
        n = 4 * hp.jsize * hp.isize
        a = torch.empty(n, dtype=torch.float64).pin_memory()
        b = torch.empty(n, dtype=torch.float64).pin_memory()
        a.numpy()[:] = hydro.download(HydroRun.U).ravel()
        b.fill_(float("nan"))
        dts, t = hydro.march_host(a.data_ptr(), b.data_ptr(), nsteps, chunk_rows)
        res = (a if nsteps % 2 == 0 else b).numpy().reshape(4, hp.jsize, hp.isize)
    assert t == t_ref
    assert_bitwise(dts, dts_ref[1:], "dt of every step")
    assert_bitwise(res[INNER], U_ref[INNER], f"{deck}: interior after {nsteps} pipelined host steps")
    # ghost cells of the final state = the boundary fill of its own interior
    full = U_ref.copy()
    oracle.make_boundaries(op, full)
    assert_bitwise(res, full, "final state incl. ghost cells")


def test_march_host_refuses_what_it_cannot_pipeline():
    import torch

    from euler2d_kokkos_b200 import Slab

    hp, _ = both_params("implode", mesh__nx=32, mesh__ny=32, run__nOutput=-1, mesh__boundary_type_ymin=3,
                        mesh__boundary_type_ymax=3)
    buf = torch.zeros(2 * 4 * 36 * 36, dtype=torch.float64).pin_memory()
    with HydroRun(hp) as h:
        assert e2d.lib().e2d_march_host(h._h, buf.data_ptr(), buf.data_ptr() + 4 * 36 * 36 * 8, 2, 0, None, None) == 5
    hp2, _ = both_params("implode", mesh__nx=32, mesh__ny=32, run__nOutput=-1)
    with HydroRun(hp2, slab=Slab(0, 2, 16, 0)) as h:
        assert e2d.lib().e2d_march_host(h._h, buf.data_ptr(), buf.data_ptr() + 4 * 36 * 36 * 8, 2, 0, None, None) == 5


# ------------------------------------------------------------------ tapered segments (launch_fused_step)
@pytest.mark.gpu
@pytest.mark.parametrize("guide,min_rows", [("1.5", "5"), ("3", "2"), ("2", "17")])
def test_tapered_segments_are_bit_identical(guide, min_rows):
    """Large grids are split into segments of decreasing length (taper_segments, e2d_kernels.cu); the switches are read
    once per process, so the table path is forced onto small decks in a child process: five decks through the
    device-resident loop, state and dt history bitwise against the oracle (tools/parity_check.py), and the same through
    two slabs on one device (the halo-publishing instantiation counts its edge blocks from the table)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, E2D_SEG_TAPER="2", E2D_SEG_GUIDE=guide, E2D_SEG_MIN_ROWS=min_rows)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "parity_check.py"), "--slabs"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("state bitwise: True") >= 6, r.stdout
