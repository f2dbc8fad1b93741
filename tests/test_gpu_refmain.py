"""The drop-in boundary exercised by the reference's own code and files.

* `saveData` (e2d_save_vtk) writes ascii .vti files that are BYTE-IDENTICAL to the ones the unmodified reference
  program (src/main.cpp on real Kokkos/OpenMP) wrote for the same decks: tests/golden/vti/ (make_golden_vti.py).
* oracle/_ref/ref_main_b200 is the reference's src/main.cpp, unmodified, compiled against include/euler2d_compat and
  linked with libeuler2d_b200.so (oracle/Makefile).  Run on the same decks it must print the reference's report lines
  and leave the same files.  (Built in the build container, where /root/reference exists; it travels to the GPU box.)
* the product's own driver (csrc/main.cpp -> euler2d_kokkos_b200/euler2d_b200) is built and run unconditionally.
* NVTX regions (Kokkos::Profiling::pushRegion names of src/HydroRun.h:242-360) are opened and balanced in profile mode.
"""
import ctypes as C
import glob
import os
import shutil
import subprocess

import pytest

import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200 import HydroRun

pytestmark = pytest.mark.gpu
pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VTI = os.path.join(ROOT, "tests", "golden", "vti")
CASES = ["implode_24x16", "shocked_bubble_40x12"]
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "ref_main_b200")
OWN_MAIN = os.path.join(ROOT, "euler2d_kokkos_b200", "euler2d_b200")


def golden_files(case):
    return sorted(glob.glob(os.path.join(VTI, case + "_*.vti")))


def assert_same_files(case, out_dir):
    gold = golden_files(case)
    assert len(gold) >= 3
    for g in gold:
        mine = os.path.join(out_dir, os.path.basename(g))
        assert os.path.exists(mine), f"{os.path.basename(g)} was not written"
        a, b = open(mine, "rb").read(), open(g, "rb").read()
        if a != b:
            la, lb = a.split(b"\n"), b.split(b"\n")
            k = next((i for i, (x, y) in enumerate(zip(la, lb)) if x != y), min(len(la), len(lb)))
            raise AssertionError(f"{os.path.basename(g)} differs from the reference's file at line {k + 1}:\n"
                                 f"  ours: {la[k][:160]!r}\n  ref : {lb[k][:160]!r}")


@pytest.mark.parametrize("case", CASES)
def test_save_vtk_is_byte_identical_to_the_reference_files(case, tmp_path):
    """python host loop (main.cpp:100-166) -> e2d_save_vtk; files compared byte for byte with the reference's"""
    text = open(os.path.join(VTI, case + ".ini")).read()
    hp = e2d.HydroParams.from_string(text + f"\n[output]\noutputDir={tmp_path}\n")
    with HydroRun(hp) as hydro:
        hydro.make_boundaries(HydroRun.U)
        hydro.make_boundaries(HydroRun.U2)
        t, n = 0.0, 0
        while t < hp.tEnd and n < hp.nStepmax:
            if n % hp.nOutput == 0:
                hydro.saveData(HydroRun.U if n % 2 == 0 else HydroRun.U2, n, "U")
            dt = hydro.compute_dt(n % 2)
            if t + dt > hp.tEnd:
                dt = hp.tEnd - t
            hydro.godunov_unsplit(n, dt)
            n += 1
            t += dt
        hydro.saveData(HydroRun.U if n % 2 == 0 else HydroRun.U2, n, "U")
    assert_same_files(case, str(tmp_path))


def run_program(exe, case, tmp_path, extra=()):
    ini = tmp_path / (case + ".ini")
    shutil.copy(os.path.join(VTI, case + ".ini"), ini)
    res = subprocess.run([exe, str(ini), *extra], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:] + res.stdout[-2000:]
    return res.stdout


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="oracle/_ref/ref_main_b200 is built where /root/reference exists")
@pytest.mark.parametrize("case", CASES)
def test_reference_main_cpp_unmodified_drives_the_library(case, tmp_path):
    out = run_program(REF_MAIN, case, tmp_path)
    want = open(os.path.join(VTI, case + ".stdout.txt")).read().splitlines()
    got = [ln for ln in out.splitlines() if ln.startswith("time step=") or ln.startswith("Output results")]
    assert got == want, "the loop's report lines (dt, t of every 10th step and of every output) differ"
    assert "Perf                 :" in out and "boundaries      time" in out
    assert_same_files(case, str(tmp_path))


@pytest.mark.parametrize("case", CASES)
def test_own_driver_is_built_and_matches_the_reference_files(case, tmp_path):
    """csrc/main.cpp (the reference's program on HydroRun.hpp) — built by `make`, never skipped"""
    if not os.path.exists(OWN_MAIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "euler2d_kokkos_b200", "csrc"), "../euler2d_b200"])
    out = run_program(OWN_MAIN, case, tmp_path)
    want = open(os.path.join(VTI, case + ".stdout.txt")).read().splitlines()
    got = [ln for ln in out.splitlines() if ln.startswith("time step=") or ln.startswith("Output results")]
    assert got == want
    assert_same_files(case, str(tmp_path))


def test_profile_regions_are_opened_and_balanced():
    L = e2d.lib()
    n0, depth = C.c_ulonglong(), C.c_int()
    L.e2d_profile_enable(1)
    try:
        L.e2d_profile_stats(C.byref(n0), C.byref(depth))
        hp = e2d.HydroParams.from_string(open(os.path.join(VTI, "implode_24x16.ini")).read() + "\n[run]\nnOutput=-1\n")
        with HydroRun(hp) as hydro:
            L.e2d_profile_push(b"main_loop")
            for n in range(4):
                hydro.godunov_unsplit(n, hydro.compute_dt(n % 2))
            L.e2d_profile_pop()
        n1 = C.c_ulonglong()
        on = L.e2d_profile_stats(C.byref(n1), C.byref(depth))
        assert on == 1 and depth.value == 0
        # per step: compute_dt, make_boundaries, hydro_impl0, compute_fluxes (+ main_loop once)
        assert n1.value - n0.value == 4 * 4 + 1
    finally:
        L.e2d_profile_enable(0)
    n2 = C.c_ulonglong()
    L.e2d_profile_push(b"ignored")
    L.e2d_profile_pop()
    assert L.e2d_profile_stats(C.byref(n2), C.byref(depth)) == 0 and n2.value == n1.value
