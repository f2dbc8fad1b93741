"""Golden ascii .vti files written by the UNMODIFIED reference program (baseline/_ref/euler2d_kokkos_omp: src/main.cpp
on real Kokkos 5.1.0 / OpenMP, baseline/build_ref_omp.sh) for small decks -> tests/golden/vti/.

    python tests/golden/make_golden_vti.py

Run in the build container (needs /root/reference to build the program).  The decks are written next to the files so
that the tests read the very same text.  tests/test_gpu_refmain.py byte-compares what e2d_save_vtk writes — driven
by the product's own loop AND by the reference's own main.cpp compiled against include/euler2d_compat.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from euler2d_kokkos_b200.decks import deck_text  # noqa: E402

EXE = os.path.join(ROOT, "baseline", "_ref", "euler2d_kokkos_omp")
OUT = os.path.join(ROOT, "tests", "golden", "vti")
CASES = {
    # name: (deck, overrides) — nStepmax / nOutput chosen so that 3 files are written (step 0, an odd step, the end)
    "implode_24x16": ("implode", dict(mesh__nx=24, mesh__ny=16, run__nStepmax=14, run__nOutput=7,
                                      output__outputPrefix="implode_24x16")),
    "shocked_bubble_40x12": ("shocked_bubble", dict(mesh__nx=40, mesh__ny=12, mesh__xmax=0.4, mesh__ymax=0.12,
                                                    run__nStepmax=9, run__nOutput=3,
                                                    output__outputPrefix="shocked_bubble_40x12")),
}

if __name__ == "__main__":
    if not os.path.exists(EXE):
        sys.exit("build baseline/_ref/euler2d_kokkos_omp first (sh baseline/build_ref_omp.sh)")
    os.makedirs(OUT, exist_ok=True)
    for name, (deck, ov) in CASES.items():
        text = deck_text(deck, **ov)
        with tempfile.TemporaryDirectory() as td:
            ini = os.path.join(td, name + ".ini")
            open(ini, "w").write(text)
            env = dict(os.environ, OMP_NUM_THREADS="2")
            res = subprocess.run([EXE, ini], cwd=td, capture_output=True, text=True, env=env, check=True)
            files = sorted(f for f in os.listdir(td) if f.endswith(".vti"))
            assert files, res.stdout[-2000:]
            for f in files:
                shutil.copy(os.path.join(td, f), os.path.join(OUT, f))
            shutil.copy(ini, os.path.join(OUT, name + ".ini"))
            report = [ln for ln in res.stdout.splitlines() if ln.startswith("time step=") or ln.startswith("Output results")]
            open(os.path.join(OUT, name + ".stdout.txt"), "w").write("\n".join(report) + "\n")
            print(name, files)
