"""The UNMODIFIED reference built with real Kokkos/CUDA for Blackwell (baseline/build_ref_cuda.sh ->
baseline/_ref/euler2d_kokkos_cuda) on the GPU, for the bench workload: "the recompiled generic kernels to beat".

usage: python tools/ref_cuda_perf.py [deck nx ny steps]     (all three implementationVersions are run)
Prints one JSON line per implementation: the program's own `Perf` line (isize*jsize*nStep / total time, ghosts
counted, main.cpp:202-203) and the same with nx*ny.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from euler2d_kokkos_b200.decks import deck_text

EXE = os.path.join(ROOT, "baseline", "_ref", "euler2d_kokkos_cuda")


def run_reference_cuda(deck="four_quadrant", nx=8192, ny=8192, steps=20, impl=0, timeout=600):
    if not os.path.exists(EXE):
        return None
    text = deck_text(deck, mesh__nx=nx, mesh__ny=ny, run__nStepmax=steps, run__tEnd=1e9, run__nOutput=-1,
                     other__implementationVersion=impl)
    with tempfile.TemporaryDirectory() as d:
        ini = os.path.join(d, "deck.ini")
        with open(ini, "w") as f:
            f.write(text)
        out = subprocess.run([EXE, ini], capture_output=True, text=True, timeout=timeout, cwd=d)
    if out.returncode != 0:
        return {"error": (out.stderr or out.stdout)[-400:]}
    m = re.search(r"Perf\s*:\s*([0-9.]+)", out.stdout)
    t = re.search(r"total\s+time\s*:\s*([0-9.]+)", out.stdout)
    tg = re.search(r"godunov\s+time\s*:\s*([0-9.]+)", out.stdout)
    perf = float(m.group(1)) if m else None
    return {"impl": "reference (Kokkos 5.1.0 CUDA backend, -arch=sm_100, unmodified sources)", "deck": deck, "nx": nx,
            "ny": ny, "steps": steps, "implementationVersion": impl, "perf_line_Mcell_per_s": perf,
            "Mcell_updates_per_s": perf * (nx * ny) / ((nx + 4) * (ny + 4)) if perf else None,
            "total_s": float(t.group(1)) if t else None, "godunov_s": float(tg.group(1)) if tg else None}


if __name__ == "__main__":
    deck = sys.argv[1] if len(sys.argv) > 1 else "four_quadrant"
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    ny = int(sys.argv[3]) if len(sys.argv) > 3 else nx
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    for impl in (0, 1, 2):
        print(json.dumps(run_reference_cuda(deck, nx, ny, steps, impl)), flush=True)
