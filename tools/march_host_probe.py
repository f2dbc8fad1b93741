"""Wall time per step of e2d_march_host against the step count and the chunk height (8192^2 four_quadrant).
Usage: python tools/march_host_probe.py [n]"""
import sys
import time

import torch

import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
hp = e2d.HydroParams.from_string(deck_text("four_quadrant", mesh__nx=n, mesh__ny=n, run__tEnd=1e9))
hyd = e2d.HydroRun(hp)
a = torch.empty(4 * (n + 4) * (n + 4), dtype=torch.float64).pin_memory()
b = torch.empty_like(a).pin_memory()
a.copy_(torch.from_numpy(hyd.download()).reshape(-1))
for chunk in (0, 64, 128, 256, 512, 1024):
    for steps in (10, 20, 40):
        hyd.march_host(a.data_ptr(), b.data_ptr(), 2, chunk_rows=chunk)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hyd.march_host(a.data_ptr(), b.data_ptr(), steps, chunk_rows=chunk)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"chunk_rows {chunk:5d} steps {steps:3d}: {dt / steps * 1e3:7.2f} ms/step  ({n * n * steps / dt * 1e-6:7.1f} Mcell/s)",
              flush=True)
