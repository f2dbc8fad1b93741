"""Development aid: e2d_step_host_streamed at 8192^2 for several chunk sizes (rows per chunk)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text

mode = sys.argv[1] if len(sys.argv) > 1 else "strict"
hp = e2d.HydroParams.from_string(deck_text("four_quadrant", mesh__nx=8192, mesh__ny=8192, run__nOutput=-1,
                                           other__arithmetic=mode))
h = e2d.HydroRun(hp)
n = 4 * h.jsize_loc * h.isize
a = torch.empty(n, dtype=torch.float64).pin_memory()
b = torch.empty(n, dtype=torch.float64).pin_memory()
e2d.check(e2d.lib().e2d_download(h._h, e2d.E2D_U, a.data_ptr(), e2d.LAYOUT_SOA))
for rows in (0, 512, 256, 128, 64, 32):
    used, dt = h.step_host_streamed(a.data_ptr(), b.data_ptr(), 0.0, rows)
    a, b = b, a
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        used, dt = h.step_host_streamed(a.data_ptr(), b.data_ptr(), dt, rows)
        a, b = b, a
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 5
    print(f"{mode} chunk_rows={rows:4d}: {t*1e3:7.2f} ms/step  {8192*8192/t*1e-6:7.1f} Mcell/s", flush=True)
