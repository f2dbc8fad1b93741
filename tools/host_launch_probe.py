import sys, time
sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text
import torch
for nx, ny in ((256, 128), (1024, 1536)):
    hp = e2d.HydroParams.from_string(deck_text("implode", mesh__nx=nx, mesh__ny=ny, run__nOutput=-1, run__nStepmax=100000, run__tEnd=1e9))
    h = e2d.HydroRun(hp)
    h.run(10)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); st = h.run(10 + 640); t1 = time.perf_counter()
    print(f"{nx}x{ny}: wall {1e6*(t1-t0)/640:.2f} us/step, events {1e6*st.seconds/640:.2f} us/step", flush=True)
    # host-driven API: godunov_unsplit without sync (enqueue cost of one fused launch + ghost copy)
    hp2 = e2d.HydroParams.from_string(deck_text("implode", mesh__nx=nx, mesh__ny=ny, run__nOutput=-1, other__implementationVersion=2))
    h2 = e2d.HydroRun(hp2)
    dt = h2.compute_dt(0)
    h2.synchronize()
    t0 = time.perf_counter()
    for n in range(200):
        h2.godunov_unsplit(n, dt * 0.5)
    t1 = time.perf_counter()
    h2.synchronize(); t2 = time.perf_counter()
    print(f"   godunov_unsplit enqueue: {1e6*(t1-t0)/200:.2f} us/call (2 launches), drained after {1e6*(t2-t0)/200:.2f} us/call", flush=True)
