"""Development aid: `arithmetic=fast` against the strict (bit-identical-to-the-reference) fused step on the GPU:
per-variable relative L1 / Linf deviation after N steps, dt sequences, and the throughput of both.

usage: python tools/fast_vs_strict.py deck nx ny steps
"""
import sys

import numpy as np

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text
from euler2d_kokkos_b200.parity import state_deviation

deck = sys.argv[1] if len(sys.argv) > 1 else "four_quadrant"
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ny = int(sys.argv[3]) if len(sys.argv) > 3 else nx
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100

out = {}
for mode in ("strict", "fast"):
    hp = e2d.HydroParams.from_string(deck_text(deck, mesh__nx=nx, mesh__ny=ny, run__nOutput=-1, run__nStepmax=steps,
                                               other__arithmetic=mode))
    h = e2d.HydroRun(hp)
    st = h.run()
    U = h.download(e2d.HydroRun.U if st.nStep % 2 == 0 else e2d.HydroRun.U2)
    out[mode] = (U[:, 2:-2, 2:-2].copy(), h.dt_history().copy(), st)
    print(f"{mode:6s} {deck} {nx}x{ny}: {st.nStep} steps, t={st.t!r}, {st.seconds*1e3:.2f} ms -> "
          f"{st.nStep*nx*ny/st.seconds*1e-6:.1f} Mcell/s", flush=True)
    h.close()
Us, ds, ss = out["strict"]
Uf, df, sf = out["fast"]
print("steps", ss.nStep, sf.nStep, " t rel", abs(sf.t - ss.t) / abs(ss.t),
      " dt rel max", float(np.max(np.abs(df - ds) / ds)) if len(df) == len(ds) else "len differs")
for name, l1, linf in state_deviation(Uf, Us):
    print(f"  {name:4s} rel L1 {l1:.3e}  rel Linf {linf:.3e}")
