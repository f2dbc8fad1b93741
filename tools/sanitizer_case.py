"""compute-sanitizer case: a few steps of every loop flavour on a small grid with all boundary kinds.
  compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitizer_case.py
  compute-sanitizer --tool memcheck python tools/sanitizer_case.py
Covers the single-GPU one-launch loop (boundary push epilogue, in-kernel bookkeeping), the two-launch loop
(E2D_TWO_LAUNCH=1 is NOT set here; the peer loop below uses the same boundary kernel), strict and fast arithmetic,
periodic and reflecting walls, and the peer-memory slab loop with two slabs on one device."""
import ctypes as C
import sys
import threading

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200 import Slab
from euler2d_kokkos_b200.decks import deck_text
from euler2d_kokkos_b200.distributed import partition_rows

PERIODIC = dict(mesh__boundary_type_xmin=3, mesh__boundary_type_xmax=3, mesh__boundary_type_ymin=3, mesh__boundary_type_ymax=3)
for mode in ("strict", "fast"):
    for ov in ({}, PERIODIC):
        hp = e2d.HydroParams.from_string(deck_text("implode", mesh__nx=300, mesh__ny=70, run__nOutput=-1, run__nStepmax=4,
                                                   other__arithmetic=mode, **ov))
        with e2d.HydroRun(hp) as h:
            st = h.run()
            print(mode, "periodic" if ov else "walls", st.nStep, st.t)
# two slabs on one device through the peer loop
hp = e2d.HydroParams.from_string(deck_text("implode", mesh__nx=300, mesh__ny=70, run__nOutput=-1, run__nStepmax=4))
counts, starts = partition_rows(hp.ny, 2)
runs = [e2d.HydroRun(hp, slab=Slab(r, 2, counts[r], starts[r])) for r in range(2)]
hs = (C.c_void_p * 2)(*[h._h for h in runs])
e2d.check(e2d.lib().e2d_peer_connect_local(hs, 2))
out = [None, None]
th = [threading.Thread(target=lambda r=r: out.__setitem__(r, runs[r].run(4))) for r in range(2)]
[t.start() for t in th]
[t.join() for t in th]
print("slabs", out[0].nStep, out[0].t, out[1].t)
[h.close() for h in runs]
