"""Development aid: where the time of a small-grid step goes.  Needs a library built with -DE2D_TIMELINE=1
(tools/build_variants.sh timeline "-DE2D_TIMELINE=1") and E2D_LIB_PATH pointing at it: every block of the single-GPU loop
kernel stamps %globaltimer at its milestones; this prints, for the last complete steps of a run, each milestone relative to
the moment the step's first block passed its dependency wait (min / median / max over the blocks, in microseconds).

usage: E2D_LIB_PATH=build/variants/timeline/libeuler2d_b200.so python tools/step_timeline.py [deck nx ny steps]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text

deck = sys.argv[1] if len(sys.argv) > 1 else "implode"
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ny = int(sys.argv[3]) if len(sys.argv) > 3 else 128
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 203
hp = e2d.HydroParams.from_string(deck_text(deck, mesh__nx=nx, mesh__ny=ny, run__tEnd=1e9, run__nStepmax=steps,
                                           run__nOutput=-1))
h = e2d.HydroRun(hp)
h.run(steps)
lib = e2d.lib()
buf = np.zeros((4, 1024, 8), dtype=np.uint64)
lib.e2d_debug_timeline.argtypes = [C.c_void_p, C.c_int]
rc = lib.e2d_debug_timeline(buf.ctypes.data, 1024)
assert rc == 0, rc
names = ["block resident", "dependency wait over", "step opened (dt known)", "prologue done (3 rows loaded + converted)",
         "marching loop done", "CFL tail + block maximum done", "epilogue done (boundary push)"]
print(f"{deck} {nx}x{ny}, {steps} steps; blocks with marks per kept step:", [(int((buf[k, :, 2] > 0).sum())) for k in range(4)])
order = sorted(range(4), key=lambda k: buf[k][buf[k, :, 2] > 0][:, 1].min() if (buf[k, :, 2] > 0).any() else 0)
prev_open = None
for k in order:
    m = buf[k][buf[k, :, 2] > 0].astype(np.int64)
    if not len(m):
        continue
    t0 = m[:, 1].min()
    print(f"-- kept step slot {k}: {len(m)} blocks" + (f"; {(t0 - prev_open) * 1e-3:.2f} us after the previous step's first block passed its wait"
                                                   if prev_open else ""))
    for c, nm in enumerate(names):
        v = (m[:, c] - t0) * 1e-3
        print(f"   {nm:44s} min {v.min():7.2f}  median {np.median(v):7.2f}  max {v.max():7.2f}")
    prev_open = t0

# the last kept step, block by block: does the marching loop's duration follow the SM's load or the block's place in the grid?
k = order[-1]
m = buf[k].astype(np.int64)
valid = m[:, 2] > 0
nb = int(valid.sum())
loop = (m[:nb, 4] - m[:nb, 3]) * 1e-3
epi = (m[:nb, 6] - m[:nb, 5]) * 1e-3
sm = m[:nb, 7] & 0xffff
bxs = (m[:nb, 7] >> 16) & 0xffff
segs = m[:nb, 7] >> 32
per_sm = np.bincount(sm, minlength=148)
print("blocks per SM histogram:", np.bincount(per_sm).tolist())
for c in sorted(set(per_sm[sm].tolist())):
    sel = per_sm[sm] == c
    print(f"   SMs holding {c} block(s): {int(sel.sum())} blocks, loop median {np.median(loop[sel]):.2f} us (min {loop[sel].min():.2f}, max {loop[sel].max():.2f}); "
          f"epilogue median {np.median(epi[sel]):.2f} (max {epi[sel].max():.2f})")
end = (m[:nb, 6] - m[:nb, 1].min()) * 1e-3
nbx, nseg = int(bxs.max()) + 1, int(segs.max()) + 1
for what, arr in (("marching loop duration", loop), ("epilogue duration", epi), ("block finished at (after the first block passed its wait)", end)):
    grid = np.full((nseg, nbx), np.nan)
    grid[segs, bxs] = arr
    print(f"{what}, us (rows: segments from the bottom; columns: column blocks):")
    for seg in range(nseg):
        print(f"   seg {seg:3d}: " + " ".join(f"{v:5.1f}" for v in grid[seg]))
