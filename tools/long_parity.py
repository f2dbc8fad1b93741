"""Long-run bitwise parity of the device-resident loop against the CPU oracle (one-off evidence, not part of the test-suite:
the oracle needs tens of seconds for these).  Thousands of steps on the reference's decks take the strict kernel through the
regimes the short tests barely touch: the exponentially small, denormal-range tails ahead of shocks, where the fast-path
guards reject and whole phases are recomputed with the plain IEEE operators; the tEnd clamp; strong shocks (Sedov).

usage: python tools/long_parity.py [quick]
"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
import oracle
from euler2d_kokkos_b200.decks import deck_text

CASES = [("implode", dict(), 3000), ("blast", dict(mesh__nx=256, mesh__ny=384), 2000),
         ("four_quadrant", dict(), -1), ("discontinuity", dict(), -1), ("shocked_bubble", dict(), 2000),
         ("sedov_blast_2d", dict(), 1500), ("blast", dict(mesh__nx=1024, mesh__ny=1536), 300)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    CASES = [(d, o, min(s, 200) if s > 0 else 200) for d, o, s in CASES[:4]]
ok = True
for deck, ov, steps in CASES:
    text = deck_text(deck, run__nOutput=-1, **(dict(ov, run__nStepmax=steps) if steps > 0 else ov))
    hp = e2d.HydroParams.from_string(text)
    with tempfile.TemporaryDirectory() as td:
        ini = os.path.join(td, "d.ini")
        open(ini, "w").write(text)
        op = oracle.params_from_ini(ini)
    t0 = time.time()
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op)
    t_cpu = time.time() - t0
    with e2d.HydroRun(hp) as h:
        st = h.run()
        U = h.download(st.nStep % 2)
        dts = h.dt_history()
    same = np.array_equal(U[:, 2:-2, 2:-2].view(np.uint64), np.ascontiguousarray(U_ref[:, 2:-2, 2:-2]).view(np.uint64))
    same_dt = len(dts) == n_ref and np.array_equal(dts, dts_ref[1:])
    tiny = int(((np.abs(U_ref) < 1e-250) & (U_ref != 0)).sum())
    print(f"{deck:16s} {hp.nx}x{hp.ny}  steps {st.nStep} (oracle {n_ref})  t {st.t!r} == {t_ref!r}: {st.t == t_ref}  state bitwise {same}  "
          f"dt bitwise {same_dt}  values below 1e-250 in the final state: {tiny}  (oracle {t_cpu:.1f} s, GPU loop {st.seconds*1e3:.0f} ms)",
          flush=True)
    ok = ok and same and same_dt and st.nStep == n_ref and st.t == t_ref
print("ALL BITWISE" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
