"""Development aid: bitwise check of the device-resident loop of the library named by E2D_LIB_PATH against the CPU oracle on
a few even-sized decks (A/B variants whose kernels have stricter alignment needs than the shipped one).
--slabs: also two y-slabs on this device through the peer-memory loop (tests/test_gpu_hydro_run.py::run_peer_slabs)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
import oracle
from euler2d_kokkos_b200.decks import deck_text

ok = True
for deck, ov, steps in (("implode", dict(mesh__nx=256, mesh__ny=128), 100), ("four_quadrant", dict(mesh__nx=200, mesh__ny=120), 60),
                        ("blast", dict(mesh__nx=128, mesh__ny=192), 80), ("shocked_bubble", dict(mesh__nx=444, mesh__ny=88), 50),
                        ("implode", dict(mesh__nx=1000, mesh__ny=700), 30)):
    text = deck_text(deck, run__nOutput=-1, **ov)
    hp = e2d.HydroParams.from_string(text)
    with tempfile.TemporaryDirectory() as td:
        ini = os.path.join(td, "d.ini")
        open(ini, "w").write(text)
        op = oracle.params_from_ini(ini)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, steps)
    with e2d.HydroRun(hp) as h:
        st = h.run(steps)
        U = h.download(st.nStep % 2)
        dts = h.dt_history()
    same = np.array_equal(U[:, 2:-2, 2:-2].view(np.uint64), np.ascontiguousarray(U_ref[:, 2:-2, 2:-2]).view(np.uint64))
    same_dt = np.array_equal(dts, dts_ref[1:])
    print(deck, ov, "state bitwise:", same, "dt bitwise:", same_dt, flush=True)
    ok = ok and same and same_dt
    if "--slabs" in sys.argv and deck in ("four_quadrant", "shocked_bubble"):
        sys.path.insert(0, "tests")
        from test_gpu_hydro_run import run_peer_slabs

        Us, st, dts = run_peer_slabs(hp, 2, steps)
        same = np.array_equal(Us.view(np.uint64), np.ascontiguousarray(U_ref[:, 2:-2, 2:-2]).view(np.uint64))
        same_dt = np.array_equal(dts, dts_ref[1:])
        print(deck, ov, "two slabs: state bitwise:", same, "dt bitwise:", same_dt, flush=True)
        ok = ok and same and same_dt
sys.exit(0 if ok else 1)
