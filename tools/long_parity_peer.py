"""Long-run bitwise parity of the multi-GPU peer-memory loop (PeerSlabRun) against the CPU oracle, under torchrun:
thousands of steps of halo rows and invDt partials travelling as NVLink peer stores behind system-scope flags, on small
grids where a step is tens of microseconds — the regime in which a flaw of the flag protocol would show.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/long_parity_peer.py
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text
from euler2d_kokkos_b200.distributed import PeerSlabRun

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
PERIODIC = dict(mesh__boundary_type_xmin=3, mesh__boundary_type_xmax=3, mesh__boundary_type_ymin=3, mesh__boundary_type_ymax=3)
CASES = [("implode", dict(), 3000), ("implode", dict(mesh__nx=96, mesh__ny=64, **PERIODIC), 3000),
         ("blast", dict(mesh__nx=256, mesh__ny=384), 1500), ("four_quadrant", dict(), -1), ("shocked_bubble", dict(), 1500),
         ("sedov_blast_2d", dict(), 1000)]
ok = True
for deck, ov, steps in CASES:
    text = deck_text(deck, run__nOutput=-1, **(dict(ov, run__nStepmax=steps) if steps > 0 else ov))
    hp = e2d.HydroParams.from_string(text)
    run = PeerSlabRun(hp, device=dev)
    st = run.run(-1)
    U = run.gather_interior(st.nStep)
    dts = run.hydro.dt_history()
    run.close()
    if rank == 0:
        import oracle

        with tempfile.TemporaryDirectory() as td:
            ini = os.path.join(td, "d.ini")
            open(ini, "w").write(text)
            op = oracle.params_from_ini(ini)
        t0 = time.time()
        U_ref, dts_ref, n_ref, t_ref = oracle.run(op)
        same = np.array_equal(U.cpu().numpy().view(np.uint64), np.ascontiguousarray(U_ref[:, 2:-2, 2:-2]).view(np.uint64))
        same_dt = len(dts) == n_ref and np.array_equal(dts, dts_ref[1:])
        print(f"{world} GPUs  {deck:16s} {hp.nx}x{hp.ny} {'periodic' if 'mesh__boundary_type_ymin' in ov else ''}  steps {st.nStep} "
              f"(oracle {n_ref})  final t equal {st.t == t_ref}  state bitwise {same}  dt bitwise {same_dt}  "
              f"({st.seconds / max(st.nStep, 1) * 1e6:.1f} us per step)", flush=True)
        ok = ok and same and same_dt and st.nStep == n_ref and st.t == t_ref
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
if rank == 0:
    print("ALL BITWISE" if ok else "MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if bool(flag.item()) else 1)
