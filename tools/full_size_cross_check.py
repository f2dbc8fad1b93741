"""Full-size parity without an oracle twin: the same deck advanced (a) by the N ranks of a torchrun job through the
peer-memory loop (y-slabs, tapered segments, halo rows over NVLink) and (b) by rank 0 alone holding the whole grid
(single-GPU loop, its own tapered segments) must agree bit for bit, dt history included — the grids are far beyond what
the CPU oracle finishes in seconds, and the small-deck tests pin (b)'s code path to the oracle.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/full_size_cross_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text
from euler2d_kokkos_b200.distributed import PeerSlabRun, partition_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
CASES = [("four_quadrant", 8192, 4096 * world, 24), ("shocked_bubble", 16384, 2048 * world, 16), ("blast", 4096, 8192, 30),
         ("implode_big", 6000, 3000 + 2 * world, 20)]
ok = True
for deck, nx, ny, steps in CASES:
    hp = e2d.HydroParams.from_string(deck_text(deck, mesh__nx=nx, mesh__ny=ny, run__nOutput=-1, run__nStepmax=steps,
                                               run__tEnd=1e9))
    run = PeerSlabRun(hp, device=dev)
    st = run.run(steps)
    mine = run.current(st.nStep)[:, 2:-2, 2:-2].numpy()
    dts = run.hydro.dt_history()
    counts, starts = partition_rows(ny, world)
    same = True
    if rank == 0:
        with e2d.HydroRun(hp) as whole:
            st1 = whole.run(steps)
            ref = whole.download(st1.nStep % 2)[:, 2:-2, 2:-2]
            dts1 = whole.dt_history()
        same_dt = np.array_equal(dts, dts1) and st.nStep == st1.nStep and st.t == st1.t
    for r in range(world):
        # every rank's slab against rank 0's whole-grid result, one slab at a time (rank 0 holds the reference)
        if r == 0:
            if rank == 0:
                same = same and np.array_equal(mine.view(np.uint64), np.ascontiguousarray(ref[:, :counts[0]]).view(np.uint64))
        else:
            if rank == r:
                dist.send(torch.from_numpy(np.ascontiguousarray(mine)).to(dev), dst=0)
            elif rank == 0:
                buf = torch.empty((4, counts[r], nx), dtype=torch.float64, device=dev)
                dist.recv(buf, src=r)
                part = buf.cpu().numpy()
                same = same and np.array_equal(part.view(np.uint64),
                                               np.ascontiguousarray(ref[:, starts[r]:starts[r] + counts[r]]).view(np.uint64))
    if rank == 0:
        print(f"{world} GPUs vs 1 GPU  {deck:15s} {nx}x{ny}  steps {st.nStep}  state bitwise {same}  dt history + step count + t bitwise {same_dt}",
              flush=True)
        ok = ok and same and same_dt
    dist.barrier()  # nobody stores into a peer any more
    run.hydro.close()
    del run
    torch.cuda.empty_cache()
    dist.barrier()
if rank == 0:
    print("ALL BITWISE" if ok else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
