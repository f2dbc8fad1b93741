"""BASELINE.json configs[3] / configs[4] under torchrun: a deck at a given size split into y-slabs over the ranks,
stepped by the device-resident peer-memory loop (PeerSlabRun).  Prints one JSON line per (deck, mode) on rank 0.

usage: torchrun --nproc-per-node N tools/scale_run.py deck nx ny steps [strict|fast]
Also reports the relative change of total mass and energy over the run (zero to round-off for decks with reflecting
walls: a size-independent property, since these sizes have no single-GPU or CPU twin to compare with).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text
from euler2d_kokkos_b200.distributed import PeerSlabRun

deck, nx, ny, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "strict"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
hp = e2d.HydroParams.from_string(deck_text(deck, mesh__nx=nx, mesh__ny=ny, run__nOutput=-1, run__nStepmax=10 ** 8,
                                           run__tEnd=1e9, other__arithmetic=mode))
run = PeerSlabRun(hp, device=dev)


class _DevArray:
    """the slab in device memory as a torch tensor (no copy), through __cuda_array_interface__"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False), "version": 2}


def totals(n):
    h = run.hydro
    U = torch.as_tensor(_DevArray(h.device_ptr(n % 2), (4, h.jsize_loc, h.isize)), device=dev)
    t = torch.stack([U[0, 2:-2, 2:-2].sum(), U[1, 2:-2, 2:-2].sum()])
    dist.all_reduce(t)
    return t.cpu().numpy()


tot0 = totals(0)
W = 3
run.run(W)
dist.barrier()
torch.cuda.synchronize()
st = run.run(W + steps)
dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([st.seconds], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
tot1 = totals(W + steps)
if rank == 0:
    sec = float(t.item())
    print(json.dumps({"deck": deck, "nx": nx, "ny": ny, "n_gpus": world, "arithmetic": mode, "steps": steps,
                      "ms_per_step": sec / steps * 1e3, "Mcell_updates_per_s": nx * ny * steps / sec * 1e-6,
                      "t": st.t, "mass_rel_change": float(abs(tot1[0] - tot0[0]) / tot0[0]),
                      "energy_rel_change": float(abs(tot1[1] - tot0[1]) / tot0[1]),
                      "state_GB_per_gpu": 2 * 4 * 8 * (nx + 4) * (ny // world + 4) * 1e-9}), flush=True)
run.close()
dist.destroy_process_group()
