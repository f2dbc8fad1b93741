"""One process per GPU (the bench.py launch shape): PeerSlabRun exchanges CUDA IPC handles through
torch.distributed once and then steps through libeuler2d_b200.so's own peer-memory loop.  Needs >= 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, os.environ["E2D_ROOT"]); sys.path.insert(0, os.path.join(os.environ["E2D_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import euler2d_kokkos_b200 as e2d
from euler2d_kokkos_b200.decks import deck_text
from euler2d_kokkos_b200.distributed import PeerSlabRun
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
hp = e2d.HydroParams.from_string(deck_text("implode", mesh__nx=256, mesh__ny=128, run__nOutput=-1))
run = PeerSlabRun(hp)
st = run.run(100)
U = run.gather_interior(st.nStep)
if rank == 0:
    np.save(os.environ["E2D_OUT"], U.cpu().numpy())
    print(json.dumps({"nStep": st.nStep, "t": st.t.hex()}))
dist.barrier(); run.close(); dist.destroy_process_group()
'''


def test_torchrun_two_ranks_ipc(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import oracle
    from util import INNER, assert_bitwise, both_params

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "U.npy"
    env = dict(os.environ, E2D_ROOT=ROOT, E2D_OUT=str(out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    meta = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    _, op = both_params("implode", mesh__nx=256, mesh__ny=128, run__nOutput=-1)
    U_ref, _, n_ref, t_ref = oracle.run(op, 100)
    assert meta["nStep"] == n_ref and float.fromhex(meta["t"]) == t_ref
    assert_bitwise(np.load(out), U_ref[INNER], "implode on 2 ranks over CUDA IPC")
