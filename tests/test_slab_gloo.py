"""Multi-rank host logic on CPU: the y-slab partition, halo exchange, boundary ownership, allreduce(max) of
invDt and device-side stop of euler2d_kokkos_b200.distributed.SlabRun, run with world_size 2 and 3 over the
gloo backend.  The per-slab operators are executed by an oracle-backed engine injected by this test (the
product's CudaEngine needs a GPU); the result must equal the single-domain oracle run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from euler2d_kokkos_b200.distributed import SlabRun, partition_rows, slab_geometry
from euler2d_kokkos_b200 import FACES_X, FACES_YMIN, FACES_YMAX, FACES_ALL
from util import both_params


class OracleEngine:
    """CPU stand-in for CudaEngine with the same contract (writes only interior cells in fused_step)."""

    def __init__(self, params, geo, device):
        from util import both_params  # noqa: F401

        self.geo = geo
        self.op = params._oracle_params

    def init_problem(self, U):
        U.copy_(torch.from_numpy(oracle.init_slab(self.op, self.geo.jsize_loc, self.geo.j_off)))

    def make_boundaries(self, U):
        oracle.make_boundaries(self.op, U.numpy(), bool(self.geo.faces & FACES_YMIN), bool(self.geo.faces & FACES_YMAX))

    def reduce_invdt(self, U, acc):
        acc[0] = max(oracle.compute_invdt(self.op, U.numpy()), 0.0)

    def fused_step(self, Uin, Uout, dt, acc_next, skip):
        acc_next.zero_()
        if int(skip.item()):
            return
        out = oracle.godunov(self.op, Uin.numpy(), float(dt.item()))
        Uout[:, 2:-2, 2:-2] = torch.from_numpy(out[:, 2:-2, 2:-2])
        acc_next[0] = oracle.compute_invdt(self.op, out)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, deck, overrides, steps, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hp, op = both_params(deck, **overrides)
        object.__setattr__(hp, "_oracle_params", op)
        run = SlabRun(hp, device=torch.device("cpu"), engine_factory=OracleEngine)
        run.keep_history = True
        run.run(steps)
        glob = run.gather_interior()
        if rank == 0:
            dts = torch.cat(run.dt_history).numpy()
            np.savez(out_path, U=glob.numpy(), nstep=run.nStep, t=run.time(), dts=dts[~np.isnan(dts)])
        dist.barrier()
    finally:
        dist.destroy_process_group()


CASES = [
    ("implode", dict(mesh__nx=40, mesh__ny=30), 2, 25),                      # reflecting walls
    ("four_quadrant", dict(mesh__nx=36, mesh__ny=35), 3, 25),                # absorbing, uneven split 12/12/11
    ("four_quadrant", dict(mesh__nx=32, mesh__ny=32, mesh__boundary_type_ymin=3, mesh__boundary_type_ymax=3,
                           mesh__boundary_type_xmin=3, mesh__boundary_type_xmax=3), 2, 20),  # periodic wrap, 2 ranks
    ("four_quadrant", dict(mesh__nx=24, mesh__ny=24, run__tEnd=0.02), 2, 200),  # stops on tEnd on the device
]


@pytest.mark.parametrize("deck,overrides,world,steps", CASES)
def test_slab_run_equals_single_domain(tmp_path, deck, overrides, world, steps):
    overrides = dict(overrides, run__nOutput=-1)
    _, op = both_params(deck, **overrides)
    U_ref, dts_ref, n_ref, t_ref = oracle.run(op, steps)
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), deck, overrides, steps, out), nprocs=world, join=True)
    res = np.load(out)
    assert int(res["nstep"]) == n_ref
    assert float(res["t"]) == t_ref
    assert np.array_equal(res["dts"], dts_ref[1:])
    assert np.array_equal(res["U"].view(np.uint64), U_ref[:, 2:-2, 2:-2].view(np.uint64))


def test_partition_and_ownership():
    assert partition_rows(10, 3) == ([4, 3, 3], [0, 4, 7])
    hp, _ = both_params("implode", mesh__nx=16, mesh__ny=16)
    g0, g1, g2 = (slab_geometry(hp, r, 3) for r in range(3))
    assert (g0.faces, g1.faces, g2.faces) == (FACES_X | FACES_YMIN, FACES_X, FACES_X | FACES_YMAX)
    assert (g0.lower, g0.upper, g2.lower, g2.upper) == (None, 1, 1, None)
    assert [g.j_off for g in (g0, g1, g2)] == [0, 6, 11] and [g.ny_loc for g in (g0, g1, g2)] == [6, 5, 5]
    assert slab_geometry(hp, 0, 1).faces == FACES_ALL
    hpp, _ = both_params("implode", mesh__nx=16, mesh__ny=16, mesh__boundary_type_ymin=3, mesh__boundary_type_ymax=3)
    p0, p2 = slab_geometry(hpp, 0, 3), slab_geometry(hpp, 2, 3)
    assert (p0.lower, p0.faces, p2.upper, p2.faces) == (2, FACES_X, 0, FACES_X)
    with pytest.raises(ValueError):
        slab_geometry(hp, 0, 16)


def test_half_periodic_y_is_refused_for_slabs():
    """only one of the two y faces periodic: no rank pair could close the wrap; whole domains may have it"""
    import euler2d_kokkos_b200 as e2d
    from euler2d_kokkos_b200.decks import deck_text
    from euler2d_kokkos_b200.distributed import slab_geometry

    hp = e2d.HydroParams.from_string(deck_text("implode", mesh__nx=32, mesh__ny=32, mesh__boundary_type_ymin=3,
                                               mesh__boundary_type_ymax=1))
    assert slab_geometry(hp, 0, 1).faces == e2d.FACES_ALL
    with pytest.raises(ValueError, match="periodic"):
        slab_geometry(hp, 0, 2)
