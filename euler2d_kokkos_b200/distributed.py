"""y-slab multi-GPU driver: one process per GPU (torchrun), torch.distributed for the plumbing.

The global grid is cut into ``world`` horizontal slabs (SoA planes keep each slab's rows contiguous).
One time step on every rank (bit-identical to the single-GPU run):

  1. halo exchange   first/last two interior rows of the current array -> the neighbours' ghost rows
                     (periodic y: rank 0 <-> rank world-1 as well)
  2. boundary fill   x faces on all local rows (incl. the received halo rows), physical y faces on the
                     ranks that own them — one launch (e2d_k_make_boundaries with a face mask)
  3. dt              allreduce(MAX) of the per-rank invDt that the previous step's fused epilogue left in
                     device memory; dt = cfl / invDt and the tEnd clamp are evaluated on the device
  4. fused step      e2d_k_fused_step(in -> out), leaving the next invDt partial in device memory

Nothing in the loop synchronises with the host.  The engine that executes 2 and 4 is injected:
``CudaEngine`` (the product: libeuler2d_b200.so on this rank's GPU) — the CPU test-suite injects an
oracle-backed engine to exercise the partition/exchange/ordering logic with the gloo backend.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _lib
from ._lib import FACES_X, FACES_YMAX, FACES_YMIN, BC_PERIODIC, Slab, check, lib
from .hydro_run import HydroParams


def partition_rows(ny: int, world: int):
    """Contiguous split of ny interior rows: the first ny % world ranks get one extra row."""
    base, rem = divmod(ny, world)
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    starts = [sum(counts[:r]) for r in range(world)]
    return counts, starts


@dataclass
class SlabGeometry:
    rank: int
    world: int
    ny_loc: int
    j_off: int  # global row (ghosts included) of local row 0 == first interior row offset
    jsize_loc: int
    faces: int  # boundary faces this rank fills itself
    lower: int | None  # neighbour ranks (None: physical boundary)
    upper: int | None


def slab_geometry(params: HydroParams, rank: int, world: int) -> SlabGeometry:
    counts, starts = partition_rows(params.ny, world)
    if min(counts) < 2:
        raise ValueError(f"ny={params.ny} is too small for {world} slabs (need >= 2 interior rows per slab)")
    if world > 1 and (params.boundary_type_ymin == BC_PERIODIC) != (params.boundary_type_ymax == BC_PERIODIC):
        # the wrap between the first and the last rank needs both y faces periodic (e2d_create refuses it as well)
        raise ValueError("y-slabs need boundary_type_ymin and boundary_type_ymax both periodic or neither")
    faces = FACES_X
    lower = rank - 1 if rank > 0 else None
    upper = rank + 1 if rank < world - 1 else None
    if world == 1:
        faces = _lib.FACES_ALL
    else:
        if rank == 0:
            if params.boundary_type_ymin == BC_PERIODIC:
                lower = world - 1
            else:
                faces |= FACES_YMIN
        if rank == world - 1:
            if params.boundary_type_ymax == BC_PERIODIC:
                upper = 0
            else:
                faces |= FACES_YMAX
    return SlabGeometry(rank, world, counts[rank], starts[rank], counts[rank] + 4, faces, lower, upper)


class CudaEngine:
    """Executes the per-slab operators on this rank's GPU through the C ABI (kernel-level entry points)."""

    def __init__(self, params: HydroParams, geo: SlabGeometry, device: torch.device):
        if lib().e2d_device_count() < 1:
            raise _lib.E2dError("no CUDA device: euler2d_kokkos_b200 has no CPU fallback")
        self.params, self.geo, self.device = params, geo, device
        self._p = C.byref(params.raw)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def init_problem(self, U: torch.Tensor):
        check(lib().e2d_k_init_problem(self._p, C.c_void_p(U.data_ptr()), self.geo.jsize_loc, self.geo.j_off,
                                       self._stream()), "e2d_k_init_problem")

    def make_boundaries(self, U: torch.Tensor):
        check(lib().e2d_k_make_boundaries(self._p, C.c_void_p(U.data_ptr()), self.geo.jsize_loc, self.geo.faces,
                                          self._stream()), "e2d_k_make_boundaries")

    def reduce_invdt(self, U: torch.Tensor, acc: torch.Tensor):
        acc.zero_()
        check(lib().e2d_k_reduce_invdt(self._p, C.c_void_p(U.data_ptr()), self.geo.jsize_loc,
                                       C.c_void_p(acc.data_ptr()), self._stream()), "e2d_k_reduce_invdt")

    def fused_step(self, Uin: torch.Tensor, Uout: torch.Tensor, dt: torch.Tensor, acc_next: torch.Tensor,
                   skip: torch.Tensor):
        acc_next.zero_()
        check(lib().e2d_k_fused_step(self._p, C.c_void_p(Uin.data_ptr()), C.c_void_p(Uout.data_ptr()),
                                     self.geo.jsize_loc, 0.0, C.c_void_p(dt.data_ptr()),
                                     C.c_void_p(acc_next.data_ptr()), C.c_void_p(skip.data_ptr()),
                                     self._stream()), "e2d_k_fused_step")


class PeerSlabRun:
    """HydroRun for one y-slab of a multi-process run, stepped by the library's own device-resident loop
    (csrc/e2d_slab.cu): halo rows and CFL partials travel as direct stores into the neighbours' memory over NVLink
    (CUDA IPC peer pointers) with system-scope flags — no NCCL call and no host round trip per step.
    ``torch.distributed`` is used once, to all-gather the IPC handles."""

    def __init__(self, params: HydroParams, rank: int | None = None, world: int | None = None,
                 device: torch.device | None = None, group=None):
        from .hydro_run import HydroRun

        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.params = params
        self.geo = slab_geometry(params, self.rank, self.world)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        slab = Slab(self.rank, self.world, self.geo.ny_loc, self.geo.j_off)
        self.hydro = HydroRun(params, slab=slab)
        if self.world > 1:
            # Sedov init: the disc-cell counts of the slabs are summed (integer all-reduce) and every rank completes
            # its initialisation with the global count (e2d_blast_renormalise; a no-op for every other problem)
            n_loc, pending = C.c_ulonglong(), C.c_int()
            check(lib().e2d_blast_inside_count(self.hydro._h, C.byref(n_loc), C.byref(pending)), "e2d_blast_inside_count")
            cnt = torch.tensor([n_loc.value, pending.value], dtype=torch.int64, device=self.device)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=self.group)
            if int(cnt[1].item()) > 0:
                check(lib().e2d_blast_renormalise(self.hydro._h, int(cnt[0].item())), "e2d_blast_renormalise")
            blob = _lib.IpcBlob()
            check(lib().e2d_ipc_export(self.hydro._h, C.byref(blob)), "e2d_ipc_export")
            mine = torch.frombuffer(bytearray(bytes(blob)), dtype=torch.uint8).to(self.device)
            allb = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allb, mine, group=self.group)
            blobs = (_lib.IpcBlob * self.world)()
            for r, tb in enumerate(allb):
                raw = bytes(tb.cpu().numpy().tobytes())
                C.memmove(C.byref(blobs[r]), raw, C.sizeof(_lib.IpcBlob))
            check(lib().e2d_ipc_connect(self.hydro._h, blobs, self.world), "e2d_ipc_connect")
            dist.barrier(group=self.group)  # every rank has mapped its peers before anyone stores into them

    def run(self, max_steps: int):
        """Steps until nStep == max_steps or t >= tEnd (every rank must pass the same max_steps).
        The ranks rendezvous first: inside the loop a kernel waits for its neighbours' halo rows only for a bounded time
        (E2D_PEER_TIMEOUT_S), so nobody may enter it while a peer is still busy elsewhere (I/O on rank 0, a gather)."""
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        return self.hydro.run(max_steps)

    def current(self, nStep: int) -> torch.Tensor:
        """This rank's slab [4][ny_loc+4][isize] after nStep steps, as a host tensor."""
        import numpy as np

        return torch.from_numpy(np.ascontiguousarray(self.hydro.download(nStep % 2)))

    def gather_interior(self, nStep: int) -> torch.Tensor | None:
        """Rank 0 receives the global interior [4][ny][nx] (tests / output)."""
        mine = self.current(nStep)[:, 2:-2, 2:-2].contiguous().to(self.device)
        if self.world == 1:
            return mine
        counts, _ = partition_rows(self.params.ny, self.world)
        if self.rank == 0:
            parts = [mine]
            for r in range(1, self.world):
                buf = torch.empty((4, counts[r], self.params.nx), dtype=torch.float64, device=self.device)
                dist.recv(buf, r, self.group)
                parts.append(buf)
            return torch.cat(parts, dim=1)
        dist.send(mine, 0, self.group)
        return None

    def close(self):
        self.hydro.close()


class SlabRun:
    """HydroRun for one y-slab of a multi-process run (same driver surface: compute_dt / make_boundaries /
    godunov_unsplit semantics folded into ``step``/``run``)."""

    def __init__(self, params: HydroParams, rank: int | None = None, world: int | None = None,
                 device: torch.device | None = None, engine_factory=CudaEngine, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.params = params
        self.geo = slab_geometry(params, self.rank, self.world)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        shape = (4, self.geo.jsize_loc, params.isize)
        self.U = torch.zeros(shape, dtype=torch.float64, device=self.device)
        self.U2 = torch.zeros(shape, dtype=torch.float64, device=self.device)
        self.engine = engine_factory(params, self.geo, self.device)
        self.invdt = [torch.zeros(1, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.dt = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.t = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.tEnd = torch.full((1,), params.tEnd, dtype=torch.float64, device=self.device)
        self.cfl = torch.full((1,), params.cfl, dtype=torch.float64, device=self.device)
        self._not_active = torch.zeros(1, dtype=torch.bool, device=self.device)
        self.nsteps_issued = 0                                   # host-side count of step() calls
        self.nstep_dev = torch.zeros(1, dtype=torch.int64, device=self.device)   # steps actually taken
        self.skip = torch.zeros(1, dtype=torch.int32, device=self.device)        # 1 once t >= tEnd
        self._send = [torch.empty((4, 2, params.isize), dtype=torch.float64, device=self.device) for _ in range(2)]
        self._recv = [torch.empty((4, 2, params.isize), dtype=torch.float64, device=self.device) for _ in range(2)]
        self.dt_history: list[torch.Tensor] = []
        self.keep_history = False
        # HydroRun::HydroRun: initial condition, U2 = U (src/HydroRun.h:185-214)
        self.engine.init_problem(self.U)
        self.U2.copy_(self.U)
        # main.cpp:87: the CFL reduction of the initial state primes the loop
        self.engine.reduce_invdt(self.U, self.invdt[0])

    # ------------------------------------------------------------------ halo exchange
    def exchange_halos(self, A: torch.Tensor):
        """Neighbour ghost rows <- my first / last two interior rows (all 4 variables, full width)."""
        g = self.geo
        if self.world == 1:
            return
        # Order matters for NCCL (which ignores tags and pairs sends with receives in issue order): every
        # rank sends [upper rows, lower rows] and receives [lower halo, upper halo], which also pairs up
        # correctly when a periodic wrap makes one peer both the lower and the upper neighbour.
        ops = []
        if g.upper is not None:
            self._send[1].copy_(A[:, -4:-2, :])
            ops.append(dist.P2POp(dist.isend, self._send[1], g.upper, self.group, tag=2))
        if g.lower is not None:
            self._send[0].copy_(A[:, 2:4, :])
            ops.append(dist.P2POp(dist.isend, self._send[0], g.lower, self.group, tag=1))
        if g.lower is not None:
            ops.append(dist.P2POp(dist.irecv, self._recv[0], g.lower, self.group, tag=2))
        if g.upper is not None:
            ops.append(dist.P2POp(dist.irecv, self._recv[1], g.upper, self.group, tag=1))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if g.lower is not None:
            A[:, 0:2, :].copy_(self._recv[0])
        if g.upper is not None:
            A[:, -2:, :].copy_(self._recv[1])

    # ------------------------------------------------------------------ one step / loop
    def step(self):
        """compute_dt + godunov_unsplit of src/main.cpp:128-143 for this slab.  No host synchronisation:
        the loop condition t < tEnd (main.cpp:100) is evaluated on the device and turns the step into a no-op
        once it fails, so the host may issue steps past the end."""
        n = self.nsteps_issued
        A, B = (self.U, self.U2) if n % 2 == 0 else (self.U2, self.U)
        inv = self.invdt[n % 2]
        if self.world > 1:
            dist.all_reduce(inv, op=dist.ReduceOp.MAX, group=self.group)
        active = self.t < self.tEnd
        torch.logical_not(active, out=self._not_active)
        self.skip.copy_(self._not_active)
        # dt = cfl / invDt (HydroRun.h:246); if (t + dt > tEnd) dt = tEnd - t (main.cpp:131-134)
        torch.div(self.cfl, inv, out=self.dt)
        torch.where(self.t + self.dt > self.tEnd, self.tEnd - self.t, self.dt, out=self.dt)
        self.exchange_halos(A)
        self.engine.make_boundaries(A)
        self.engine.fused_step(A, B, self.dt, self.invdt[(n + 1) % 2], self.skip)
        torch.where(active, self.t + self.dt, self.t, out=self.t)
        self.nstep_dev += active
        self.nsteps_issued = n + 1
        if self.keep_history:
            self.dt_history.append(torch.where(active, self.dt, torch.full_like(self.dt, float("nan"))))

    def run(self, max_steps: int):
        """Loop of src/main.cpp:100: issues steps until nStep == max_steps; steps issued after t reached tEnd
        are device-side no-ops.  The host looks at the device once every 32 steps to stop early."""
        while self.nsteps_issued < max_steps:
            if self.nsteps_issued % 32 == 31 and bool(self.skip.item()):
                break
            self.step()
        return self.nStep

    @property
    def nStep(self) -> int:
        """Steps actually taken (synchronises)."""
        return int(self.nstep_dev.item())

    def time(self) -> float:
        return float(self.t.item())

    def current(self) -> torch.Tensor:
        return self.U if self.nStep % 2 == 0 else self.U2

    def gather_interior(self) -> torch.Tensor | None:
        """Rank 0 receives the global interior [4][ny][nx] (tests / output)."""
        mine = self.current()[:, 2:-2, 2:-2].contiguous()
        if self.world == 1:
            return mine
        counts, _ = partition_rows(self.params.ny, self.world)
        if self.rank == 0:
            parts = [mine]
            for r in range(1, self.world):
                buf = torch.empty((4, counts[r], self.params.nx), dtype=torch.float64, device=self.device)
                dist.recv(buf, r, self.group)
                parts.append(buf)
            return torch.cat(parts, dim=1)
        dist.send(mine, 0, self.group)
        return None
