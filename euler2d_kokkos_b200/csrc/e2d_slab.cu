// The device-resident time loop, single- and multi-GPU (y-slabs), with NO host round trip and NO collective
// library call per step: dt, t and nStep live in device memory; the halo rows and the per-rank CFL partials travel
// as plain stores into the neighbours' memory over NVLink (peer pointers: cudaDeviceEnablePeerAccess inside one
// process, CUDA IPC between the processes of a torchrun job) and are published with system-scope flags.
//
// One step n on every rank (A = the array the step reads, by the parity of n; all ranks are at the same n):
//
//   k_slab_push        (multi-GPU, FIRST step of an e2d_run call only)  my first / last two interior rows of A  ->
//                      the neighbours' ghost rows of A; my invDt partial  ->  slot [n&1][rank] of EVERY rank; then,
//                      by the last block to finish and after a system fence, the flags halo_flag / invdt_flag := n+1
//                      in the receivers' memory.  For every later step the PREVIOUS fused step has already done all
//                      of this itself: its edge segments are scheduled first and store their rows into the
//                      neighbours' ghost rows as they produce them, and its last block publishes the invDt partial
//                      (k_fused_step<.., LINKED>, e2d_kernels.cu) — the exchange overlaps the interior compute.
//   k_slab_boundaries  waits (spinning on its OWN memory) for the neighbours' halo flags and — one thread — for all
//                      invDt flags; fills the boundaries of A (x faces on all local rows incl. the received halo
//                      rows, physical y faces where this rank owns them: SURVEY.md §8e order, bit-exact corners);
//                      the one thread closes step n-1 (t += dt, nStep++, dt history) and opens step n:
//                      dt = cfl / max_k invDt_k  (max is exact => identical on every rank and to a single-GPU run),
//                      the tEnd clamp and the loop condition of src/main.cpp:100,131-134.
//   k_fused_step       reads dt / done from device memory, leaves the next invDt partial in device memory.
//
// Why this is race free: flags carry the step number and only ever grow.  Halo data is double-buffered by the
// ping-pong arrays: a neighbour can store into my ghost rows of A for step n+2 only after its fused step n+1, which
// waited for my push of step n+1, which is stream-ordered after my fused step n — the last reader of A's ghosts.
// invDt slots are double-buffered by step parity: a rank publishes step n+2 only after its fused step n+1, which
// needed every rank's step-n+1 partial, which each rank publishes (stream order) after it consumed the step-n slots.
// No kernel ever waits for a kernel that waits for it: pushes wait for nothing.
#include <cstdio>

#include "e2d_bc.cuh"
#include "e2d_internal.h"

namespace e2d
{

namespace
{

__device__ __forceinline__ unsigned long long
ld_acquire_sys(const unsigned long long * p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Spin until *flag >= seq.  Bounded (SlabStepArgs::timeout_clocks, 60 s of SM clocks unless E2D_PEER_TIMEOUT_S says
// otherwise): if a peer died, raise the error flag AND `done`, so that every later step of the batch is a no-op
// instead of marching on halo rows that never arrived; the host reports the error and marks the handle invalid.
__device__ __forceinline__ void
wait_flag(const unsigned long long * flag, unsigned long long seq, SlabState * st, long long timeout_clocks)
{
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq)
    if (clock64() - t0 > timeout_clocks)
    {
      st->error = 1;
      st->done = 1;
      break;
    }
}

__device__ __forceinline__ void
st_release_sys(unsigned long long * p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_slab_push(SlabPushArgs a)
{
  const int  isize = a.isize;
  const long per_dir = 8L * isize; // 4 variables x 2 rows x isize
  const long total = 2 * per_dir;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long)gridDim.x * blockDim.x)
  {
    const int  dir = k >= per_dir;     // 0: to the lower neighbour, 1: to the upper neighbour
    const long kk = k - dir * per_dir; // (v, row, i), i fastest
    const int  i = (int)(kk % isize);
    const int  row = (int)((kk / isize) & 1);
    const int  v = (int)(kk / (2L * isize));
    double *   dst = dir ? a.upperA : a.lowerA;
    if (!dst)
      continue;
    // lower neighbour: my rows 2,3 -> its top ghost rows; upper neighbour: my last interior rows -> its rows 0,1
    const int    js = dir ? a.jsize - 4 + row : 2 + row;
    const int    jd = dir ? row : a.lower_jsize - 2 + row;
    const int    jsize_d = dir ? a.upper_jsize : a.lower_jsize;
    const size_t so = (size_t)i + (size_t)isize * ((size_t)js + (size_t)a.jsize * v);
    const size_t d_o = (size_t)i + (size_t)isize * ((size_t)jd + (size_t)jsize_d * v);
    dst[d_o] = a.A[so];
  }
  if (blockIdx.x == 0 && threadIdx.x < a.nranks)
    a.comm[threadIdx.x]->invdt_slot[a.parity][a.rank] = a.st->invdt_acc;

  // publish: every block fences its stores, the last one to arrive raises the flags in the receivers' memory
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    SlabComm *         mine = a.comm[a.rank];
    const unsigned int prev = atomicAdd(&mine->push_blocks_done, 1u);
    if (prev == gridDim.x - 1)
    {
      mine->push_blocks_done = 0;
      __threadfence_system();
      if (a.lower >= 0)
        st_release_sys(&a.comm[a.lower]->halo_flag[1], a.seq); // I am its upper neighbour
      if (a.upper >= 0)
        st_release_sys(&a.comm[a.upper]->halo_flag[0], a.seq); // I am its lower neighbour
      for (int k = 0; k < a.nranks; ++k)
        st_release_sys(&a.comm[k]->invdt_flag[a.rank], a.seq);
    }
  }
}

// closes the previous step and opens the next one (src/main.cpp:100-143; HydroRun.h:246)
__device__ __forceinline__ void
loop_scalars(const SlabStepArgs & a, bool open_next)
{
  SlabState s = *a.st;
  if (s.pending)
  {
    if (a.dt_hist && s.nStep < a.hist_cap)
      a.dt_hist[s.nStep] = s.dt;
    s.t += s.dt; // main.cpp:142-143
    s.nStep += 1;
    s.pending = 0;
  }
  s.done = !(s.t < a.tEnd && s.nStep < a.max_steps) || (*(volatile int *)&a.st->error != 0); // main.cpp:100
  if (open_next && !s.done)
  {
    unsigned long long bits = s.invdt_acc;
    if (a.nranks > 1)
    {
      bits = 0ull; // invDt >= 0: its IEEE bit pattern orders like an unsigned integer
      for (int k = 0; k < a.nranks; ++k)
      {
        const unsigned long long b = ld_acquire_sys(&a.mine->invdt_slot[a.parity][k]); // stored by rank k
        bits = b > bits ? b : bits;
      }
    }
    const double invDt = __longlong_as_double((long long)bits);
    double       dt = a.cfl / invDt; // HydroRun.h:246
    if (s.t + dt > a.tEnd)           // main.cpp:131-134
      dt = a.tEnd - s.t;
    s.dt = dt;
    s.pending = 1;
  }
  if (open_next)
    s.invdt_acc = 0ull; // consumed; the fused step accumulates the next partial from zero
  // field-wise write-back: `error` may be raised concurrently by a waiting block
  a.st->t = s.t;
  a.st->dt = s.dt;
  a.st->invdt_acc = s.invdt_acc;
  a.st->nStep = s.nStep;
  a.st->done = s.done;
  a.st->pending = s.pending;
}

__global__ void __launch_bounds__(128)
k_slab_boundaries(Geom g, BcArgs bc, double * __restrict__ A, SlabStepArgs a)
{
  pdl_wait_for_predecessor(); // the previous fused step has completed (its state, its invDt partial)
  pdl_release_successor();    // the fused step behind this kernel may be scheduled; it waits for our completion
  if (threadIdx.x == 0)
  {
    // once the loop is over the fused steps are no-ops and publish nothing: there is nothing to wait for
    const bool over = *(volatile int *)&a.st->done != 0;
    if (a.has_lower && !over)
      wait_flag(&a.mine->halo_flag[0], a.seq, a.st, a.timeout_clocks);
    if (a.has_upper && !over)
      wait_flag(&a.mine->halo_flag[1], a.seq, a.st, a.timeout_clocks);
    if (blockIdx.x == 0)
    {
      if (a.nranks > 1 && !over)
        for (int k = 0; k < a.nranks; ++k)
          wait_flag(&a.mine->invdt_flag[k], a.seq, a.st, a.timeout_clocks);
      loop_scalars(a, true);
    }
  }
  __syncthreads();
  bc_fill_cell(g, bc, A, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void
k_slab_finish(SlabStepArgs a)
{
  loop_scalars(a, false);
}

} // namespace

// CUDA loads kernels lazily, and loading may wait for running kernels: a rank whose boundary kernel is already
// spinning on the device must never be the reason a peer's first launch cannot load.  Only matters when several
// ranks share one device (tests), but costs nothing: resolve every kernel of the loop before the first step.
cudaError_t
preload_slab_kernels()
{
  cudaFuncAttributes fa;
  cudaError_t        e = cudaFuncGetAttributes(&fa, k_slab_push);
  if (e == cudaSuccess)
    e = cudaFuncGetAttributes(&fa, k_slab_boundaries);
  if (e == cudaSuccess)
    e = cudaFuncGetAttributes(&fa, k_slab_finish);
  return e;
}

cudaError_t
launch_slab_push(const SlabPushArgs & a, cudaStream_t st)
{
  const long total = 16L * a.isize;
  int        blocks = (int)((total + 255) / 256);
  if (blocks > 148)
    blocks = 148; // all resident at once; grid-stride covers the rest
  k_slab_push<<<blocks, 256, 0, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_slab_boundaries(const e2d_params & p, const Geom & g, double * A, int faces, const SlabStepArgs & a,
                       cudaStream_t st, bool pdl)
{
  const int          n = 4 * g.isize + 4 * g.jsize;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((n + 127) / 128));
  cfg.blockDim = dim3(128);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k_slab_boundaries, g, make_bc_args(p, faces), A, a);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t
launch_slab_finish(const SlabStepArgs & a, cudaStream_t st)
{
  k_slab_finish<<<1, 1, 0, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

} // namespace e2d
