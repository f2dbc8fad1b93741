// Internal C++ interface between the C ABI (e2d_capi.cu) and the kernels (e2d_kernels.cu).
#ifndef E2D_INTERNAL_H
#define E2D_INTERNAL_H

#include <cuda_runtime.h>

#include "../../include/euler2d_b200.h"
#include "e2d_math.cuh"

namespace e2d
{

// geometry of one slab as the kernels see it
struct Geom
{
  int isize, jsize; // extent incl. ghosts (jsize = local rows)
  int nx, ny;       // interior extent (ny = local interior rows)
  int j_off;        // global row of local row 0
};

// device-resident time-loop state (src/main.cpp:61-62,100-143 keeps these on the host)
struct LoopState
{
  double             t;
  double             dt;
  unsigned long long invdt_cur;  // bit pattern of max invDt of the current state
  unsigned long long invdt_next; // accumulated by the step kernel for the state it writes
  int                nStep;
  int                done;
};

// ---- the slab loop (e2d_slab.cu): state and peer links, all in device memory ----
constexpr int kMaxRanks = 16;

// written by the peers (and by this rank's own push): one block per rank, cudaMalloc'ed so that it can be shared
// through CUDA IPC
struct SlabComm
{
  unsigned long long invdt_slot[2][kMaxRanks]; // [step parity][rank]: bit pattern of that rank's invDt partial
  unsigned long long invdt_flag[kMaxRanks];    // step+1 of the last partial rank k published here
  unsigned long long halo_flag[2];             // step+1 of the last halo stored by the lower [0] / upper [1] neighbour
  unsigned int       push_blocks_done;         // last-block election of this rank's own push kernel
  unsigned int       fused_cnt[3];             // elections inside the fused step: lower-halo / upper-halo / all blocks
};

// what the fused step needs to publish its halo rows and its invDt partial itself (device pointers; by value)
struct FusedLink
{
  unsigned int *       cnt;                     // -> SlabComm::fused_cnt of this rank
  unsigned int         n_lo, n_hi, n_all;       // blocks holding rows {2,3} / the last two interior rows / all
  unsigned long long * flag_lo;                 // lower neighbour's halo_flag[1], upper neighbour's halo_flag[0]
  unsigned long long * flag_hi;
  SlabComm *           comm[kMaxRanks];
  int                  nranks, rank;
  int                  parity_next;             // slot parity and flag value of the NEXT step
  unsigned long long   seq_next;
};

struct SlabState
{
  double             t;
  double             dt;
  unsigned long long invdt_acc; // accumulated by the fused step (atomicMax on the bit pattern)
  int                nStep;
  int                done;
  int                pending; // dt of an opened step not yet added to t
  int                error;   // a wait for a peer timed out
  // single-GPU loop folded into the step kernel (SoloLoop): rings indexed by (step & 3) — the time at which a step
  // starts and the invDt maximum of the state it reads.  Step n reads slot n, accumulates the next maximum into slot
  // n+1, and its first block writes T[n+1] and clears acc[n+2]: no slot is read and written in the same launch.
  unsigned long long solo_acc[4];
  double             solo_T[4];
};

// Single-GPU device-resident loop with ONE launch per step: the step kernel itself opens the step (every block
// derives dt = cfl / invDt and the tEnd clamp from device memory: src/HydroRun.h:246, src/main.cpp:100,131-134),
// its first block does the bookkeeping of main.cpp:142-143 on the way in (t += dt, nStep++, dt history — all blocks
// derive the same dt, so nothing has to be elected or fenced at the end of the kernel), and every block pushes the
// boundary fill of the rows it has produced into the output array's ghost cells (what make_boundaries would do at
// the start of the next step, e2d_bc.cuh).
struct SoloLoop
{
  SlabState * st = nullptr;
  double      cfl = 0.0, tEnd = 0.0;
  int         max_steps = 0;
  int         step = 0; // nStep of the state this launch reads: ring slot step & 3
  double *    dt_hist = nullptr;
  long        hist_cap = 0;
  int         bc_xmin = 0, bc_xmax = 0, bc_ymin = 0, bc_ymax = 0;
};

struct SlabPushArgs
{
  const double *    A; // the array the coming step reads
  int               isize, jsize;
  double *          lowerA; // the same-parity array of the lower / upper neighbour (peer pointers), or nullptr
  double *          upperA;
  int               lower_jsize, upper_jsize;
  SlabComm *        comm[kMaxRanks]; // every rank's comm block (comm[rank] is local)
  const SlabState * st;
  int               nranks, rank, lower, upper; // neighbour ranks or -1
  int               parity;
  unsigned long long seq; // step + 1
};

struct SlabStepArgs
{
  SlabComm *         mine;
  SlabState *        st;
  unsigned long long seq;
  int                has_lower, has_upper, nranks, parity;
  double             cfl, tEnd;
  int                max_steps;
  double *           dt_hist;
  long               hist_cap;
  long long          timeout_clocks = 120000000000ll; // bound of a wait for a peer, in SM clocks (e2d_run sets it)
};

cudaError_t launch_slab_push(const SlabPushArgs & a, cudaStream_t st);
// pdl: programmatic dependent launch — the kernel may be scheduled while its predecessor in the stream drains; it
// starts with griddepcontrol.wait, so nothing of it runs before the predecessor has completed and flushed
cudaError_t launch_slab_boundaries(const e2d_params & p, const Geom & g, double * A, int faces, const SlabStepArgs & a,
                                   cudaStream_t st, bool pdl = false);
cudaError_t launch_slab_finish(const SlabStepArgs & a, cudaStream_t st);
cudaError_t preload_slab_kernels(); // force-load the loop's kernels (lazy loading may wait for running kernels)
cudaError_t preload_step_kernels();

Settings make_settings(const e2d_params & p);
Geom     make_geom(const e2d_params & p, int jsize_loc, int j_off);

// every launcher returns cudaGetLastError() after the launch
// Sedov (blast with total_energy_inside > 0): n_inside_out != nullptr -> only count the disc cells of the local rows
// [count_jlo, count_jhi) and return (the energy is NOT renormalised yet); n_inside_given >= 0 -> renormalise with that
// global count; neither -> count and renormalise in one call (whole domain).
cudaError_t launch_init_problem(const e2d_params & p, const Geom & g, double * U, cudaStream_t st,
                                unsigned long long * n_inside_out = nullptr, long long n_inside_given = -1,
                                int count_jlo = 0, int count_jhi = 0);
cudaError_t launch_make_boundaries(const e2d_params & p, const Geom & g, double * U, int faces, const int * d_done,
                                   cudaStream_t st);
cudaError_t launch_reduce_invdt(const e2d_params & p, const Geom & g, const double * U, unsigned long long * d_bits,
                                cudaStream_t st);
cudaError_t launch_convert_to_primitives(const e2d_params & p, const Geom & g, const double * U, double * Q,
                                         cudaStream_t st);
cudaError_t launch_compute_and_store_fluxes(const e2d_params & p, const Geom & g, const double * Q, double * Fx,
                                            double * Fy, double dtdx, double dtdy, cudaStream_t st);
cudaError_t launch_update(const e2d_params & p, const Geom & g, double * U, const double * Fx, const double * Fy,
                          cudaStream_t st);
cudaError_t launch_compute_slopes(const e2d_params & p, const Geom & g, const double * Q, double * Sx, double * Sy,
                                  cudaStream_t st);
cudaError_t launch_trace_and_fluxes(const e2d_params & p, const Geom & g, const double * Q, const double * Sx,
                                    const double * Sy, double * F, double dtdx, double dtdy, int dir,
                                    cudaStream_t st);
cudaError_t launch_update_dir(const e2d_params & p, const Geom & g, double * U, const double * F, int dir,
                              cudaStream_t st);
struct MarchPeers
{
  double * lo = nullptr; // same-parity OUTPUT array of the lower / upper neighbour
  double * hi = nullptr;
  int      lo_jsize = 0, hi_jsize = 0;
};
// link/peers: nullptr for a single GPU or when the caller exchanges halos itself
cudaError_t launch_fused_step(const e2d_params & p, const Geom & g, const double * Uin, double * Uout, double dt,
                              const double * d_dt, unsigned long long * d_invdt_bits, const int * d_done,
                              cudaStream_t st, const MarchPeers * peers = nullptr, FusedLink * link = nullptr,
                              int j_first = 2, int j_last = 0 /* <= 0: all rows; else rows [j_first, j_last) */,
                              bool pdl = false /* see launch_slab_boundaries */,
                              const SoloLoop * solo = nullptr /* single-GPU loop: dt, invDt, done come from solo->st */);
int         device_sm_count(); // multiprocessors of the current device (cached per device)
// refined reciprocal of the strict division sequence for denominator d (device-evaluated once per value, cached;
// synchronises on a miss — e2d_create warms it for dx, dy so that no launch inside a loop ever misses)
double      refined_reciprocal(double d);

// Programmatic dependent launch (sm_90+).  Both are no-ops in a kernel launched without the attribute.
#if defined(__CUDACC__)
__device__ __forceinline__ void
pdl_wait_for_predecessor()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void
pdl_release_successor()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
// out's two-cell ghost frame <- in's (what the reference's deep_copy(out, in) leaves there, src/HydroRun.h:302)
cudaError_t launch_copy_ghost_frame(const Geom & g, const double * in, double * out, cudaStream_t st);
// x-ghost columns of rows [jlo, jhi) (faces & E2D_FACES_X)
cudaError_t launch_bc_x_rows(const e2d_params & p, const Geom & g, double * U, int faces, int jlo, int jhi,
                             cudaStream_t st);
// scalar bookkeeping of the device-resident loop
cudaError_t launch_loop_begin_step(LoopState * st_dev, double cfl, double tEnd, cudaStream_t st);
cudaError_t launch_loop_end_step(LoopState * st_dev, double tEnd, int max_steps, double * dt_hist, long hist_cap,
                                 cudaStream_t st);
cudaError_t launch_eval(const e2d_params & p, int func, const double * d_in, double * d_out, long n,
                        cudaStream_t st);

// ---- post-processing / output (e2d_post.cu) ----
struct RadialArgs
{
  double xmin, ymin, dx, dy, cx, cy, rmax;
  int    nbins, gw;
  int    j_lo, j_hi; // local rows [j_lo, j_hi) of the slab are binned
};
RadialArgs  make_radial_args(const e2d_params & p, int nbins, int j_lo, int j_hi);
int         radial_segments(const Geom & g, int nbins, int rows);
// part_sum / part_cnt: nbins * nseg * isize entries each (scratch); d_sums / d_counts: nbins entries
cudaError_t launch_radial_profile(const Geom & g, const RadialArgs & a, const double * U, int nseg, double * part_sum,
                                  int * part_cnt, double * d_sums, int * d_counts, cudaStream_t st);
// interior cells of rows [j_lo, j_lo + n_rows), ghost columns stripped: out[var][row][nx]
cudaError_t launch_gather_interior(const Geom & g, const double * U, double * out, int j_lo, int n_rows,
                                   cudaStream_t st);

int  solver_for(const e2d_params & p); // 2 (HLLC) unless honourRiemannSolver
void count_launch(int n = 1);

} // namespace e2d

#endif
