// sm_100a kernels of the unsplit MUSCL-Hancock Godunov step.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (strict: bit parity with the
// reference's FMA-free x86 build).  Data layout: SoA planes, off = i + isize*(j + jsize*var);
// threadIdx.x always walks i, so every global access of a warp is a contiguous 256-byte run.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <utility>
#include <vector>

#include "e2d_bc.cuh"
#include "e2d_internal.h"
#include "e2d_march.cuh" // includes e2d_lean.cuh

namespace e2d
{

namespace
{

#ifndef E2D_BX
#  define E2D_BX 128
#endif
#ifndef E2D_STRICT_MIN_BLOCKS
#  define E2D_STRICT_MIN_BLOCKS (384 / E2D_BX)
#endif
constexpr int kBX = E2D_BX;    // threads per block of the marching kernel (= columns incl. 4 halo columns)
constexpr int kMarchMinBlocks = E2D_STRICT_MIN_BLOCKS; // 384 threads per SM (<= 168 registers per thread)
#ifndef E2D_FAST_MIN_BLOCKS
#  define E2D_FAST_MIN_BLOCKS 4
#endif
constexpr int kMarchMinBlocksFast = E2D_FAST_MIN_BLOCKS; // same for the `arithmetic=fast` instantiations
#ifndef E2D_MARCH_UNROLL
#  define E2D_MARCH_UNROLL 2
#endif
// rows per trip of the marching loop of the fast instantiations (ring slot r & 1 becomes a constant, fewer carried
// register moves: +2 %; the strict kernel, at 158 registers, loses 4 % to the same unrolling)
constexpr int kMarchUnrollFast = E2D_MARCH_UNROLL;
#ifndef E2D_STRICT_UNROLL
#  define E2D_STRICT_UNROLL 1
#endif
constexpr int kMarchUnrollStrict = E2D_STRICT_UNROLL;
constexpr int
march_min_blocks(int math)
{
  return math == 1 ? kMarchMinBlocksFast : kMarchMinBlocks;
}

// row handled by a block of a grid_rows() launch: gridDim.y is capped at 32768 rows, gridDim.z carries the rest
__device__ __forceinline__ int
grid_row()
{
  return (int)(blockIdx.y + blockIdx.z * gridDim.y);
}

__host__ __device__ __forceinline__ size_t
cell(const Geom & g, int i, int j, int v)
{
  return (size_t)i + (size_t)g.isize * ((size_t)j + (size_t)g.jsize * (size_t)v);
}

__device__ __forceinline__ void
load4(const double * __restrict__ A, const Geom & g, int i, int j, double q[4])
{
  const size_t plane = (size_t)g.isize * g.jsize;
  const size_t o = (size_t)i + (size_t)g.isize * j;
#pragma unroll
  for (int v = 0; v < 4; ++v)
    q[v] = A[o + v * plane];
}

__device__ __forceinline__ void
store4(double * __restrict__ A, const Geom & g, int i, int j, const double q[4])
{
  const size_t plane = (size_t)g.isize * g.jsize;
  const size_t o = (size_t)i + (size_t)g.isize * j;
#pragma unroll
  for (int v = 0; v < 4; ++v)
    A[o + v * plane] = q[v];
}

// ------------------------------------------------------------------------------------------
// Problem initialisers: Init*Functor, src/HydroRunFunctors.h:1347-1827.
// Cell centre  x = xmin + dx/2 + (i - ghostWidth)*dx  (:1384-1385) with the GLOBAL row index.
// ------------------------------------------------------------------------------------------
struct InitArgs
{
  int    problem;
  double xmin, ymin, dx, dy, gamma0;
  double blast_radius, blast_center_x, blast_center_y, blast_density_in, blast_density_out;
  double blast_pressure_in, blast_pressure_out;
  double blast_energy_density; // total_energy_inside / volume_inside, or < 0 when not used
  double bubble_radius, bubble_center_x, bubble_center_y, bubble_density, bubble_pressure;
  double preshock_density, preshock_pressure, postshock_density, postshock_pressure, postshock_velocity;
  double shock_loc;
};

__global__ void __launch_bounds__(128)
k_init_problem(Geom g, InitArgs a, double * __restrict__ U, unsigned long long * __restrict__ n_inside, int count_jlo,
               int count_jhi)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = grid_row();
  if (i >= g.isize || j >= g.jsize)
    return;
  const int    gw = 2;
  const double x = a.xmin + a.dx / 2 + (i - gw) * a.dx;
  const double y = a.ymin + a.dy / 2 + (j + g.j_off - gw) * a.dy;
  double       u[4] = { 0.0, 0.0, 0.0, 0.0 };

  if (a.problem == E2D_PROBLEM_BLAST)
  { // :1466-1545
    const double radius2 = a.blast_radius * a.blast_radius;
    const double d2 = (x - a.blast_center_x) * (x - a.blast_center_x) +
                      (y - a.blast_center_y) * (y - a.blast_center_y);
    if (d2 < radius2)
    {
      u[ID] = a.blast_density_in;
      u[IP] = a.blast_pressure_in / (a.gamma0 - 1.0);
      if (n_inside && j >= count_jlo && j < count_jhi) // a slab counts the rows it owns (no halo row twice)
        atomicAdd(n_inside, 1ull);
      if (a.blast_energy_density >= 0.0)
        u[IP] = a.blast_energy_density;
    }
    else
    {
      u[ID] = a.blast_density_out;
      u[IP] = a.blast_pressure_out / (a.gamma0 - 1.0);
    }
  }
  else if (a.problem == E2D_PROBLEM_FOUR_QUADRANT)
  { // :1578-1665, Lax-Liu configuration 3 split at (0.8, 0.8)
    const double xt = 0.8, yt = 0.8;
    double       rho, p, vx, vy;
    if (x < xt)
    {
      if (y < yt)
      {
        rho = 0.138, p = 0.029, vx = 1.206, vy = 1.206;
      }
      else
      {
        rho = 0.5323, p = 0.3, vx = 1.206, vy = 0.0;
      }
    }
    else
    {
      if (y < yt)
      {
        rho = 0.5323, p = 0.3, vx = 0.0, vy = 1.206;
      }
      else
      {
        rho = 1.5, p = 1.5, vx = 0.0, vy = 0.0;
      }
    }
    u[ID] = rho; // primToCons :1578-1591
    u[IU] = vx * rho;
    u[IV] = vy * rho;
    u[IP] = p / (a.gamma0 - 1.0) + rho * (vx * vx + vy * vy) * 0.5;
  }
  else if (a.problem == E2D_PROBLEM_DISCONTINUITY)
  { // :1713-1729
    u[ID] = (x + y < 1) ? 1.0 + x * x : 0.25;
    u[IP] = 1.0 / (a.gamma0 - 1.0);
  }
  else if (a.problem == E2D_PROBLEM_SHOCKED_BUBBLE)
  { // :1776-1820
    double pres;
    if (x < a.shock_loc)
    {
      u[ID] = a.postshock_density;
      u[IU] = a.postshock_density * a.postshock_velocity;
      pres = a.postshock_pressure;
    }
    else
    {
      const double radius = sqrt((x - a.bubble_center_x) * (x - a.bubble_center_x) +
                                 (y - a.bubble_center_y) * (y - a.bubble_center_y));
      if (radius < a.bubble_radius)
      {
        u[ID] = a.bubble_density;
        pres = a.bubble_pressure;
      }
      else
      {
        u[ID] = a.preshock_density;
        pres = a.preshock_pressure;
      }
    }
    const double rho_eint = pres / (a.gamma0 - 1);
    u[IE] = rho_eint + 0.5 * (u[IU] * u[IU] + u[IV] * u[IV]) / u[ID];
  }
  else
  { // implode :1384-1401 (also the reference's fallback for an unknown problem, HydroRun.h:205-211)
    const double tmp = x + y * y;
    if (tmp > 0.5 && tmp < 1.5)
    {
      u[ID] = 1.0;
      u[IP] = 1.0 / (a.gamma0 - 1.0);
    }
    else
    {
      u[ID] = 0.125;
      u[IP] = 0.14 / (a.gamma0 - 1.0);
    }
  }
  store4(U, g, i, j, u);
}

// ------------------------------------------------------------------------------------------
// Boundary fill (e2d_bc.cuh): the four MakeBoundariesFunctor<face> launches as ONE launch
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_make_boundaries(Geom g, BcArgs a, double * __restrict__ U, const int * __restrict__ d_done)
{
  if (d_done && *d_done)
    return;
  bc_fill_cell(g, a, U, blockIdx.x * blockDim.x + threadIdx.x);
}

// x-ghost columns of rows [jlo, jhi) only (a.faces = X faces): the host-streamed step fills each chunk of rows as
// it lands; the y faces follow once their source rows are there (same XMIN, XMAX -> YMIN, YMAX order as
// HydroRun::make_boundaries, src/HydroRun.h:390-399)
__global__ void __launch_bounds__(128)
k_bc_x_rows(Geom g, BcArgs a, double * __restrict__ U, int jlo, int jhi)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < 4 * (jhi - jlo))
    bc_fill_cell(g, a, U, 4 * g.isize + 4 * jlo + k);
}

// The two-cell frame of ghost cells, copied from `in` to `out`: the part of the reference's deep_copy(data_out,
// data_in) (src/HydroRun.h:302) that survives the update, which only rewrites the interior.
__global__ void __launch_bounds__(128)
k_copy_ghost_frame(Geom g, const double * __restrict__ in, double * __restrict__ out)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  int       i, j;
  if (k < 4 * g.isize)
  { // rows 0, 1, jsize-2, jsize-1, full width
    const int r = k / g.isize;
    i = k - r * g.isize;
    j = r < 2 ? r : g.jsize - 4 + r;
  }
  else
  { // columns 0, 1, isize-2, isize-1 of the rows in between
    const int kk = k - 4 * g.isize;
    if (kk >= 4 * (g.jsize - 4))
      return;
    const int c = kk & 3;
    j = 2 + (kk >> 2);
    i = c < 2 ? c : g.isize - 4 + c;
  }
  const size_t plane = (size_t)g.isize * g.jsize;
  const size_t o = (size_t)j * g.isize + i;
#pragma unroll
  for (int v = 0; v < 4; ++v)
    out[o + v * plane] = in[o + v * plane];
}

// ------------------------------------------------------------------------------------------
// CFL reduction: ComputeDtFunctor, src/HydroRunFunctors.h:17-79.
// max is exact and order independent, so any reduction tree reproduces the reference's value.
// invDt is non-negative, so its IEEE bit pattern orders like an unsigned integer: the cross-block
// step is a single atomicMax on 64-bit integers (no CAS loop, no second kernel).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double
warp_max(double x)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

__device__ __forceinline__ void
block_max_to_global(double x, unsigned long long * __restrict__ bits)
{
  __shared__ double warp_part[32];
  const int         lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int         nw = (blockDim.x + 31) >> 5;
  x = warp_max(x);
  if (lane == 0)
    warp_part[w] = x;
  __syncthreads();
  if (w == 0)
  {
    x = lane < nw ? warp_part[lane] : 0.0;
    x = warp_max(x);
    if (lane == 0 && x > 0.0)
      atomicMax(bits, (unsigned long long)__double_as_longlong(x));
  }
}

__global__ void __launch_bounds__(256)
k_reduce_invdt(Geom g, Settings s, const double * __restrict__ U, unsigned long long * __restrict__ bits)
{
  const size_t plane = (size_t)g.isize * g.jsize;
  double       m = 0.0;
  for (int j = 2 + blockIdx.y; j < g.jsize - 2; j += gridDim.y)
    for (int i = 2 + blockIdx.x * blockDim.x + threadIdx.x; i < g.isize - 2; i += gridDim.x * blockDim.x)
    {
      const size_t o = (size_t)i + (size_t)g.isize * j;
      const double v = cfl_inv_dt(s, U[o], U[o + plane], U[o + 2 * plane], U[o + 3 * plane]);
      m = fmax(m, v); // fmax drops a NaN operand, like the reference's fmax(invDt, ...) at :72
    }
  block_max_to_global(m, bits);
}

// ------------------------------------------------------------------------------------------
// ConvertToPrimitivesFunctor, src/HydroRunFunctors.h:84-143 (whole array incl. ghosts)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_convert_to_primitives(Geom g, Settings s, const double * __restrict__ U, double * __restrict__ Q)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = grid_row();
  if (i >= g.isize || j >= g.jsize)
    return;
  double u[4], q[4];
  load4(U, g, i, j, u);
  compute_primitives_noc(s, u[ID], u[IP], u[IU], u[IV], q[ID], q[IP], q[IU], q[IV]);
  store4(Q, g, i, j, q);
}

// ------------------------------------------------------------------------------------------
// ComputeAndStoreFluxesFunctor, src/HydroRunFunctors.h:412-651 — the unfused flux kernel of
// implementation 0: thread (i,j) produces the flux through its west and south faces.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void
cell_slopes(const Settings & s, const Geom & g, const double * __restrict__ Q, int i, int j, double q[4],
            double dqX[4], double dqY[4])
{
  double qp[4], qm[4];
  load4(Q, g, i, j, q);
  load4(Q, g, i + 1, j, qp);
  load4(Q, g, i - 1, j, qm);
  slopes_dir(s, q, qp, qm, dqX);
  load4(Q, g, i, j + 1, qp);
  load4(Q, g, i, j - 1, qm);
  slopes_dir(s, q, qp, qm, dqY);
}

template <int SOLVER>
__global__ void __launch_bounds__(128)
k_compute_and_store_fluxes(Geom g, Settings s, const double * __restrict__ Q, double * __restrict__ Fx,
                           double * __restrict__ Fy, double dtdx, double dtdy)
{
  const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 2 + grid_row();
  if (i > g.isize - 2 || j > g.jsize - 2)
    return;
  double q[4], dqX[4], dqY[4], s0[4], qn[4], dqXn[4], dqYn[4], s0n[4], ql[4], qr[4], f[4];
  double f_d, f_e, f_n, f_t;

  cell_slopes(s, g, Q, i, j, q, dqX, dqY);
  trace_sources(s, q, dqX, dqY, s0);

  // west face (:519-575)
  cell_slopes(s, g, Q, i - 1, j, qn, dqXn, dqYn);
  trace_sources(s, qn, dqXn, dqYn, s0n);
  trace_face<-1>(s, q, dqX, s0, dtdx, qr);
  trace_face<+1>(s, qn, dqXn, s0n, dtdx, ql);
  riemann<SOLVER>(s, ql[ID], ql[IP], ql[IU], ql[IV], qr[ID], qr[IP], qr[IU], qr[IV], f_d, f_e, f_n, f_t);
  f[ID] = f_d * dtdx;
  f[IP] = f_e * dtdx;
  f[IU] = f_n * dtdx;
  f[IV] = f_t * dtdx;
  store4(Fx, g, i, j, f);

  // south face (:581-640), IU<->IV swapped around the solve
  cell_slopes(s, g, Q, i, j - 1, qn, dqXn, dqYn);
  trace_sources(s, qn, dqXn, dqYn, s0n);
  trace_face<-1>(s, q, dqY, s0, dtdy, qr);
  trace_face<+1>(s, qn, dqYn, s0n, dtdy, ql);
  riemann<SOLVER>(s, ql[ID], ql[IP], ql[IV], ql[IU], qr[ID], qr[IP], qr[IV], qr[IU], f_d, f_e, f_n, f_t);
  f[ID] = f_d * dtdy;
  f[IP] = f_e * dtdy;
  f[IU] = f_t * dtdy;
  f[IV] = f_n * dtdy;
  store4(Fy, g, i, j, f);
}

// UpdateFunctor, src/HydroRunFunctors.h:656-723
__global__ void __launch_bounds__(256)
k_update(Geom g, double * __restrict__ U, const double * __restrict__ Fx, const double * __restrict__ Fy)
{
  const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 2 + grid_row();
  if (i >= g.isize - 2 || j >= g.jsize - 2)
    return;
  const size_t plane = (size_t)g.isize * g.jsize;
  const size_t o = (size_t)i + (size_t)g.isize * j;
#pragma unroll
  for (int v = 0; v < 4; ++v)
  {
    double x = U[o + v * plane];
    x += Fx[o + v * plane];
    x -= Fx[o + 1 + v * plane];
    x += Fy[o + v * plane];
    x -= Fy[o + g.isize + v * plane];
    U[o + v * plane] = x;
  }
}

// ------------------------------------------------------------------------------------------
// implementation 1: ComputeSlopesFunctor :1061-1162, ComputeTraceAndFluxes_Functor<dir> :1167-1342,
// UpdateDirFunctor<dir> :986-1055
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_compute_slopes(Geom g, Settings s, const double * __restrict__ Q, double * __restrict__ Sx,
                 double * __restrict__ Sy)
{
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + grid_row();
  if (i > g.isize - 2 || j > g.jsize - 2)
    return;
  double q[4], dqX[4], dqY[4];
  cell_slopes(s, g, Q, i, j, q, dqX, dqY);
  store4(Sx, g, i, j, dqX);
  store4(Sy, g, i, j, dqY);
}

template <int SOLVER, int DIR>
__global__ void __launch_bounds__(128)
k_trace_and_fluxes(Geom g, Settings s, const double * __restrict__ Q, const double * __restrict__ Sx,
                   const double * __restrict__ Sy, double * __restrict__ F, double dtdx, double dtdy)
{
  const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 2 + grid_row();
  if (i > g.isize - 2 || j > g.jsize - 2)
    return;
  const int in = DIR == 1 ? i - 1 : i, jn = DIR == 1 ? j : j - 1;
  double    q[4], dqX[4], dqY[4], s0[4], ql[4], qr[4], f[4];
  double    f_d, f_e, f_n, f_t;
  load4(Q, g, i, j, q);
  load4(Sx, g, i, j, dqX);
  load4(Sy, g, i, j, dqY);
  trace_sources(s, q, dqX, dqY, s0);
  if (DIR == 1)
    trace_face<-1>(s, q, dqX, s0, dtdx, qr);
  else
    trace_face<-1>(s, q, dqY, s0, dtdy, qr);
  load4(Q, g, in, jn, q);
  load4(Sx, g, in, jn, dqX);
  load4(Sy, g, in, jn, dqY);
  trace_sources(s, q, dqX, dqY, s0);
  if (DIR == 1)
  {
    trace_face<+1>(s, q, dqX, s0, dtdx, ql);
    riemann<SOLVER>(s, ql[ID], ql[IP], ql[IU], ql[IV], qr[ID], qr[IP], qr[IU], qr[IV], f_d, f_e, f_n, f_t);
    f[ID] = f_d * dtdx;
    f[IP] = f_e * dtdx;
    f[IU] = f_n * dtdx;
    f[IV] = f_t * dtdx;
  }
  else
  {
    trace_face<+1>(s, q, dqY, s0, dtdy, ql);
    riemann<SOLVER>(s, ql[ID], ql[IP], ql[IV], ql[IU], qr[ID], qr[IP], qr[IV], qr[IU], f_d, f_e, f_n, f_t);
    f[ID] = f_d * dtdy;
    f[IP] = f_e * dtdy;
    f[IU] = f_t * dtdy;
    f[IV] = f_n * dtdy;
  }
  store4(F, g, i, j, f);
}

template <int DIR>
__global__ void __launch_bounds__(256)
k_update_dir(Geom g, double * __restrict__ U, const double * __restrict__ F)
{
  const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 2 + grid_row();
  if (i >= g.isize - 2 || j >= g.jsize - 2)
    return;
  const size_t plane = (size_t)g.isize * g.jsize;
  const size_t o = (size_t)i + (size_t)g.isize * j;
  const size_t on = DIR == 1 ? o + 1 : o + g.isize;
#pragma unroll
  for (int v = 0; v < 4; ++v)
  {
    double x = U[o + v * plane];
    x += F[o + v * plane];
    x -= F[on + v * plane];
    U[o + v * plane] = x;
  }
}

// ------------------------------------------------------------------------------------------
// The fused step (e2d_march.cuh)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void
st_release_sys_u64(unsigned long long * p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Epilogue of the peer-publishing instantiation (multi-GPU, e2d_slab.cu).  Kept out of line: inlined, its live
// values and addressing leak into the register allocation of the marching loop (ncu: +17 register moves per row and
// 4 % of the fast kernel's throughput).
// the rows [j0, j1) and the column a thread of this block owns, recomputed from the block indices: the epilogues take
// nothing from the marching loop's registers (values kept alive for them cost the loop a spill and three S2Rs)
// Which column block and row segment a thread block works on.  Blocks are handed to the SMs in the order of their
// linear index, and the first ones to become resident are the first ones to finish (measured: tools/step_timeline.py):
//   LOOP == 1 (peers): the two EDGE segments come first, so that the halo rows are on their way (and usually landed)
//             while the interior is still being computed;
//   LOOP == 2 (single-GPU loop), E2D_BORDER_FIRST=1 only (experiment, off): every block whose epilogue pushes
//             boundary cells comes first — the first and last segment, then the first and last column block of the
//             other segments, then the interior.  Measured (profiles/r3q_border_first_ab.txt): no gain — which of the
//             blocks sharing an SM finishes first is decided by the SM's warp slots, not by the block index.
#ifndef E2D_BORDER_FIRST
#  define E2D_BORDER_FIRST 0
#endif
template <int LOOP>
__device__ __forceinline__ void
block_place(int & bx, int & seg)
{
  bx = blockIdx.x;
  seg = blockIdx.y;
  const int nbx = gridDim.x, nseg = gridDim.y;
  if (LOOP == 1)
    seg = (blockIdx.y == 0) ? 0 : (blockIdx.y == 1 ? nseg - 1 : (int)blockIdx.y - 1);
  if (LOOP == 2 && E2D_BORDER_FIRST && nbx >= 3 && nseg >= 3)
  {
    int L = blockIdx.y * nbx + blockIdx.x;
    if (L < 2 * nbx)
    {
      seg = L < nbx ? 0 : nseg - 1;
      bx = L < nbx ? L : L - nbx;
      return;
    }
    L -= 2 * nbx;
    if (L < 2 * (nseg - 2))
    {
      bx = (L & 1) ? nbx - 1 : 0;
      seg = 1 + (L >> 1);
      return;
    }
    L -= 2 * (nseg - 2);
    seg = L / (nbx - 2);
    bx = 1 + (L - seg * (nbx - 2));
    seg += 1;
  }
}

template <int BX, int LOOP>
__device__ __forceinline__ bool
block_rows(const MarchArgs & a, int & j0, int & j1, int & bx)
{
  int seg;
  block_place<LOOP>(bx, seg);
  return segment_rows(a, seg, j0, j1);
}

template <int BX>
__device__ __noinline__ void
publish_to_peers(const MarchArgs & a, const FusedLink & link)
{
  int        j0, j1, bx;
  const bool active = block_rows<BX, 1>(a, j0, j1, bx);
  const int  t = threadIdx.x, i = bx * (BX - 4) + t;
  const bool store = (t >= 2) && (t <= BX - 3) && (i >= 2) && (i <= a.isize - 3);
  const bool lo = active && a.peer_lo && j0 <= 3 && j1 > 2;
  const bool hi = active && a.peer_hi && j0 <= a.jsize - 3 && j1 > a.jsize - 4;
  // Edge segments: every thread copies the edge rows of ITS column (which it stored itself a moment ago: program
  // order makes them visible to it) into the neighbour's ghost rows of the same-parity array.
  if ((lo || hi) && store)
  {
    const size_t plane = (size_t)a.isize * a.jsize;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      // k = 0, 1: my rows 2, 3 -> the lower neighbour's top ghost rows
      // k = 2, 3: my last two interior rows -> the upper neighbour's rows 0, 1
      const bool to_lo = k < 2;
      const int  jr = to_lo ? 2 + k : a.jsize - 4 + (k - 2);
      if (!(to_lo ? lo : hi) || jr < j0 || jr >= j1)
        continue;
      const double * src = a.Uout + (jr * a.isize + i);
      const int      jsize_d = to_lo ? a.peer_lo_jsize : a.peer_hi_jsize;
      const int      jd = to_lo ? jsize_d - 2 + k : k - 2;
      double *       dst = (to_lo ? a.peer_lo : a.peer_hi) + (jd * a.isize + i);
      const size_t   plane_d = (size_t)a.isize * jsize_d;
#pragma unroll
      for (int v = 0; v < 4; ++v)
        dst[v * plane_d] = src[v * plane];
    }
  }
  // Publication (e2d_slab.cu): every block fences its stores (peer rows, atomicMax), then three elections by
  // arrival count.  The last block holding the lower / upper edge rows raises the neighbour's halo flag; the
  // last block of the grid copies the finished invDt partial into every rank's slot and raises the invDt flags.
  // (only the edge blocks have peer stores to drain at system scope; for the others the device-scope fence orders
  //  their atomicMax before their arrival count, which is all the last block's read needs)
  if (lo || hi)
    __threadfence_system();
  else
    __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    if (lo && atomicAdd(&link.cnt[0], 1u) == link.n_lo - 1)
    {
      link.cnt[0] = 0;
      __threadfence_system();
      st_release_sys_u64(link.flag_lo, link.seq_next);
    }
    if (hi && atomicAdd(&link.cnt[1], 1u) == link.n_hi - 1)
    {
      link.cnt[1] = 0;
      __threadfence_system();
      st_release_sys_u64(link.flag_hi, link.seq_next);
    }
    if (atomicAdd(&link.cnt[2], 1u) == link.n_all - 1)
    {
      link.cnt[2] = 0;
      __threadfence_system();
      const unsigned long long part = atomicMax(a.invdt_bits, 0ull); // every block's atomicMax precedes its count
      for (int k = 0; k < link.nranks; ++k)
        link.comm[k]->invdt_slot[link.parity_next][link.rank] = part;
      __threadfence_system();
      for (int k = 0; k < link.nranks; ++k)
        st_release_sys_u64(&link.comm[k]->invdt_flag[link.rank], link.seq_next);
    }
  }
}

// Epilogue of the single-GPU loop instantiation (SoloLoop, e2d_internal.h).  Out of line for the same reason.
//     Boundary push: every ghost cell whose SOURCE cell (e2d_bc.cuh: bc_map) this block has just produced is written
//     now, into the output array — the fill HydroRun::make_boundaries would do at the start of the next step
//     (src/HydroRun.h:296, :390-399), same composed index maps, same signs, bit for bit.  Only blocks that hold one of
//     the columns {2, 3, nx, nx+1} or rows {2, 3, ny, ny+1} have anything to do.
// dt of the step a SoloLoop launch performs, from the device-resident scalars (HydroRun.h:246, main.cpp:131-134)
__device__ __forceinline__ double
solo_dt(const SoloLoop & solo, double t)
{
  const double invDt = __longlong_as_double((long long)solo.st->solo_acc[solo.step & 3]);
  double       dt = solo.cfl / invDt;
  if (t + dt > solo.tEnd)
    dt = solo.tEnd - t;
  return dt;
}

// Opens the step of a SoloLoop launch: every thread derives the same dt from the same device-resident scalars
// (HydroRun.h:246, main.cpp:100,131-134), which no block of THIS launch writes (rings, see SlabState); one thread of
// the grid does the bookkeeping of main.cpp:142-143 on the way in.  Returns -1 once the loop is over.  Out of line so
// that none of this competes for the marching loop's registers.
__device__ __noinline__ double
solo_open_step(const SoloLoop & solo)
{
  SlabState *  st = solo.st;
  const int    k = solo.step & 3;
  const double t = st->solo_T[k];
  const bool   over = !(t < solo.tEnd && solo.step < solo.max_steps); // main.cpp:100
  const double dt = over ? -1.0 : solo_dt(solo, t);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
  {
    st->solo_T[(k + 1) & 3] = over ? t : t + dt;
    st->solo_acc[(k + 2) & 3] = 0ull; // the accumulator of the next launch
    if (!over)
    {
      if (solo.dt_hist && solo.step < solo.hist_cap)
        solo.dt_hist[solo.step] = dt;
      st->t = t + dt;
      st->dt = dt;
      st->nStep = solo.step + 1;
      st->done = !(t + dt < solo.tEnd && solo.step + 1 < solo.max_steps);
    }
    else
      st->solo_acc[(k + 1) & 3] = st->solo_acc[k]; // nothing changes any more: carry the state's invDt along
  }
  return dt;
}

template <int BX>
__device__ __noinline__ void
solo_epilogue(const MarchArgs & a, const SoloLoop & solo)
{
  int        j0, j1, bx;
  const bool active = block_rows<BX, 2>(a, j0, j1, bx);
  __syncthreads(); // this block's stores to Uout are visible to all its threads
  if (active)
  {
    Geom g;
    g.isize = a.isize;
    g.jsize = a.jsize;
    g.nx = a.isize - 4;
    g.ny = a.jsize - 4;
    g.j_off = 0;
    BcArgs bc;
    bc.bc_xmin = solo.bc_xmin;
    bc.bc_xmax = solo.bc_xmax;
    bc.bc_ymin = solo.bc_ymin;
    bc.bc_ymax = solo.bc_ymax;
    bc.faces = E2D_FACES_ALL;
    const int    nx = g.nx, ny = g.ny;
    const int    I0 = bx * (BX - 4) + 2;                                   // produced columns [I0, I1)
    const int    I1 = (I0 + BX - 4 < a.isize - 2) ? I0 + BX - 4 : a.isize - 2;
    const size_t plane = (size_t)a.isize * a.jsize;
    const bool   col_src = (I0 <= 3) || (I1 > nx); // holds a column of {2, 3, nx, nx+1}
    const bool   row_src = (j0 <= 3) || (j1 > ny); // holds a row of {2, 3, ny, ny+1}
    if (col_src)
    { // x-ghost columns of the rows [j0, j1): a long segment gives every thread several cells; their loads are issued
      // together, four cells at a time, before the first store (loads and stores go through the same pointer, so the
      // compiler would otherwise serialise one L2 round trip per cell — ~10 us at the tail of the border blocks of a
      // 249-row segment)
      const int n = 4 * (j1 - j0);
      for (int k0 = threadIdx.x; k0 < n; k0 += 4 * BX)
      {
        double val[4][4];
        size_t dst[4];
        bool   take[4];
#pragma unroll
        for (int b = 0; b < 4; ++b)
        {
          const int k = k0 + b * BX;
          take[b] = k < n;
          const int kk = take[b] ? k : k0;
          const int gsel = kk & 3, j = j0 + (kk >> 2);
          const int i = gsel < 2 ? gsel : nx + gsel;
          int       i0, jj0;
          bool      in_x, in_y, flip_u, flip_v;
          bc_map(g, bc, i, j, i0, jj0, in_x, in_y, flip_u, flip_v);
          take[b] = take[b] && i0 >= I0 && i0 < I1;
          if (!take[b])
            i0 = I0; // a cell of this block: the load stays valid, the value is dropped
          dst[b] = (size_t)i + (size_t)a.isize * j;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            val[b][v] = bc_value(a.Uout, (size_t)i0 + (size_t)a.isize * jj0 + v * plane, v, in_x, in_y, flip_u, flip_v);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (take[b])
          {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              a.Uout[dst[b] + v * plane] = val[b][v];
          }
      }
    }
    if (row_src)
    { // y-ghost rows: this block's own columns and the four x-ghost columns (corners); loads batched like above
      const int own = I1 - I0, ncols = own + 4, n = 4 * ncols;
      for (int k0 = threadIdx.x; k0 < n; k0 += 4 * BX)
      {
        double val[4][4];
        size_t dst[4];
        bool   take[4];
#pragma unroll
        for (int b = 0; b < 4; ++b)
        {
          const int k = k0 + b * BX;
          take[b] = k < n;
          const int kk = take[b] ? k : k0;
          const int gsel = kk / ncols, c = kk - gsel * ncols;
          const int j = gsel < 2 ? gsel : ny + gsel;
          const int i = c < own ? I0 + c : ((c - own) < 2 ? (c - own) : nx + (c - own));
          int       i0, jj0;
          bool      in_x, in_y, flip_u, flip_v;
          bc_map(g, bc, i, j, i0, jj0, in_x, in_y, flip_u, flip_v);
          take[b] = take[b] && i0 >= I0 && i0 < I1 && jj0 >= j0 && jj0 < j1;
          if (!take[b])
          { // a cell of this block: the load stays valid, the value is dropped
            i0 = I0;
            jj0 = j0;
          }
          dst[b] = (size_t)i + (size_t)a.isize * j;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            val[b][v] = bc_value(a.Uout, (size_t)i0 + (size_t)a.isize * jj0 + v * plane, v, in_x, in_y, flip_u, flip_v);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (take[b])
          {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              a.Uout[dst[b] + v * plane] = val[b][v];
          }
      }
    }
  }
}

// E2D_TIMELINE=1 (development build, tools/step_timeline.py): every block of the single-GPU loop kernel stamps the
// global nanosecond timer at its milestones into a device buffer that e2d_debug_timeline() copies out — where the time
// of a small-grid step goes (profiles/r3n_step_timeline.txt).  Compiled out of the product.
#ifndef E2D_TIMELINE
#  define E2D_TIMELINE 0
#endif
#if E2D_TIMELINE
constexpr int                 kTimelineBlocks = 1024, kTimelineMarks = 8;
__device__ unsigned long long g_timeline[4][kTimelineBlocks][kTimelineMarks]; // [step & 3]: the last four steps
__device__ __forceinline__ unsigned long long
timeline_now()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void
timeline_mark(int step, int k, unsigned long long t)
{
  const unsigned b = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0 && b < (unsigned)kTimelineBlocks)
    g_timeline[step & 3][b][k] = t;
}
#  define E2D_MARK(k) timeline_mark(solo.step, k, timeline_now())
#else
#  define E2D_MARK(k)
#endif

// LOOP: 0 = one step, dt from the arguments;  1 = multi-GPU slab loop (publishes halo rows + invDt partial to the
// peers, e2d_slab.cu);  2 = single-GPU loop with one launch per step (SoloLoop: dt, boundary push, bookkeeping).
// PEEL: the march with the first and last row of a segment treated apart (MarchThread::phaseB<1>, <2>: three of a
// segment's Riemann solves serve nobody).  A kernel of its own, used for SHORT uniform segments only: in the one kernel
// the two extra phase A bodies cost the long marches 4 % (code size, spills: profiles/r2a_variants.txt, r4m), while a
// grid of 2- to 64-row segments gains 2-9 %.
template <int SOLVER, bool FUSE_DT, int LOOP, int MATH = 0, int TYP = 0, bool PEEL = false>
__global__ void __launch_bounds__(kBX, march_min_blocks(MATH))
k_fused_step(const __grid_constant__ MarchArgs a, const int * __restrict__ d_done,
             const __grid_constant__ FusedLink link, const __grid_constant__ SoloLoop solo)
{
#if E2D_TIMELINE
  const unsigned long long tl0 = timeline_now();
#endif
  pdl_wait_for_predecessor(); // (no-ops unless launched with the programmatic-serialization attribute)
  pdl_release_successor();
#if E2D_TIMELINE
  const unsigned long long tl1 = timeline_now();
#endif
  if (d_done && *d_done)
    return;
  double dt = a.d_dt ? *a.d_dt : a.dt;
  if (LOOP == 2)
  {
    dt = solo_open_step(solo);
    if (dt < 0.0) // the loop is over (dt is positive otherwise: cfl / invDt or tEnd - t with t < tEnd)
      return;
  }
#if E2D_TIMELINE
  timeline_mark(solo.step, 0, tl0); // (a launch past the end of the loop has returned above and leaves no marks)
  timeline_mark(solo.step, 1, tl1);
#endif
  E2D_MARK(2);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MarchSmem<kBX> &                  sm = *reinterpret_cast<MarchSmem<kBX> *>(smem_raw);
  MarchThread<kBX, SOLVER, FUSE_DT, MATH, TYP> th;
  int bx, seg;
  block_place<LOOP>(bx, seg);
  const bool active = th.init(a, sm, threadIdx.x, bx, seg, dt);
#if E2D_TIMELINE
  {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    timeline_mark(solo.step, 7, (unsigned long long)smid | ((unsigned long long)bx << 16) | ((unsigned long long)seg << 32));
  }
#endif
  if (active)
  {
    __syncthreads();
    E2D_MARK(3);
    if (PEEL)
    {
      th.phaseA(a, sm, th.j0 - 1);
      __syncthreads();
      th.template phaseB<1>(a, sm, th.j0 - 1);
#pragma unroll 1
      for (int r = th.j0; r < th.j1; ++r)
      {
        th.phaseA(a, sm, r);
        __syncthreads();
        th.template phaseB<0>(a, sm, r);
      }
      th.phaseA(a, sm, th.j1);
      __syncthreads();
      th.template phaseB<2>(a, sm, th.j1);
    }
    else
    {
#pragma unroll(MATH == 1 ? kMarchUnrollFast : kMarchUnrollStrict)
      for (int r = th.j0 - 1; r <= th.j1; ++r)
      {
        th.phaseA(a, sm, r);
        __syncthreads();
        th.phaseB(a, sm, r);
      }
    }
    E2D_MARK(4);
    th.finish(a);
    if (FUSE_DT && a.invdt_bits)
      block_max_to_global(th.invdt, a.invdt_bits);
    E2D_MARK(5);
  }
  if (LOOP == 1)
    publish_to_peers<kBX>(a, link);
  if (LOOP == 2)
    solo_epilogue<kBX>(a, solo);
  E2D_MARK(6);
}

// ------------------------------------------------------------------------------------------
// scalar bookkeeping of the device-resident loop (src/main.cpp:100-143)
// ------------------------------------------------------------------------------------------
__global__ void
k_loop_begin_step(LoopState * st, double cfl, double tEnd)
{
  if (st->done)
    return;
  const double invDt = __longlong_as_double((long long)st->invdt_cur);
  double       dt = cfl / invDt; // HydroRun.h:246
  if (st->t + dt > tEnd)         // main.cpp:131-134
    dt = tEnd - st->t;
  st->dt = dt;
  st->invdt_next = 0ull;
}

__global__ void
k_loop_end_step(LoopState * st, double tEnd, int max_steps, double * dt_hist, long hist_cap)
{
  if (st->done)
    return;
  if (dt_hist && st->nStep < hist_cap)
    dt_hist[st->nStep] = st->dt;
  st->nStep += 1; // main.cpp:142-143
  st->t += st->dt;
  st->invdt_cur = st->invdt_next;
  st->done = !(st->t < tEnd && st->nStep < max_steps); // main.cpp:100
}

// ------------------------------------------------------------------------------------------
// function-level evaluation (known-answer tests of the __device__ functions)
// ------------------------------------------------------------------------------------------
__global__ void
k_eval(Settings s, int func, const double * __restrict__ in, double * __restrict__ out, long n)
{
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n)
    return;
  switch (func)
  {
    case 0:
    { // prim
      const double * a = in + 4 * r;
      double *       o = out + 5 * r;
      compute_primitives(s, a[ID], a[IP], a[IU], a[IV], o[ID], o[IP], o[IU], o[IV], o[4]);
      break;
    }
    case 1:
    { // slope
      const double * a = in + 20 * r;
      double *       o = out + 8 * r;
      double         dq[4];
      slopes_dir(s, a, a + 4, a + 8, dq);
      for (int v = 0; v < 4; ++v)
        o[v] = dq[v];
      slopes_dir(s, a, a + 12, a + 16, dq);
      for (int v = 0; v < 4; ++v)
        o[4 + v] = dq[v];
      break;
    }
    case 2:
    { // trace
      const double * a = in + 14 * r;
      double *       o = out + 16 * r;
      double         s0[4], f[4];
      trace_sources(s, a, a + 4, a + 8, s0);
      trace_face<-1>(s, a, a + 4, s0, a[12], f);
      for (int v = 0; v < 4; ++v)
        o[v] = f[v];
      trace_face<+1>(s, a, a + 4, s0, a[12], f);
      for (int v = 0; v < 4; ++v)
        o[4 + v] = f[v];
      trace_face<-1>(s, a, a + 8, s0, a[13], f);
      for (int v = 0; v < 4; ++v)
        o[8 + v] = f[v];
      trace_face<+1>(s, a, a + 8, s0, a[13], f);
      for (int v = 0; v < 4; ++v)
        o[12 + v] = f[v];
      break;
    }
    case 3:
    case 6:
    { // hllc / hll
      const double * a = in + 8 * r;
      double *       o = out + 4 * r;
      if (func == 3)
        riemann_hllc(s, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP],
                     o[IU], o[IV]);
      else
        riemann_hll(s, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP],
                    o[IU], o[IV]);
      break;
    }
    case 4:
    { // approx
      const double * a = in + 8 * r;
      double *       o = out + 8 * r;
      riemann_approx(s, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP],
                     o[IU], o[IV], o[4 + ID], o[4 + IP], o[4 + IU], o[4 + IV]);
      break;
    }
    case 7:
    { // hllc_lean: the fused kernel's solver, fast path then (guard failed) plain operators; out[4] = guard
      const double *   a = in + 8 * r;
      double *         o = out + 5 * r;
      const StepConsts c = make_step_consts(s);
      bool             ok = true;
      hllc_lean<true>(s, c, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP], o[IU],
                      o[IV], ok);
      o[4] = ok ? 1.0 : 0.0;
      if (!ok)
        hllc_lean<false>(s, c, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP],
                         o[IU], o[IV], ok);
      break;
    }
    case 8:
    { // prim_lean + cfl_lean: q[4], invDt integrand, guard
      const double *   a = in + 4 * r;
      double *         o = out + 6 * r;
      const StepConsts c = make_step_consts(s);
      bool             ok = true, unused = true;
      Recip            rd;
      const Recip      rdx = recip_of<true, false>(s.dx, unused), rdy = recip_of<true, false>(s.dy, unused);
      prim_lean<true>(s, c, a, o, rd, ok);
      o[4] = cfl_lean<true>(s, c, rdx, rdy, a, ok);
      o[5] = ok ? 1.0 : 0.0;
      if (!ok)
      {
        prim_lean<false>(s, c, a, o, rd, ok);
        o[4] = cfl_lean<false>(s, c, rdx, rdy, a, ok);
      }
      break;
    }
    case 9:
    { // slope_lean + trace_sources_lean + faces: same record as "trace" but slopes are recomputed from the
      // 5-point stencil: in = q, qPlusX, qMinusX, qPlusY, qMinusY, dtdx, dtdy (22); out = dqX, dqY, 4 faces, guard (25)
      const double * a = in + 22 * r;
      double *       o = out + 25 * r;
      const bool     limited = (s.slope_type == 1.0) || (s.slope_type == 2.0);
      double         dqX[4], dqY[4], s0[4], f[4];
      slopes_lean(s.slope_type, limited, a, a + 4, a + 8, dqX);
      slopes_lean(s.slope_type, limited, a, a + 12, a + 16, dqY);
      bool  ok = true, unused = true;
      Recip rd = recip_of<true, true>(a[ID], unused);
      trace_sources_lean<true>(s, a, rd, dqX, dqY, s0, ok);
      o[24] = ok ? 1.0 : 0.0;
      if (!ok)
        trace_sources_lean<false>(s, a, rd, dqX, dqY, s0, ok);
      for (int v = 0; v < 4; ++v)
      {
        o[v] = dqX[v];
        o[4 + v] = dqY[v];
      }
      trace_face_lean<-1>(s, a, dqX, s0, a[20], f);
      for (int v = 0; v < 4; ++v)
        o[8 + v] = f[v];
      trace_face_lean<+1>(s, a, dqX, s0, a[20], f);
      for (int v = 0; v < 4; ++v)
        o[12 + v] = f[v];
      trace_face_lean<-1>(s, a, dqY, s0, a[21], f);
      for (int v = 0; v < 4; ++v)
        o[16 + v] = f[v];
      trace_face_lean<+1>(s, a, dqY, s0, a[21], f);
      for (int v = 0; v < 4; ++v)
        o[20 + v] = f[v];
      break;
    }
    case 10:
    { // div: a, d -> shared-reciprocal quotient (ZERO_OK), its guard, the `/` operator, same with ZERO_OK off
      const double * a = in + 2 * r;
      double *       o = out + 5 * r;
      bool           ok = true, ok2 = true;
      const Recip    rp = recip_of<true, true>(a[1], ok);
      o[0] = div_by<true, true>(a[0], rp, ok);
      o[1] = ok ? 1.0 : 0.0;
      o[2] = a[0] / a[1];
      const Recip rn = recip_of<true, false>(a[1], ok2);
      o[3] = div_by<true, false>(a[0], rn, ok2);
      o[4] = ok2 ? 1.0 : 0.0;
      break;
    }
    case 11:
    { // sqrt: x -> lean sqrt, guard, sqrt()
      const double * a = in + r;
      double *       o = out + 3 * r;
      bool           ok = true;
      o[0] = sqrt_pos<true>(a[0], ok);
      o[1] = ok ? 1.0 : 0.0;
      o[2] = sqrt(a[0]);
      break;
    }
    case 12:
    { // fast_div: (a, d) -> a * fast::rcp(d), a / d
      const double * a = in + 2 * r;
      double *       o = out + 2 * r;
      o[0] = a[0] * fast::rcp(a[1]);
      o[1] = a[0] / a[1];
      break;
    }
    case 13:
    { // fast_sqrt: x -> fast::sqrt_pos(x), sqrt(x)
      const double * a = in + r;
      double *       o = out + 2 * r;
      o[0] = fast::sqrt_pos(a[0]);
      o[1] = sqrt(a[0]);
      break;
    }
    case 14:
    { // fast_hllc: same record as hllc
      const double *   a = in + 8 * r;
      double *         o = out + 4 * r;
      const StepConsts c = make_step_consts(s);
      fast::hllc(s, c, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP], o[IU],
                 o[IV]);
      break;
    }
    case 15:
    { // fast_cell: u[4] -> q[4], cfl integrand
      const double *   a = in + 4 * r;
      double *         o = out + 5 * r;
      const StepConsts c = make_step_consts(s);
      double           ry;
      fast::prim(s, c, a, o, ry);
      o[4] = fast::cfl_tail(s, 1.0 / s.dx, 1.0 / s.dy, o, ry);
      break;
    }
    case 16:
    { // fast_slope: same record as slope
      const double * a = in + 20 * r;
      double *       o = out + 8 * r;
      const bool     limited = (s.slope_type == 1.0) || (s.slope_type == 2.0);
      fast::slopes(limited ? s.slope_type : 0.0, a, a + 4, a + 8, o);
      fast::slopes(limited ? s.slope_type : 0.0, a, a + 12, a + 16, o + 4);
      break;
    }
    case 17:
    { // fast_trace: same record as trace (q, dqX, dqY, dtdx, dtdy) -> xmin, xmax, ymin, ymax
      const double * a = in + 14 * r;
      double *       o = out + 16 * r;
      fast::trace(s, a, fast::rcp(a[ID]), a + 4, a + 8, 0.5 * a[12], 0.5 * a[13], o, o + 4, o + 8, o + 12);
      break;
    }
    case 18:
    { // rusanov: same record as hllc
      const double * a = in + 8 * r;
      double *       o = out + 4 * r;
      riemann_rusanov(s, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP], o[IU],
                      o[IV]);
      break;
    }
    case 5:
    { // cmpflx
      const double * a = in + 4 * r;
      double *       o = out + 4 * r;
      cmpflx(s, a[ID], a[IP], a[IU], a[IV], o[ID], o[IP], o[IU], o[IV]);
      break;
    }
  }
}

} // namespace

// ==========================================================================================
// host side
// ==========================================================================================
Settings
make_settings(const e2d_params & p)
{
  Settings s;
  s.gamma0 = p.gamma0;
  s.gamma6 = p.gamma6;
  s.cfl = p.cfl;
  s.slope_type = p.slope_type;
  s.smallr = p.smallr;
  s.smallc = p.smallc;
  s.smallp = p.smallp;
  s.smallpp = p.smallpp;
  s.dx = p.dx;
  s.dy = p.dy;
  return s;
}

Geom
make_geom(const e2d_params & p, int jsize_loc, int j_off)
{
  Geom g;
  g.isize = p.isize;
  g.jsize = jsize_loc;
  g.nx = p.nx;
  g.ny = jsize_loc - 2 * p.ghostWidth;
  g.j_off = j_off;
  return g;
}

int
solver_for(const e2d_params & p)
{
  // the reference parses `riemann=` and never reads it: every kernel hard-codes riemann_hllc
  // (src/HydroRunFunctors.h:567,631)
  return p.honourRiemannSolver ? p.riemannSolverType : E2D_RIEMANN_HLLC;
}

// one block row per grid row; gridDim.y may not exceed 65535, so tall slabs spill into gridDim.z (grid_row())
static inline dim3
grid_rows(int ncols, int nrows, int block)
{
  const int ycap = 32768;
  const int ny = nrows < ycap ? nrows : ycap;
  return dim3((unsigned)((ncols + block - 1) / block), (unsigned)(ny < 1 ? 1 : ny), (unsigned)((nrows + ycap - 1) / ycap > 0 ? (nrows + ycap - 1) / ycap : 1));
}

cudaError_t
launch_init_problem(const e2d_params & p, const Geom & g, double * U, cudaStream_t st, unsigned long long * n_inside_out,
                    long long n_inside_given, int count_jlo, int count_jhi)
{
  InitArgs a;
  a.problem = p.problemType;
  a.xmin = p.xmin;
  a.ymin = p.ymin;
  a.dx = p.dx;
  a.dy = p.dy;
  a.gamma0 = p.gamma0;
  a.blast_radius = p.blast_radius;
  a.blast_center_x = p.blast_center_x;
  a.blast_center_y = p.blast_center_y;
  a.blast_density_in = p.blast_density_in;
  a.blast_density_out = p.blast_density_out;
  a.blast_pressure_in = p.blast_pressure_in;
  a.blast_pressure_out = p.blast_pressure_out;
  a.blast_energy_density = -1.0;
  a.bubble_radius = p.bubble_radius;
  a.bubble_center_x = p.bubble_center_x;
  a.bubble_center_y = p.bubble_center_y;
  a.bubble_density = p.bubble_density;
  a.bubble_pressure = p.bubble_pressure;
  a.preshock_density = p.preshock_density;
  a.preshock_pressure = p.preshock_pressure;
  a.postshock_density = p.postshock_density;
  a.postshock_pressure = p.postshock_pressure;
  a.postshock_velocity = p.postshock_velocity;
  a.shock_loc = p.shock_loc;

  const dim3 grid = grid_rows(g.isize, g.jsize, 128);
  if (count_jhi <= 0)
    count_jhi = g.jsize;
  if (p.problemType == E2D_PROBLEM_BLAST && p.blast_total_energy_inside > 0)
  {
    // Sedov variant (:1445-1463): count the cells inside the disc, form the volume the way a serial
    // Kokkos::Sum would (repeated += dx*dy), then overwrite the energy inside with E_tot / volume.
    // A slab counts its own rows and hands the number out (n_inside_out); the caller sums over the ranks and comes
    // back with the global count (n_inside_given >= 0): integers, hence exact and order independent.
    unsigned long long n_inside = 0;
    if (n_inside_given >= 0)
      n_inside = (unsigned long long)n_inside_given;
    else
    {
      unsigned long long * d_n = nullptr;
      cudaError_t          e = cudaMalloc(&d_n, sizeof(unsigned long long));
      if (e != cudaSuccess)
        return e;
      cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), st);
      k_init_problem<<<grid, 128, 0, st>>>(g, a, U, d_n, count_jlo, count_jhi);
      count_launch();
      cudaMemcpyAsync(&n_inside, d_n, sizeof n_inside, cudaMemcpyDeviceToHost, st);
      e = cudaStreamSynchronize(st);
      cudaFree(d_n);
      if (e != cudaSuccess)
        return e;
      if (n_inside_out)
      { // the caller completes the initialisation once it knows the global count
        *n_inside_out = n_inside;
        return cudaGetLastError();
      }
    }
    double volume = 0.0;
    for (unsigned long long k = 0; k < n_inside; ++k)
      volume += p.dx * p.dy;
    a.blast_energy_density = p.blast_total_energy_inside / volume;
  }
  k_init_problem<<<grid, 128, 0, st>>>(g, a, U, nullptr, 0, 0);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_make_boundaries(const e2d_params & p, const Geom & g, double * U, int faces, const int * d_done,
                       cudaStream_t st)
{
  const BcArgs a = make_bc_args(p, faces);
  const int    n = 4 * g.isize + 4 * g.jsize;
  k_make_boundaries<<<(n + 127) / 128, 128, 0, st>>>(g, a, U, d_done);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_bc_x_rows(const e2d_params & p, const Geom & g, double * U, int faces, int jlo, int jhi, cudaStream_t st)
{
  if (jhi <= jlo || !(faces & E2D_FACES_X))
    return cudaSuccess;
  const int n = 4 * (jhi - jlo);
  k_bc_x_rows<<<(n + 127) / 128, 128, 0, st>>>(g, make_bc_args(p, faces & E2D_FACES_X), U, jlo, jhi);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_copy_ghost_frame(const Geom & g, const double * in, double * out, cudaStream_t st)
{
  const int n = 4 * g.isize + 4 * (g.jsize - 4);
  k_copy_ghost_frame<<<(n + 127) / 128, 128, 0, st>>>(g, in, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_reduce_invdt(const e2d_params & p, const Geom & g, const double * U, unsigned long long * d_bits,
                    cudaStream_t st)
{
  const int bx = (g.nx + 255) / 256;
  int       by = g.ny < 1 ? 1 : g.ny;
  // enough blocks to fill the machine, few enough that the atomics stay negligible
  const int max_by = (device_sm_count() * 8 + bx - 1) / bx;
  if (by > max_by)
    by = max_by;
  k_reduce_invdt<<<dim3(bx < 1 ? 1 : bx, by), 256, 0, st>>>(g, make_settings(p), U, d_bits);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_convert_to_primitives(const e2d_params & p, const Geom & g, const double * U, double * Q, cudaStream_t st)
{
  k_convert_to_primitives<<<grid_rows(g.isize, g.jsize, 256), 256, 0, st>>>(g, make_settings(p), U, Q);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_compute_and_store_fluxes(const e2d_params & p, const Geom & g, const double * Q, double * Fx, double * Fy,
                                double dtdx, double dtdy, cudaStream_t st)
{
  const dim3     grid = grid_rows(g.isize - 3, g.jsize - 3, 128);
  const Settings s = make_settings(p);
  switch (solver_for(p))
  {
    case E2D_RIEMANN_APPROX:
      k_compute_and_store_fluxes<0><<<grid, 128, 0, st>>>(g, s, Q, Fx, Fy, dtdx, dtdy);
      break;
    case E2D_RIEMANN_HLL:
      k_compute_and_store_fluxes<1><<<grid, 128, 0, st>>>(g, s, Q, Fx, Fy, dtdx, dtdy);
      break;
    case E2D_RIEMANN_RUSANOV:
      k_compute_and_store_fluxes<3><<<grid, 128, 0, st>>>(g, s, Q, Fx, Fy, dtdx, dtdy);
      break;
    default:
      k_compute_and_store_fluxes<2><<<grid, 128, 0, st>>>(g, s, Q, Fx, Fy, dtdx, dtdy);
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_update(const e2d_params &, const Geom & g, double * U, const double * Fx, const double * Fy, cudaStream_t st)
{
  k_update<<<grid_rows(g.nx, g.ny, 256), 256, 0, st>>>(g, U, Fx, Fy);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_compute_slopes(const e2d_params & p, const Geom & g, const double * Q, double * Sx, double * Sy,
                      cudaStream_t st)
{
  k_compute_slopes<<<grid_rows(g.isize - 2, g.jsize - 2, 128), 128, 0, st>>>(g, make_settings(p), Q, Sx, Sy);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_trace_and_fluxes(const e2d_params & p, const Geom & g, const double * Q, const double * Sx, const double * Sy,
                        double * F, double dtdx, double dtdy, int dir, cudaStream_t st)
{
  const dim3     grid = grid_rows(g.isize - 3, g.jsize - 3, 128);
  const Settings s = make_settings(p);
  const int      sol = solver_for(p);
#define E2D_TF(SOL, DIR) k_trace_and_fluxes<SOL, DIR><<<grid, 128, 0, st>>>(g, s, Q, Sx, Sy, F, dtdx, dtdy)
  if (dir == 1)
  {
    if (sol == 0)
      E2D_TF(0, 1);
    else if (sol == 1)
      E2D_TF(1, 1);
    else if (sol == 3)
      E2D_TF(3, 1);
    else
      E2D_TF(2, 1);
  }
  else
  {
    if (sol == 0)
      E2D_TF(0, 2);
    else if (sol == 1)
      E2D_TF(1, 2);
    else if (sol == 3)
      E2D_TF(3, 2);
    else
      E2D_TF(2, 2);
  }
#undef E2D_TF
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_update_dir(const e2d_params &, const Geom & g, double * U, const double * F, int dir, cudaStream_t st)
{
  if (dir == 1)
    k_update_dir<1><<<grid_rows(g.nx, g.ny, 256), 256, 0, st>>>(g, U, F);
  else
    k_update_dir<2><<<grid_rows(g.nx, g.ny, 256), 256, 0, st>>>(g, U, F);
  count_launch();
  return cudaGetLastError();
}

// The refined reciprocal the strict division sequence uses for a denominator (e2d_lean.cuh: recip_of), evaluated on
// the device once per distinct value and cached: dx and dy enter the CFL integrand of every cell, and as kernel
// arguments they cost no registers.  NaN on failure: the fast-path guards then reject and the plain operators run.
namespace
{
__global__ void
k_refined_reciprocal(double d, double * out)
{
  bool ok = true;
  *out = recip_of<true, false>(d, ok).y;
}
} // namespace

double
refined_reciprocal(double d)
{
  static std::mutex                             mu;
  static std::vector<std::pair<double, double>> cache;
  std::lock_guard<std::mutex>                   lock(mu);
  for (const auto & e : cache)
    if (e.first == d)
      return e.second;
  double   y = std::numeric_limits<double>::quiet_NaN();
  double * dev = nullptr;
  if (cudaMalloc(&dev, sizeof(double)) == cudaSuccess)
  {
    k_refined_reciprocal<<<1, 1>>>(d, dev);
    count_launch();
    if (cudaMemcpy(&y, dev, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      y = std::numeric_limits<double>::quiet_NaN();
    cudaFree(dev);
  }
  if (y == y)
    cache.emplace_back(d, y);
  return y;
}

// multiprocessors of the current device (148 on a B200), cached per device
int
device_sm_count()
{
  static std::atomic<int> cached[64] = {};
  int                     dev = 0;
  cudaGetDevice(&dev);
  int n = cached[dev & 63].load(std::memory_order_relaxed);
  if (n <= 0)
  {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

// Rows per block segment.  Every segment re-traces 2 rows and re-converts 3, so segments should be long; but
// the grid (nbx column blocks x nseg segments) should also fill the SMs x blocks_per_sm resident slots in
// an integral number of equal waves, because a block lives for a whole segment.  Pick the segment count that
// minimises  waves x (rows per segment + per-segment overhead).  Grids of less than one wave get segments as short
// as two rows: the step is then a latency chain of (rows + overhead) row-times, not a throughput problem.
static int
choose_seg_rows(int nbx, int ny, int blocks_per_sm)
{
  const int  slots = device_sm_count() * blocks_per_sm;
  const int  overhead_rows = 4;
  long       best_cost = -1;
  int        best_rows = ny;
  const int  max_seg = ny < 4096 ? ny : 4096;
  static const int min_rows = [] { // development switch: shortest segment considered
    const char * e = std::getenv("E2D_MIN_SEG_ROWS");
    const int    v = e ? std::atoi(e) : 2;
    return v > 0 ? v : 2;
  }();
  static const int forced_rows = [] { // development switch: this many rows per segment, whatever the grid
    const char * e = std::getenv("E2D_SEG_ROWS");
    return e ? std::atoi(e) : 0;
  }();
  if (forced_rows > 0)
    return forced_rows < ny ? forced_rows : ny;
  for (int nseg = 1; nseg <= max_seg; ++nseg)
  {
    const int  rows = (ny + nseg - 1) / nseg;
    const int  used = (ny + rows - 1) / rows; // segments actually needed with this row count
    const long blocks = (long)nbx * used;
    const long waves = (blocks + slots - 1) / slots;
    const long cost = waves * (rows + overhead_rows);
    if (best_cost < 0 || cost < best_cost)
    {
      best_cost = cost;
      best_rows = rows;
    }
    if (rows <= min_rows)
      break;
  }
  return best_rows < 1 ? 1 : best_rows;
}

// Tapered segments.  Blocks are handed to the SMs in the order of their segment index, an SM shares its issue slots
// among its resident blocks, and when the queue runs dry every SM finishes what it holds with two, then one block —
// at 65 % and 43 % of its throughput (tools/step_timeline.py).  With uniform segments that drain costs about half a
// block's duration per step (3 % at 8192^2 with 155-row segments).  So the segment length follows the work that is
// left ("guided" scheduling): a block takes 1/guide (1.5) of the fair share of the remaining rows per resident slot,
// never less than min_rows (32; profiles/r3v_seg_taper_sweep.txt: 1.25-1.75 and 32-48 are all within 0.5 % on the large
// grids, 1.0 falls off a cliff on wide ones) — long blocks (few re-traced rows) while there is plenty to do, short ones at
// the end.  That also pays for a grid of a single wave of uniform blocks (4096^2: +8 %): the blocks that share an SM do
// not finish together, so a lone wave drains just the same.  Grids whose fair share of rows per resident slot is below
// 6 minimal segments (3072^2 and smaller: the re-traced rows then cost more than the drain), and grids that would need
// more segments than the table holds, keep uniform segments.
// Returns the number of segments.  (E2D_SEG_TAPER=0 switches it off, 2 forces it on grids of any size — the parity tests
// of the table path; E2D_SEG_GUIDE / E2D_SEG_MIN_ROWS tune it: A/B runs.)
static int
taper_segments(int nbx, int rows, int blocks_per_sm, int uniform_rows, int uniform_nseg, MarchArgs & a)
{
  static const int    enabled = [] { const char * e = std::getenv("E2D_SEG_TAPER"); return e ? std::atoi(e) : 1; }();
  static const double guide = [] { const char * e = std::getenv("E2D_SEG_GUIDE"); return e ? std::atof(e) : 1.5; }();
  static const int    min_rows = [] { const char * e = std::getenv("E2D_SEG_MIN_ROWS"); return e ? std::atoi(e) : 32; }();
  a.seg_tab_n = 0;
  const int slots = device_sm_count() * blocks_per_sm;
  if (!enabled || guide < 1.0 || min_rows < 2)
    return uniform_nseg;
  // worth it once the fair share of rows per resident slot is several minimal segments long — also for a grid of a
  // single wave of uniform blocks: the blocks sharing an SM do not finish together, so that wave drains just the same
  static const double fair_min = [] { const char * e = std::getenv("E2D_SEG_TAPER_FAIR"); return e ? std::atof(e) : 6.0; }();
  (void)uniform_rows;
  if (enabled != 2 && (double)rows * nbx / slots < fair_min * min_rows) // 2: tests force it on small grids
    return uniform_nseg;
  int n = 0, at = 0;
  a.seg_tab[0] = 0;
  while (at < rows)
  {
    if (n == MarchArgs::kSegTabMax)
      return uniform_nseg; // does not fit the table: uniform segments (seg_tab_n stays 0)
    int len = (int)((double)(rows - at) * nbx / (guide * slots));
    if (len < min_rows)
      len = min_rows;
    if (rows - at - len < min_rows / 2) // no sliver at the end
      len = rows - at;
    at += len;
    a.seg_tab[++n] = at;
  }
  a.seg_tab_n = n;
  return n;
}

namespace
{
// one-time, per device, per instantiation: opt in to the dynamic shared memory the kernel needs
template <typename K>
cudaError_t
configure_once(K kernel, size_t smem, std::atomic<unsigned long long> & done_mask)
{
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done_mask.load(std::memory_order_acquire) & bit)
    return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    done_mask.fetch_or(bit, std::memory_order_release);
  return e;
}

template <int SOL, bool FUSE, int LOOP, int MATH, int TYP, bool PEEL = false>
cudaError_t
launch_instance(const dim3 & grid, size_t smem, cudaStream_t st, bool pdl, const MarchArgs & a, const int * d_done,
                const FusedLink & lk, const SoloLoop & so)
{
  static std::atomic<unsigned long long> configured{ 0 };
  auto                                   kernel = k_fused_step<SOL, FUSE, LOOP, MATH, TYP, PEEL>;
  if (cudaError_t e = configure_once(kernel, smem, configured))
    return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kBX);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, a, d_done, lk, so);
}

// mode: 0 plain without the fused CFL reduction, 1 plain with it, 2 peers (multi-GPU loop), 3 solo (single-GPU loop)
template <int SOL, int MATH, int TYP>
cudaError_t
launch_mode(int mode, const dim3 & grid, size_t smem, cudaStream_t st, bool pdl, const MarchArgs & a,
            const int * d_done, const FusedLink & lk, const SoloLoop & so)
{
  switch (mode)
  {
    case 0:
      return launch_instance<SOL, false, 0, MATH, TYP>(grid, smem, st, pdl, a, d_done, lk, so);
    case 1:
      return launch_instance<SOL, true, 0, MATH, TYP>(grid, smem, st, pdl, a, d_done, lk, so);
    case 2:
      return launch_instance<SOL, true, 1, MATH, TYP>(grid, smem, st, pdl, a, d_done, lk, so);
    default:
      return launch_instance<SOL, true, 2, MATH, TYP>(grid, smem, st, pdl, a, d_done, lk, so);
  }
}
} // namespace

cudaError_t
launch_fused_step(const e2d_params & p, const Geom & g, const double * Uin, double * Uout, double dt,
                  const double * d_dt, unsigned long long * d_invdt_bits, const int * d_done, cudaStream_t st,
                  const MarchPeers * peers, FusedLink * link, int j_first, int j_last, bool pdl, const SoloLoop * solo)
{
  MarchArgs a;
  a.Uin = Uin;
  a.Uout = Uout;
  a.isize = g.isize;
  a.jsize = g.jsize;
  a.s = make_settings(p);
  a.c = make_step_consts(a.s);
  a.dt = dt;
  a.d_dt = d_dt;
  a.invdt_bits = d_invdt_bits;
  const bool fast_arith = p.arithmetic == E2D_ARITH_FAST && solver_for(p) == E2D_RIEMANN_HLLC;
  a.rdx_y = fast_arith ? 1.0 / a.s.dx : refined_reciprocal(a.s.dx);
  a.rdy_y = fast_arith ? 1.0 / a.s.dy : refined_reciprocal(a.s.dy);
  const int nbx = (g.nx + (kBX - 4) - 1) / (kBX - 4);
  // rows [j_first, j_last) only (host-streamed step); default: the whole slab
  const int rows = j_last > 0 ? j_last - j_first : g.ny;
  if (j_last > 0)
  {
    if (link || solo || j_first < 2 || j_last > g.jsize - 2 || rows < 1)
      return cudaErrorInvalidValue;
    a.j_first = j_first;
    a.j_last = j_last;
  }
  const int  sol = solver_for(p);
  // `[other] arithmetic=fast` (e2d_fast.cuh) exists for the HLLC solver, i.e. for everything the reference can run
  const bool fastm = p.arithmetic == E2D_ARITH_FAST && sol == E2D_RIEMANN_HLLC;
  a.seg_rows = choose_seg_rows(nbx, rows, march_min_blocks(fastm ? 1 : 0));
  int nseg = (rows + a.seg_rows - 1) / a.seg_rows;
  nseg = taper_segments(nbx, rows, march_min_blocks(fastm ? 1 : 0), a.seg_rows, nseg, a);
  const dim3 grid((unsigned)nbx, (unsigned)nseg, 1);
  const bool fuse = d_invdt_bits != nullptr;
  FusedLink  lk{};
  SoloLoop   so{};
  int        mode = fuse ? 1 : 0;
  if (link)
  {
    if (!fuse || !peers || solo)
      return cudaErrorInvalidValue;
    a.peer_lo = peers->lo;
    a.peer_hi = peers->hi;
    a.peer_lo_jsize = peers->lo_jsize;
    a.peer_hi_jsize = peers->hi_jsize;
    // segments holding interior rows {0,1} / {ny-2, ny-1} (ny >= 2)
    auto seg_of = [&](int row) {
      if (a.seg_tab_n == 0)
        return row / a.seg_rows;
      int k = 0;
      while (k + 1 < a.seg_tab_n && a.seg_tab[k + 1] <= row)
        ++k;
      return k;
    };
    link->n_lo = (unsigned)nbx * (seg_of(0) != seg_of(1) ? 2u : 1u);
    link->n_hi = (unsigned)nbx * (seg_of(g.ny - 2) != seg_of(g.ny - 1) ? 2u : 1u);
    link->n_all = (unsigned)nbx * (unsigned)nseg;
    lk = *link;
    mode = 2;
  }
  if (solo)
  {
    if (!fuse || !solo->st)
      return cudaErrorInvalidValue;
    so = *solo;
    mode = 3;
  }
  // what is known about the deck becomes a template argument of the HLLC strict kernel (MarchThread<.., TYP>)
  const int    typ = !a.c.limited ? 0 : (p.dx == p.dy ? 1 : 2);
  // the peeled march: single-GPU loop, strict HLLC, short uniform segments (k_fused_step<.., PEEL>); E2D_NO_PEEL=1: A/B runs
  static const bool no_peel = std::getenv("E2D_NO_PEEL") != nullptr;
  const bool        peel = mode == 3 && !fastm && sol == E2D_RIEMANN_HLLC && a.seg_tab_n == 0 && a.seg_rows <= 64 && !no_peel &&
                    !E2D_BULK_FETCH;
  const size_t smem = sizeof(MarchSmem<kBX>);
  cudaError_t  launch_err;
  if (sol == E2D_RIEMANN_APPROX)
    launch_err = launch_mode<0, 0, 0>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  else if (sol == E2D_RIEMANN_HLL)
    launch_err = launch_mode<1, 0, 0>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  else if (sol == E2D_RIEMANN_RUSANOV)
    launch_err = launch_mode<3, 0, 0>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  else if (fastm && typ == 1)
    launch_err = launch_mode<2, 1, 1>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  else if (fastm)
    launch_err = launch_mode<2, 1, 0>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  else if (peel && typ == 1)
    launch_err = launch_instance<2, true, 2, 0, 1, true>(grid, smem, st, pdl, a, d_done, lk, so);
  else if (peel && typ == 2)
    launch_err = launch_instance<2, true, 2, 0, 2, true>(grid, smem, st, pdl, a, d_done, lk, so);
  else if (typ == 1)
    launch_err = launch_mode<2, 0, 1>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  else if (typ == 2)
    launch_err = launch_mode<2, 0, 2>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  else
    launch_err = launch_mode<2, 0, 0>(mode, grid, smem, st, pdl, a, d_done, lk, so);
  count_launch();
  return launch_err != cudaSuccess ? launch_err : cudaGetLastError();
}

#if E2D_TIMELINE
} // namespace e2d
extern "C" int
e2d_debug_timeline(unsigned long long * out, int nblocks)
{
  (void)nblocks; // out: [4][kTimelineBlocks = 1024][kTimelineMarks = 8]
  return (int)cudaMemcpyFromSymbol(out, e2d::g_timeline, sizeof(e2d::g_timeline));
}
namespace e2d
{
#endif

cudaError_t
preload_step_kernels()
{
  cudaFuncAttributes fa;
  cudaError_t        e = cudaFuncGetAttributes(&fa, k_reduce_invdt);
#define E2D_PRE(SOL, MATH, TYP)                                                 \
  if (e == cudaSuccess)                                                         \
    e = cudaFuncGetAttributes(&fa, k_fused_step<SOL, true, 1, MATH, TYP>);      \
  if (e == cudaSuccess)                                                         \
    e = cudaFuncGetAttributes(&fa, k_fused_step<SOL, true, 2, MATH, TYP>);      \
  if (e == cudaSuccess)                                                         \
    e = cudaFuncGetAttributes(&fa, k_fused_step<SOL, true, 0, MATH, TYP>);      \
  if (e == cudaSuccess)                                                         \
    e = cudaFuncGetAttributes(&fa, k_fused_step<SOL, false, 0, MATH, TYP>);
  E2D_PRE(0, 0, 0)
  E2D_PRE(1, 0, 0)
  E2D_PRE(3, 0, 0)
  E2D_PRE(2, 0, 0)
  E2D_PRE(2, 0, 1)
  E2D_PRE(2, 0, 2)
  E2D_PRE(2, 1, 0)
  E2D_PRE(2, 1, 1)
#undef E2D_PRE
  if (e == cudaSuccess)
    e = cudaFuncGetAttributes(&fa, k_fused_step<2, true, 2, 0, 1, true>);
  if (e == cudaSuccess)
    e = cudaFuncGetAttributes(&fa, k_fused_step<2, true, 2, 0, 2, true>);
  return e;
}

cudaError_t
launch_loop_begin_step(LoopState * st_dev, double cfl, double tEnd, cudaStream_t st)
{
  k_loop_begin_step<<<1, 1, 0, st>>>(st_dev, cfl, tEnd);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_loop_end_step(LoopState * st_dev, double tEnd, int max_steps, double * dt_hist, long hist_cap,
                     cudaStream_t st)
{
  k_loop_end_step<<<1, 1, 0, st>>>(st_dev, tEnd, max_steps, dt_hist, hist_cap);
  count_launch();
  return cudaGetLastError();
}

cudaError_t
launch_eval(const e2d_params & p, int func, const double * d_in, double * d_out, long n, cudaStream_t st)
{
  k_eval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(make_settings(p), func, d_in, d_out, n);
  count_launch();
  return cudaGetLastError();
}

} // namespace e2d
