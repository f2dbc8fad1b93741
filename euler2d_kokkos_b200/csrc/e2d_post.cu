// Post-processing and output kernels: the Sedov radial profile and the interior gather for snapshots.
//
// Radial profile — replaces ComputeRadialProfileFunctor (src/ComputeRadialProfileFunctor.h:86-167): every cell of
// the array, ghost cells included (:106), is binned by the distance of its centre from the box centre,
//   bin = (int)(distance / max_radial_distance * nbins);  histo[bin] += 1;  profile[bin] += rho     (:151-165)
// The reference accumulates with atomics (order unspecified, so its sums are reproducible to round-off only) and
// indexes past the end of its bin arrays for the corner ghost cells (distance > max_radial_distance is never
// tested); here those samples are dropped and the sums are DETERMINISTIC: one thread per grid column walks its
// rows in order into a private strip of bins (coalesced across the columns of a warp: part[bin][seg][i]), then
// one block per bin folds the strips with a fixed-shape tree.  Same formulas, operation for operation, for x, y,
// distance and the bin (IEEE sqrt and division, no FMA contraction), so the integer histogram is exact.
#include "e2d_internal.h"

namespace e2d
{

namespace
{

__global__ void __launch_bounds__(128)
k_radial_partial(Geom g, RadialArgs a, const double * __restrict__ U, double * __restrict__ part_sum,
                 int * __restrict__ part_cnt)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.isize)
    return;
  const int seg = blockIdx.y, nseg = gridDim.y;
  const int rows = (a.j_hi - a.j_lo + nseg - 1) / nseg;
  const int j0 = a.j_lo + seg * rows;
  const int j1 = min(j0 + rows, a.j_hi);
  const double x = a.xmin + a.dx / 2 + (i - a.gw) * a.dx; // :148
  const double ddx = (x - a.cx) * (x - a.cx);
  for (int j = j0; j < j1; ++j)
  {
    const double y = a.ymin + a.dy / 2 + (j + g.j_off - a.gw) * a.dy;     // :149
    const double distance = sqrt(ddx + (y - a.cy) * (y - a.cy));          // :151-152
    const int    bin = (int)(distance / a.rmax * a.nbins);                // :155
    if (bin < 0 || bin >= a.nbins)
      continue; // out of bounds in the reference
    const size_t o = ((size_t)bin * nseg + seg) * g.isize + i;
    part_cnt[o] += 1;                                                     // :158
    part_sum[o] += U[(size_t)i + (size_t)g.isize * j];                    // :162 (ID plane)
  }
}

// one block per bin: n = nseg * isize contiguous partials, thread t takes t, t + 256, ... in order, then a
// fixed-shape tree in shared memory (the launch shape is a constant, so the summation order is too)
__global__ void __launch_bounds__(256)
k_radial_fold(int n, const double * __restrict__ part_sum, const int * __restrict__ part_cnt,
              double * __restrict__ sums, int * __restrict__ counts)
{
  __shared__ double ss[256];
  __shared__ int    sc[256];
  const int         bin = blockIdx.x;
  const double *    ps = part_sum + (size_t)bin * n;
  const int *       pc = part_cnt + (size_t)bin * n;
  double            s = 0.0;
  int               c = 0;
  for (int k = threadIdx.x; k < n; k += 256)
  {
    s += ps[k];
    c += pc[k];
  }
  ss[threadIdx.x] = s;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1)
  {
    if (threadIdx.x < w)
    {
      ss[threadIdx.x] += ss[threadIdx.x + w];
      sc[threadIdx.x] += sc[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    sums[bin] = ss[0];
    counts[bin] = sc[0];
  }
}

// interior cells of one variable plane, ghosts stripped, i fastest (the order HydroRun::saveVTK writes them,
// src/HydroRun.h:573-598) -> a dense nx x rows block; one block per row, coalesced both ways
__global__ void __launch_bounds__(256)
k_gather_interior(Geom g, const double * __restrict__ U, double * __restrict__ out, int j_lo, int n_rows)
{
  const size_t plane = (size_t)g.isize * g.jsize;
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x)
  {
    const int j = j_lo + r;
    for (int v = 0; v < 4; ++v)
    {
      const double * src = U + v * plane + (size_t)j * g.isize + 2;
      double *       dst = out + ((size_t)v * n_rows + r) * g.nx;
      for (int i = threadIdx.x; i < g.nx; i += blockDim.x)
        dst[i] = src[i];
    }
  }
}

} // namespace

RadialArgs
make_radial_args(const e2d_params & p, int nbins, int j_lo, int j_hi)
{
  RadialArgs a;
  a.xmin = p.xmin;
  a.ymin = p.ymin;
  a.dx = p.dx;
  a.dy = p.dy;
  a.cx = (p.xmin + p.xmax) / 2; // :95-96
  a.cy = (p.ymin + p.ymax) / 2;
  const double Dx = (p.xmax - p.xmin) / 2, Dy = (p.ymax - p.ymin) / 2; // :99-100
  a.rmax = sqrt(Dx * Dx + Dy * Dy);                                    // :101 (host sqrt, like the reference)
  a.nbins = nbins;
  a.gw = p.ghostWidth;
  a.j_lo = j_lo;
  a.j_hi = j_hi;
  return a;
}

int
radial_segments(const Geom & g, int nbins, int rows)
{
  // enough column strips to fill the GPU, within ~64 MB of partials
  long nseg = (148L * 8 * 128 + g.isize - 1) / g.isize;
  const long cap = (64L << 20) / ((long)nbins * g.isize * 12);
  if (nseg > cap)
    nseg = cap;
  if (nseg > rows)
    nseg = rows;
  return nseg < 1 ? 1 : (int)nseg;
}

cudaError_t
launch_radial_profile(const Geom & g, const RadialArgs & a, const double * U, int nseg, double * part_sum,
                      int * part_cnt, double * d_sums, int * d_counts, cudaStream_t st)
{
  const size_t n = (size_t)a.nbins * nseg * g.isize;
  cudaError_t  e = cudaMemsetAsync(part_sum, 0, n * sizeof(double), st);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(part_cnt, 0, n * sizeof(int), st);
  if (e != cudaSuccess)
    return e;
  k_radial_partial<<<dim3((g.isize + 127) / 128, nseg), 128, 0, st>>>(g, a, U, part_sum, part_cnt);
  k_radial_fold<<<a.nbins, 256, 0, st>>>(nseg * g.isize, part_sum, part_cnt, d_sums, d_counts);
  count_launch(2);
  return cudaGetLastError();
}

cudaError_t
launch_gather_interior(const Geom & g, const double * U, double * out, int j_lo, int n_rows, cudaStream_t st)
{
  if (n_rows <= 0)
    return cudaSuccess;
  const int blocks = n_rows < 148 * 8 ? n_rows : 148 * 8;
  k_gather_interior<<<blocks, 256, 0, st>>>(g, U, out, j_lo, n_rows);
  count_launch();
  return cudaGetLastError();
}

} // namespace e2d
