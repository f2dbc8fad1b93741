// Boundary fill shared by k_make_boundaries (e2d_kernels.cu) and the slab loop's boundary kernel (e2d_slab.cu).
//
// The four MakeBoundariesFunctor<face> launches of HydroRun::make_boundaries
// (src/HydroRunFunctors.h:1832-2030, src/HydroRun.h:390-399) as ONE pass.  The reference runs XMIN, XMAX over all
// rows and then YMIN, YMAX over all columns, so a corner ghost ends up as  (U(i0, j0) * sign_x) * sign_y.  Every
// source cell (i0, j0) is an interior cell, which no pass writes, so each ghost cell can be produced independently
// by composing the two index maps — same values, same signs (multiplying by +-1.0 is exact), no ordering between
// threads.
#ifndef E2D_BC_CUH
#define E2D_BC_CUH

#include "e2d_internal.h"

namespace e2d
{

struct BcArgs
{
  int bc_xmin, bc_xmax, bc_ymin, bc_ymax;
  int faces;
};

__device__ __forceinline__ int
bc_source_lo(int bc, int k, int n, double & sign, bool is_normal)
{ // ghost index k in {0,1}; :1893-1906 / :1968-1981
  if (bc == E2D_BC_DIRICHLET)
  {
    if (is_normal)
      sign = -1.0;
    return 3 - k;
  }
  if (bc == E2D_BC_NEUMANN)
    return 2;
  return n + k; // periodic
}

__device__ __forceinline__ int
bc_source_hi(int bc, int k, int n, double & sign, bool is_normal)
{ // ghost index k in {n+2, n+3}; :1931-1944 / :2006-2019
  if (bc == E2D_BC_DIRICHLET)
  {
    if (is_normal)
      sign = -1.0;
    return 2 * n + 3 - k;
  }
  if (bc == E2D_BC_NEUMANN)
    return n + 1;
  return k - n; // periodic
}

// Source of ghost cell (i, j): the interior cell (i0, j0) it copies, whether it lies in an x / y ghost strip of an
// active face, and whether the normal momentum changes sign there (reflecting wall).  (i, j) interior: itself.
__device__ __forceinline__ void
bc_map(const Geom & g, const BcArgs & a, int i, int j, int & i0, int & j0, bool & in_x, bool & in_y, bool & flip_u,
       bool & flip_v)
{
  const int nx = g.nx, ny = g.ny;
  in_x = (i < 2 && (a.faces & 1)) || (i >= nx + 2 && (a.faces & 2));
  in_y = (j < 2 && (a.faces & E2D_FACES_YMIN)) || (j >= ny + 2 && (a.faces & E2D_FACES_YMAX));
  i0 = i;
  j0 = j;
  double sx = 1.0, sy = 1.0;
  if (in_x)
    i0 = (i < 2) ? bc_source_lo(a.bc_xmin, i, nx, sx, true) : bc_source_hi(a.bc_xmax, i, nx, sx, true);
  if (in_y)
    j0 = (j < 2) ? bc_source_lo(a.bc_ymin, j, ny, sy, true) : bc_source_hi(a.bc_ymax, j, ny, sy, true);
  flip_u = sx < 0.0;
  flip_v = sy < 0.0;
}

// value of ghost cell (i, j), variable v, from array U:  (U(i0, j0) * sign_x) * sign_y, the composition the
// reference's XMIN, XMAX -> YMIN, YMAX sequence produces (multiplying by +-1.0 is exact)
__device__ __forceinline__ double
bc_value(const double * __restrict__ U, size_t src, int v, bool in_x, bool in_y, bool flip_u, bool flip_v)
{
  // .cg: the source may be a halo row that a peer GPU stored while this kernel was already running, or a cell another
  // thread of this block has just written
  double val = __ldcg(U + src);
  if (in_x)
    val = val * ((v == IU && flip_u) ? -1.0 : 1.0);
  if (in_y)
    val = val * ((v == IV && flip_v) ? -1.0 : 1.0);
  return val;
}

// ghost cell number k of the slab (k < 4*isize: the y-ghost rows, full width; then the x-ghost columns)
__device__ __forceinline__ void
bc_fill_cell(const Geom & g, const BcArgs & a, double * __restrict__ U, int k)
{
  const int    nx = g.nx, ny = g.ny;
  const size_t plane = (size_t)g.isize * g.jsize;
  const int    n_y = 4 * g.isize; // y-ghost rows, full width (i fastest)
  const int    n_x = 4 * g.jsize; // x-ghost columns

  int i, j;
  if (k < n_y)
  {
    const int gsel = k / g.isize;
    i = k - gsel * g.isize;
    j = gsel < 2 ? gsel : ny + gsel;
    if (!(a.faces & (gsel < 2 ? E2D_FACES_YMIN : E2D_FACES_YMAX)))
      return;
  }
  else if (k < n_y + n_x)
  {
    const int kk = k - n_y;
    j = kk >> 2;
    const int gsel = kk & 3;
    i = gsel < 2 ? gsel : nx + gsel;
    if (!(a.faces & (gsel < 2 ? 1 : 2)))
      return;
    // rows that an active y face rewrites are produced by the first branch
    if ((j < 2 && (a.faces & E2D_FACES_YMIN)) || (j >= ny + 2 && (a.faces & E2D_FACES_YMAX)))
      return;
  }
  else
    return;

  int  i0, j0;
  bool in_x, in_y, flip_u, flip_v;
  bc_map(g, a, i, j, i0, j0, in_x, in_y, flip_u, flip_v);
#pragma unroll
  for (int v = 0; v < 4; ++v)
    U[(size_t)i + (size_t)g.isize * j + v * plane] =
      bc_value(U, (size_t)i0 + (size_t)g.isize * j0 + v * plane, v, in_x, in_y, flip_u, flip_v);
}

inline BcArgs
make_bc_args(const e2d_params & p, int faces)
{
  BcArgs a;
  a.bc_xmin = p.boundary_type_xmin;
  a.bc_xmax = p.boundary_type_xmax;
  a.bc_ymin = p.boundary_type_ymin;
  a.bc_ymax = p.boundary_type_ymax;
  a.faces = faces;
  return a;
}

} // namespace e2d

#endif // E2D_BC_CUH
