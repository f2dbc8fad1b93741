// The fused step as a row-marching ("2.5-D") kernel body.
//
// One thread owns one grid column i and marches along +j through a segment of rows, keeping a
// three-row window of primitive variables in registers.  Per row it
//   A) fetches the west/east primitives of the row from shared memory, limits the slopes, does the
//      MUSCL-Hancock half-step trace to the four faces of its cell, solves the y-face Riemann problem
//      against the YMAX state it kept from the previous row, and publishes its XMAX face state and the
//      next row's primitives;
//   -- one __syncthreads --
//   B) solves the x-face Riemann problem against the XMAX state of its west neighbour, publishes the
//      x flux, completes the conservative update of the PREVIOUS row (whose east x flux and north y
//      flux are now known), stores it, and folds the next step's CFL reduction into the same pass.
//
// Each cell's slopes and trace are computed exactly once, each face's Riemann problem exactly once
// (the reference's flux kernel computes 3 slope sets, 4 traces and 2 solves per cell,
// src/HydroRunFunctors.h:451-644), every conservative value is read from HBM once (+4/(BX-4) column
// halo) and written once.  The update applies  U + Fx(i) - Fx(i+1) + Fy(j) - Fy(j+1)  in the
// order of UpdateFunctor (src/HydroRunFunctors.h:695-713) with fluxes pre-scaled by dt/dx, dt/dy
// (:572-575,:637-640), so the result is bit-identical to the reference's implementation 0.
//
// Shared memory is double-buffered on the row parity, which is what allows a single barrier per row:
//   buffer[r&1].Q    primitives of row r+1      written in A(r)   read in A(r+1)
//   buffer[r&1].XMAX XMAX face states of row r  written in A(r)   read in B(r)
//   buffer[r&1].FX   x fluxes of row r-1        written in B(r-1) read in B(r)
//
// The body is written as a per-thread state machine (init / phaseA / phaseB) so that the test-suite
// can run the very same code on the host, one "thread" after the other with the barrier between the
// phases (tests/host_emulation) — the product only ever runs it on the GPU.
#ifndef E2D_MARCH_CUH
#define E2D_MARCH_CUH

#include "e2d_math.cuh"

namespace e2d
{

struct MarchArgs
{
  const double *       Uin;
  double *             Uout;
  int                  isize, jsize; // slab extent incl. ghosts
  int                  seg_rows;     // interior rows per block segment
  Settings             s;
  double               dt;         // used when d_dt == nullptr
  const double *       d_dt;       // device-resident dt (optional)
  unsigned long long * invdt_bits; // optional: atomicMax target for the next step's CFL reduction
};

template <int BX>
struct MarchSmem
{
  double Q[2][4][BX];
  double XMAX[2][4][BX];
  double FX[2][4][BX];
};

#if defined(__CUDACC__)
#  define E2D_UNROLL _Pragma("unroll")
#else
#  define E2D_UNROLL
#endif

template <int BX, int SOLVER, bool FUSE_DT>
struct MarchThread
{
  // geometry
  int    t, tm, tp; // lane in the block, clamped west / east lanes
  int    i, ic;     // grid column, clamped grid column
  int    j0, j1;    // interior rows [j0, j1) are produced by this block
  bool   store;     // this thread owns an output column
  size_t plane;     // isize * jsize
  double dtdx, dtdy;
  // carried across rows
  double qS[4], qC[4], qN[4]; // primitives of rows r-1, r, r+1
  double uC[4], uN[4];        // conservatives of rows r, r+1
  double xminC[4];            // XMIN face state of row r (A -> B)
  double ymaxP[4];            // YMAX face state of row r-1
  double fyP[4], fyN[4];      // y fluxes at the south faces of rows r-1 and r
  double pend[4];             // U(r-1) + Fx(i, r-1)
  double invdt;

  E2D_HD void
  load_row(const MarchArgs & a, int j, double u[4]) const
  {
    const double * p = a.Uin + (size_t)j * a.isize + ic;
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
      u[v] = p[v * plane];
  }

  E2D_HD void
  to_prim(const MarchArgs & a, const double u[4], double q[4]) const
  {
    compute_primitives_noc(a.s, u[ID], u[IP], u[IU], u[IV], q[ID], q[IP], q[IU], q[IV]);
  }

  // returns false when the block has no rows to produce (uniform over the block)
  E2D_HD bool
  init(const MarchArgs & a, MarchSmem<BX> & sm, int lane, int bx, int seg)
  {
    t = lane;
    tm = t > 0 ? t - 1 : 0;
    tp = t < BX - 1 ? t + 1 : BX - 1;
    i = bx * (BX - 4) + t;
    ic = i < a.isize ? i : a.isize - 1;
    store = (t >= 2) && (t <= BX - 3) && (i >= 2) && (i <= a.isize - 3);
    plane = (size_t)a.isize * a.jsize;
    j0 = 2 + seg * a.seg_rows;
    j1 = j0 + a.seg_rows;
    if (j1 > a.jsize - 2)
      j1 = a.jsize - 2;
    if (j0 >= j1)
      return false;
    const double dt = a.d_dt ? *a.d_dt : a.dt;
    dtdx = dt / a.s.dx; // HydroRun.h:290-291
    dtdy = dt / a.s.dy;
    invdt = 0.0;

    double u[4];
    load_row(a, j0 - 2, u);
    to_prim(a, u, qS);
    load_row(a, j0 - 1, uC);
    to_prim(a, uC, qC);
    load_row(a, j0, uN);
    to_prim(a, uN, qN);
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      sm.Q[(j0 - 2) & 1][v][t] = qC[v]; // primitives of the first traced row, r = j0-1
      ymaxP[v] = 0.0;
      fyP[v] = 0.0;
      fyN[v] = 0.0;
      pend[v] = 0.0;
    }
    return true;
  }

  // r = row being traced, j0-1 <= r <= j1
  E2D_HD void
  phaseA(const MarchArgs & a, MarchSmem<BX> & sm, int r)
  {
    const Settings & s = a.s;
    double           qW[4], qE[4], dqX[4], dqY[4], s0[4], xmax[4], ymin[4], ymax[4];
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      qW[v] = sm.Q[(r - 1) & 1][v][tm];
      qE[v] = sm.Q[(r - 1) & 1][v][tp];
    }
    slopes_dir(s, qC, qE, qW, dqX);
    slopes_dir(s, qC, qN, qS, dqY);
    trace_sources(s, qC, dqX, dqY, s0);
    trace_face<-1>(s, qC, dqX, s0, dtdx, xminC);
    trace_face<+1>(s, qC, dqX, s0, dtdx, xmax);
    trace_face<-1>(s, qC, dqY, s0, dtdy, ymin);
    trace_face<+1>(s, qC, dqY, s0, dtdy, ymax);

    if (r >= j0)
    {
      // south face of row r: left = YMAX of row r-1, right = YMIN of row r, IU<->IV swapped
      // (HydroRunFunctors.h:621-640)
      double f_d, f_e, f_n, f_t;
      riemann<SOLVER>(s, ymaxP[ID], ymaxP[IP], ymaxP[IV], ymaxP[IU], ymin[ID], ymin[IP], ymin[IV], ymin[IU],
                      f_d, f_e, f_n, f_t);
      fyN[ID] = f_d * dtdy;
      fyN[IP] = f_e * dtdy;
      fyN[IU] = f_t * dtdy;
      fyN[IV] = f_n * dtdy;
    }
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      sm.XMAX[r & 1][v][t] = xmax[v];
      sm.Q[r & 1][v][t] = qN[v];
      ymaxP[v] = ymax[v];
    }
  }

  E2D_HD void
  phaseB(const MarchArgs & a, MarchSmem<BX> & sm, int r)
  {
    const Settings & s = a.s;
    double           uP[4];
    const bool       more = r < j1;
    if (more)
      load_row(a, r + 2, uP); // issued early, consumed at the bottom

    double     fx[4] = { 0.0, 0.0, 0.0, 0.0 };
    const bool xface = (r >= j0) && (r < j1);
    if (xface)
    {
      // west face of cell (i, r): left = XMAX of (i-1, r), right = XMIN of (i, r)
      // (HydroRunFunctors.h:559-575)
      double xl[4];
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
        xl[v] = sm.XMAX[r & 1][v][tm];
      double f_d, f_e, f_n, f_t;
      riemann<SOLVER>(s, xl[ID], xl[IP], xl[IU], xl[IV], xminC[ID], xminC[IP], xminC[IU], xminC[IV], f_d, f_e,
                      f_n, f_t);
      fx[ID] = f_d * dtdx;
      fx[IP] = f_e * dtdx;
      fx[IU] = f_n * dtdx;
      fx[IV] = f_t * dtdx;
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
        sm.FX[(r + 1) & 1][v][t] = fx[v];
    }

    if (r >= j0 + 1)
    {
      // complete row r-1: UpdateFunctor order (HydroRunFunctors.h:695-713)
      double un[4];
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
      {
        double x = pend[v];           // U + Fx(i, j)
        x -= sm.FX[r & 1][v][tp];     //   - Fx(i+1, j)
        x += fyP[v];                  //   + Fy(i, j)
        x -= fyN[v];                  //   - Fy(i, j+1)
        un[v] = x;
      }
      if (store)
      {
        double * po = a.Uout + (size_t)(r - 1) * a.isize + i;
        E2D_UNROLL
        for (int v = 0; v < 4; ++v)
          po[v * plane] = un[v];
        if (FUSE_DT)
          invdt = fmax(invdt, cfl_inv_dt(s, un[ID], un[IP], un[IU], un[IV]));
      }
    }

    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      if (xface)
        pend[v] = uC[v] + fx[v];
      fyP[v] = fyN[v];
    }
    if (more)
    {
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
      {
        qS[v] = qC[v];
        qC[v] = qN[v];
        uC[v] = uN[v];
        uN[v] = uP[v];
      }
      to_prim(a, uP, qN);
    }
  }
};

} // namespace e2d

#endif // E2D_MARCH_CUH
