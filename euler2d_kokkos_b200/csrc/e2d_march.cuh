// The fused step as a row-marching ("2.5-D") kernel body.
//
// One thread owns one grid column i and marches along +j through a segment of rows.  Per row r:
//   A) reads the primitives of the five-point stencil from a three-row ring in shared memory,
//      limits the slopes, does the MUSCL-Hancock half-step trace to the four faces of cell (i, r),
//      publishes the XMAX face state (for the east neighbour) and parks the YMAX face state (for
//      its own next row);
//   -- one __syncthreads --
//   B) reads its ring rows and at once issues the asynchronous fetch (cp.async) of row r+3 into the
//      slot just read; solves the x-face Riemann problem (west face of (i, r)) and the y-face one
//      (south face) back to back in one straight-line block, so the two solves, the CFL integrand of
//      the row completed one phase earlier and the primitive conversion of row r+2 overlap in the
//      FP64 pipe; publishes the x flux, completes the conservative update of row r-1 (its east x
//      flux and north y flux are now known), stores it, folds the next step's CFL reduction in, and
//      writes the primitives of row r+2 into the ring.
//
// Each cell's slopes and trace are computed exactly once, each face's Riemann problem exactly once
// (the reference's flux kernel computes 3 slope sets, 4 traces and 2 solves per cell,
// src/HydroRunFunctors.h:451-644), every conservative value is read from HBM once (+4/(BX-4) column
// halo) and written once.  The update applies  U + Fx(i) - Fx(i+1) + Fy(j) - Fy(j+1)  in the
// order of UpdateFunctor (src/HydroRunFunctors.h:695-713) with fluxes pre-scaled by dt/dx, dt/dy
// (:572-575,:637-640), so the result is bit-identical to the reference's implementation 0.
//
// Shared memory (ring slots are row % 3 or row & 1; one barrier per row suffices; each row of
// states is four planes of doubles in the strict kernel, two planes of 16-byte pairs in the fast one):
//   Q[r%3]       primitives of rows r-1, r, r+1      written in B(r-2)   read in A(r-1..r+1)
//   RY[r%3]      refined 1/rho of the same rows      own column only
//   U[r%3]       conservatives of rows r, r+1, r+2   own column only     row r+3 is fetched by cp.async issued in
//                                                                        B(r) into the slot of row r (just read); it is
//                                                                        read in B(r+1) (-> Q) and B(r+3)
//   XMAX[r&1]    XMAX face states of row r           written in A(r)     read in B(r) by the east lane
//   YMAX[r&1]    YMAX face states of row r           own column only     written A(r), read B(r+1)
//   FX[(r+1)&1]  x fluxes of row r                   written in B(r)     read in B(r+1) by the west lane
// Own-column rows live in shared memory rather than registers: cp.async needs a shared-memory target anyway,
// and it keeps the register file for the solver's temporaries (ncu: 3 blocks/SM at <=168 registers beat 4
// blocks/SM at 128, profiles/r1_fused_step_variants.txt).
//
// All work is unconditional (rows outside [j0, j1) compute on valid-but-unused data); only the side
// effects are predicated.  The arithmetic is the lean-but-exact form of e2d_lean.cuh; when a fast-path
// guard fails anywhere in a phase the phase is recomputed with plain IEEE operators.
//
// The body is written as a per-thread state machine (init / phaseA / phaseB) so that the test-suite
// can run the very same code on the host, one "thread" after the other with the barrier between the
// phases (tests/host_emulation) — the product only ever runs it on the GPU.
#ifndef E2D_MARCH_CUH
#define E2D_MARCH_CUH

#include "e2d_lean.cuh"
#include "e2d_fast.cuh"

namespace e2d
{

struct MarchArgs
{
  const double *       Uin;
  double *             Uout;
  int                  isize, jsize; // slab extent incl. ghosts
  int                  seg_rows;     // interior rows per block segment (uniform segments: seg_tab_n == 0)
  // Tapered segments (launch_fused_step, e2d_kernels.cu): segment k produces the rows [seg_tab[k], seg_tab[k+1]) counted
  // from j_first, long segments first and ever shorter ones towards the end of the block queue, so that the SMs run
  // dry together instead of each finishing a long block alone.
  static constexpr int kSegTabMax = 191;
  int                  seg_tab_n = 0;
  int                  seg_tab[kSegTabMax + 1];
  int                  j_first = 2;  // rows [j_first, j_last) are produced (j_last <= 0: jsize - 2); a sub-range lets the
  int                  j_last = 0;   // host-streamed step (e2d_capi.cu) advance chunk by chunk as the rows arrive
  Settings             s;
  StepConsts           c;
  double               rdx_y = 0.0, rdy_y = 0.0; // reciprocals of dx, dy for the CFL integrand (MarchThread::recip_dx)
  double               dt;         // used when d_dt == nullptr
  const double *       d_dt;       // device-resident dt (optional)
  unsigned long long * invdt_bits; // optional: atomicMax target for the next step's CFL reduction
  // multi-GPU (e2d_slab.cu): the same-parity OUTPUT arrays of the lower / upper y-neighbour as peer pointers.  The
  // blocks that produce this slab's first / last two interior rows copy them into the neighbour's ghost rows
  // (NVLink stores) as soon as their segment is done, so the halo exchange of the NEXT step rides on this kernel —
  // overlapped with the interior segments — instead of following it (k_fused_step, e2d_kernels.cu).
  double * peer_lo = nullptr;
  double * peer_hi = nullptr;
  int      peer_lo_jsize = 0, peer_hi_jsize = 0;
};

#if defined(__CUDACC__)
#  define E2D_UNROLL _Pragma("unroll")
#else
#  define E2D_UNROLL
#endif

// The four variables of a cell are kept as two 16-byte pairs, (rho, p|E) and (u|mx, v|my), each pair array indexed
// by the lane: a row of states is read and written with 128-bit shared-memory accesses (2 per state instead of 4;
// consecutive lanes are 16 bytes apart, so every quarter-warp covers all 32 banks — conflict-free, also for the
// west / east neighbour's lane).
struct alignas(16) Pair
{
  double a, b;
};

// E2D_BULK_FETCH = 1 (experiment, profiles/r2p_bulk_fetch_ab.txt): the next row of conservative states is fetched by
// ONE elected thread per block with four bulk asynchronous copies (cp.async.bulk, one 1 KB run per variable plane)
// that complete on an mbarrier, instead of four 8-byte cp.async (LDGSTS) per thread.  A bulk copy lands contiguous
// bytes, so the U ring becomes planar (UB) and four slots deep: the slot overwritten by the copy of row r+3 was last
// read in phase B(r-1), which every thread has left when it passes the barrier of row r.
#ifndef E2D_BULK_FETCH
#  define E2D_BULK_FETCH 0
#endif
#define E2D_BULK (E2D_BULK_FETCH && E2D_LEAN_DEVICE)

// interior rows [j0, j1) of segment `seg`; false when the segment is empty
E2D_HD bool
segment_rows(const MarchArgs & a, int seg, int & j0, int & j1)
{
  const int j_end = a.j_last > 0 ? a.j_last : a.jsize - 2;
  if (a.seg_tab_n > 0)
  {
    j0 = a.j_first + a.seg_tab[seg];
    j1 = a.j_first + a.seg_tab[seg + 1];
  }
  else
  {
    j0 = a.j_first + seg * a.seg_rows;
    j1 = j0 + a.seg_rows;
  }
  if (j1 > j_end)
    j1 = j_end;
  return j0 < j1;
}

template <int BX>
struct MarchSmem
{
  Pair   Q[3][2][BX];
  double RY[3][BX];
  Pair   U[3][2][BX];
  Pair   XMAX[2][2][BX];
  Pair   YMAX[2][2][BX];
  Pair   FX[2][2][BX];
#if E2D_BULK_FETCH
  alignas(16) double             UB[4][4][BX]; // planar U ring of the bulk-fetch variant (row & 3)
  alignas(8) unsigned long long mbar[4];      // one transaction barrier per UB slot
#endif
};

// PACKED = false: the same storage read as four planes double[4][BX] with 64-bit accesses (the strict kernel, whose
// register allocation is 3 % better off without the pairing: profiles/r1_fused_step_variants.txt).
template <bool PACKED, int BX>
E2D_HD double *
slot_of(Pair (&row)[2][BX], int v, int t)
{
  if (PACKED)
    return (v & 1) ? &row[v >> 1][t].b : &row[v >> 1][t].a;
  return reinterpret_cast<double *>(&row[0][0]) + v * BX + t;
}

template <bool PACKED, int BX>
E2D_HD void
ld4(const Pair (&row)[2][BX], int t, double v[4])
{
  if (PACKED)
  {
    const Pair x = row[0][t], y = row[1][t];
    v[0] = x.a;
    v[1] = x.b;
    v[2] = y.a;
    v[3] = y.b;
  }
  else
  {
    const double * r = reinterpret_cast<const double *>(&row[0][0]) + t;
    E2D_UNROLL
    for (int k = 0; k < 4; ++k)
      v[k] = r[k * BX];
  }
}

template <bool PACKED, int BX>
E2D_HD void
st4(Pair (&row)[2][BX], int t, const double v[4])
{
  if (PACKED)
  {
    Pair x, y;
    x.a = v[0];
    x.b = v[1];
    y.a = v[2];
    y.b = v[3];
    row[0][t] = x;
    row[1][t] = y;
  }
  else
  {
    double * r = reinterpret_cast<double *>(&row[0][0]) + t;
    E2D_UNROLL
    for (int k = 0; k < 4; ++k)
      r[k * BX] = v[k];
  }
}


// MATH: 0 = strict (bit-identical to the reference's x86 arithmetic), 1 = fast (e2d_fast.cuh: explicit FMAs and
// reciprocal-multiply division, within north_star's 1e-12 of the reference; `[other] arithmetic=fast`).  With
// MATH == 1 the RY ring holds the fast reciprocal of the density, FX / fyP carry UNSCALED fluxes (the update applies
// dt/dx, dt/dy inside its fma chain) and the solver is the fast HLLC when SOLVER == 2.
// TYP: what is known about the deck at compile time.  1: limited slopes (slope_type 1 or 2) on square cells (dx == dy
// bit for bit, hence dt/dx == dt/dy) — four of the reference's five decks;  2: limited slopes, dx != dy (its
// shocked_bubble deck: 0.445f/445 and 0.089f/89 differ in the last bits);  0: nothing — slope_type and dx == dy are
// tested at run time, which costs the strict kernel ~40 issue slots per row (predicated multiplies, selects against
// zero slopes, two DSETPs).
template <int BX, int SOLVER, bool FUSE_DT, int MATH = 0, int TYP = 0>
struct MarchThread
{
#ifndef E2D_STRICT_PACK
#  define E2D_STRICT_PACK 1
#endif
  static constexpr bool PACK = (MATH == 1) || E2D_STRICT_PACK; // shared-memory rows as 16-byte pairs
  static constexpr bool LIMITED = (TYP != 0);
  static constexpr bool SQUARE = (TYP == 1);
  static constexpr bool UNFL = E2D_WINDOW_GUARDS != 0; // face densities left unfloored (e2d_lean.cuh)
  // geometry
  int    t, tm, tp; // lane in the block, clamped west / east lanes
  int    i, ic;     // grid column, clamped grid column
  int    j0, j1;    // interior rows [j0, j1) are produced by this block
  bool   store;     // this thread owns an output column
  size_t plane;     // isize * jsize
  double dtdx, dtdy;
  double hdtdx, hdtdy; // MATH == 1: dt/dx/2, dt/dy/2
  int    m3;       // (row being traced) % 3
  // carried across the barrier / rows
  double xmin[4], ymin[4]; // XMIN / YMIN face states of row r (A -> B)
  double fyP[4];           // y flux at the south face of row r-1
  double pend[4];          // U(r-1) + Fx(i, r-1)
  double unD[4];           // updated state of the row completed by the previous phase B (CFL integrand deferred)
  double invdt;

  // dx, dy with their reciprocals for the CFL integrand: kernel arguments (constant bank), no registers.  Strict: the
  // refined reciprocal of the division sequence (MarchArgs::rdx_y, computed once on the device); fast: 1.0 / dx.
  E2D_HD static Recip
  recip_dx(const MarchArgs & a)
  {
    Recip r;
    r.d = a.s.dx;
    r.y = a.rdx_y;
    return r;
  }
  E2D_HD static Recip
  recip_dy(const MarchArgs & a)
  {
    Recip r;
    r.d = SQUARE ? a.s.dx : a.s.dy;
    r.y = SQUARE ? a.rdx_y : a.rdy_y;
    return r;
  }

  E2D_HD void
  load_row(const MarchArgs & a, int j, double u[4]) const
  {
    // 32-bit in-plane offset (isize*jsize < 2^31 is checked at the ABI); the plane stride is block-uniform
    const double * p = a.Uin + (j * a.isize + ic);
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
      u[v] = p[v * plane];
  }

  // asynchronous fetch of this thread's cell of row j into U ring slot `slot` (cp.async: no registers are
  // held while the load is in flight, so the compiler cannot sink it towards its consumer)
  E2D_HD void
  prefetch_row(const MarchArgs & a, MarchSmem<BX> & sm, int j, int slot) const
  {
    const double * p = a.Uin + (j * a.isize + ic);
#if E2D_LEAN_DEVICE
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(slot_of<PACK>(sm.U[slot], v, t));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(p + v * plane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#else
    double u[4];
    for (int v = 0; v < 4; ++v)
      u[v] = p[v * plane];
    st4<PACK>(sm.U[slot], t, u);
#endif
  }

  E2D_HD void
  wait_prefetch() const
  {
#if E2D_LEAN_DEVICE
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  }

#if E2D_BULK
  // ---- bulk-fetch variant (strict arithmetic only) ----
  __device__ static unsigned
  smem_u32(const void * p)
  {
    return (unsigned)__cvta_generic_to_shared(p);
  }
  // row j of this block's BX columns -> UB[j & 3], by the calling (elected) thread: four bulk copies, one barrier
  __device__ void
  bulk_fetch_row(const MarchArgs & a, MarchSmem<BX> & sm, int j_src, int j) const
  { // j: the row the loop believes it fetches (ring slot, barrier phase); j_src: the same, clamped to the array
    const int      i0 = i - t; // first column of the block
    int            nvalid = a.isize - i0;
    nvalid = nvalid < BX ? nvalid : BX;
    const unsigned bytes = (unsigned)nvalid * 8u;
    const unsigned bar = smem_u32(&sm.mbar[j & 3]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * bytes) : "memory");
    const double * src = a.Uin + (j_src * a.isize + i0);
#pragma unroll
    for (int v = 0; v < 4; ++v)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(&sm.UB[j & 3][v][0])),
                   "l"(src + v * plane), "r"(bytes), "r"(bar)
                   : "memory");
  }
  // wait until the bulk copy of row j has landed; rows are fetched in order starting with row j0 + 1
  __device__ void
  bulk_wait_row(MarchSmem<BX> & sm, int j) const
  {
    const unsigned bar = smem_u32(&sm.mbar[j & 3]);
    const unsigned parity = (unsigned)(((j - (j0 + 1)) >> 2) & 1);
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
  }
  __device__ static void
  ldb4(const double (&row)[4][BX], int t, double v[4])
  {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      v[k] = row[k][t];
  }
  __device__ static void
  stb4(double (&row)[4][BX], int t, const double v[4])
  {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      row[k][t] = v[k];
  }
#endif

  // conservative -> primitive of one cell of a fetched row, into ring slot `slot`
  E2D_HD void
  convert_into(const MarchArgs & a, MarchSmem<BX> & sm, const double u[4], int slot) const
  {
    double q[4];
    Recip  rd;
    if (MATH == 1)
      fast::prim(a.s, a.c, u, q, rd.y);
    else
    {
      bool ok = a.c.lean_ok != 0;
      prim_lean<true>(a.s, a.c, u, q, rd, ok);
      if (!ok)
      {
        Recip unused;
        prim_lean<false>(a.s, a.c, u, q, unused, ok);
      }
    }
    st4<PACK>(sm.Q[slot], t, q);
    sm.RY[slot][t] = rd.y;
  }

  // returns false when the block has no rows to produce (uniform over the block)
  E2D_HD bool
  init(const MarchArgs & a, MarchSmem<BX> & sm, int lane, int bx, int seg, double dt)
  {
    t = lane;
    tm = t > 0 ? t - 1 : 0;
    tp = t < BX - 1 ? t + 1 : BX - 1;
    i = bx * (BX - 4) + t;
    ic = i < a.isize ? i : a.isize - 1;
    store = (t >= 2) && (t <= BX - 3) && (i >= 2) && (i <= a.isize - 3);
    plane = (size_t)a.isize * a.jsize;
    if (!segment_rows(a, seg, j0, j1))
      return false;
    dtdx = dt / a.s.dx; // HydroRun.h:290-291
    dtdy = dt / a.s.dy;
    hdtdx = 0.5 * dtdx;
    hdtdy = 0.5 * dtdy;
    invdt = 0.0;
    m3 = (j0 - 1) % 3;

    // rows j0-2, j0-1, j0 -> primitive ring; rows j0-1, j0 -> conservative ring
    double u[4];
    load_row(a, j0 - 2, u);
    convert_into(a, sm, u, (j0 - 2) % 3);
    load_row(a, j0 - 1, u);
    convert_into(a, sm, u, (j0 - 1) % 3);
#if E2D_BULK
    const bool bulk = (MATH == 0);
    if (bulk)
      stb4(sm.UB[(j0 - 1) & 3], t, u);
    else
#endif
      st4<PACK>(sm.U[(j0 - 1) % 3], t, u);
    load_row(a, j0, u);
    convert_into(a, sm, u, j0 % 3);
#if E2D_BULK
    if (bulk)
    {
      stb4(sm.UB[j0 & 3], t, u);
      // columns past the end of the array are never touched by a bulk copy: give them a valid state in the other slots
      if (i >= a.isize)
      {
        stb4(sm.UB[(j0 + 1) & 3], t, u);
        stb4(sm.UB[(j0 + 2) & 3], t, u);
      }
      if (t == 0)
      {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.mbar[k])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
    }
    else
#endif
      st4<PACK>(sm.U[j0 % 3], t, u);
    // The first phase B (r = j0-1) solves the south face of row j0-1 and "completes" row j0-2; neither result is used,
    // but the completed state goes through the deferred CFL integrand of the next phase B, whose guards would send
    // the whole phase down the plain-operator path if it were not a healthy state (once per segment: measured as
    // 0.7 row-times).  So: the face's left state is row j0-1's own primitive state (a flux of physical size), and the
    // pending sum starts from row j0's state instead of zero.
    {
      double qb[4];
      ld4<PACK>(sm.Q[(j0 - 1) % 3], t, qb);
      st4<PACK>(sm.YMAX[(j0 - 2) & 1], t, qb);
    }
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      fyP[v] = 0.0;
      pend[v] = u[v];
      unD[v] = u[v]; // any valid state
    }
    st4<PACK>(sm.FX[(j0 - 1) & 1], t, fyP); // zeros: read (and unused) by the first phase B
    st4<PACK>(sm.FX[j0 & 1], t, fyP);       // zeros: what phase B(j0) reads when the first phase B is the peeled one (phaseB<1>)
    // row j0+1, read as "row r+2" by the first phase B (r = j0-1)
#if E2D_BULK
    if (bulk)
    { // by the elected thread, once the initialised barriers are visible to everybody
      __syncthreads(); // (uniform: every thread of an active block gets here)
      if (t == 0)
      {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulk_fetch_row(a, sm, j0 + 1 < a.jsize ? j0 + 1 : a.jsize - 1, j0 + 1);
      }
      return true;
    }
#endif
    prefetch_row(a, sm, (j0 + 1 < a.jsize) ? j0 + 1 : a.jsize - 1, (j0 + 1) % 3);
    return true;
  }

  // r = row being traced, j0-1 <= r <= j1
  E2D_HD void
  phaseA(const MarchArgs & a, MarchSmem<BX> & sm, int r)
  {
    const Settings & s = a.s;
    const int        sC = m3, sS = (m3 == 0) ? 2 : m3 - 1, sN = (m3 == 2) ? 0 : m3 + 1;
    double           qC[4], qW[4], qE[4], qS[4], qN[4], dqX[4], dqY[4], s0[4], xmax[4], ymax[4];
    (void)r;
    ld4<PACK>(sm.Q[sC], t, qC);
    ld4<PACK>(sm.Q[sC], tm, qW);
    ld4<PACK>(sm.Q[sC], tp, qE);
    ld4<PACK>(sm.Q[sS], t, qS);
    ld4<PACK>(sm.Q[sN], t, qN);
    Recip rd;
    rd.d = qC[ID];
    rd.y = sm.RY[sC][t];

    // slope_unsplit_hydro_2d (src/HydroBaseFunctor.h:473-516): slope_type outside {1,2} -> zero slopes
    const bool limited = LIMITED || a.c.limited != 0;
    if (MATH == 1)
    {
      const double st0 = limited ? s.slope_type : 0.0;
      fast::slopes(st0, qC, qE, qW, dqX);
      fast::slopes(st0, qC, qN, qS, dqY);
      fast::trace<SQUARE>(s, qC, rd.y, dqX, dqY, hdtdx, SQUARE ? hdtdx : hdtdy, xmin, xmax, ymin, ymax);
    }
    else
    {
      slopes_lean<LIMITED>(s.slope_type, limited, qC, qE, qW, dqX);
      slopes_lean<LIMITED>(s.slope_type, limited, qC, qN, qS, dqY);

      bool ok = a.c.lean_ok != 0;
      trace_sources_lean<true>(s, qC, rd, dqX, dqY, s0, ok);
      if (!ok)
        trace_sources_lean<false>(s, qC, rd, dqX, dqY, s0, ok);
      trace_faces_lean<TYP, UNFL>(s, qC, dqX, dqY, s0, dtdx, SQUARE ? dtdx : dtdy, xmin, xmax, ymin, ymax);
    }

    st4<PACK>(sm.XMAX[r & 1], t, xmax);
    st4<PACK>(sm.YMAX[r & 1], t, ymax);
  }

  // Everything of phase B that is arithmetic: the x-face and y-face solves, the update of row r-1, and — advanced
  // in lock step (e2d_lean.cuh) — the primitives of the fetched row r+2 together with the primitives + CFL
  // integrand of the row completed by the PREVIOUS phase B (deferred by one row so that its long serial chain
  // overlaps other work instead of trailing the update).
  // PART 2: the last phase B of a segment (r = j1) in the peeled march — the x fluxes of row j1 and the primitives of
  // row j1+2 belong to the segment above; only the south face is solved, row j1-1 completed and the deferred CFL
  // integrand evaluated.
  template <bool LEAN, int PART = 0>
  E2D_HD void
  compute_B(const MarchArgs & a, const double xl[4], const double yl[4], const double fxE[4], const double uP[4],
            double fx[4], double fy[4], double un[4], double qP[4], double & ryP, double & cflv, bool & ok) const
  {
    const Settings & s = a.s;
    if (SOLVER == 2)
    {
      // west face of cell (i, r): left = XMAX of (i-1, r), right = XMIN of (i, r) (HydroRunFunctors.h:559-575)
      if (PART == 0)
        hllc_lean<LEAN, true>(s, a.c, xl[ID], xl[IP], xl[IU], xl[IV], xmin[ID], xmin[IP], xmin[IU], xmin[IV], fx[ID],
                              fx[IP], fx[IU], fx[IV], ok);
      // south face of row r: left = YMAX of row r-1, right = YMIN of row r, IU<->IV swapped (:621-640)
      hllc_lean<LEAN, true>(s, a.c, yl[ID], yl[IP], yl[IV], yl[IU], ymin[ID], ymin[IP], ymin[IV], ymin[IU], fy[ID], fy[IP],
                      fy[IV], fy[IU], ok);
    }
    else
    {
      if (PART == 0)
        riemann<SOLVER>(s, xl[ID], xl[IP], xl[IU], xl[IV], xmin[ID], xmin[IP], xmin[IU], xmin[IV], fx[ID], fx[IP], fx[IU],
                        fx[IV]);
      riemann<SOLVER>(s, yl[ID], yl[IP], yl[IV], yl[IU], ymin[ID], ymin[IP], ymin[IV], ymin[IU], fy[ID], fy[IP], fy[IV],
                      fy[IU]);
    }
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      if (PART == 0)
        fx[v] = fx[v] * dtdx;
      fy[v] = fy[v] * (SQUARE ? dtdx : dtdy);
    }
    // complete row r-1: UpdateFunctor order (HydroRunFunctors.h:695-713)
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      double x = pend[v]; // U + Fx(i, j)
      x -= fxE[v];        //   - Fx(i+1, j)
      x += fyP[v];        //   + Fy(i, j)
      x -= fy[v];         //   - Fy(i, j+1)
      un[v] = x;
    }
    cflv = 0.0;
    if (PART == 2)
    {
      if (FUSE_DT)
      {
        double u1[1][4], q1[1][4];
        Recip  rd1[1];
        E2D_UNROLL
        for (int v = 0; v < 4; ++v)
          u1[0][v] = unD[v];
        prim_lean_multi<LEAN, 1>(s, a.c, u1, q1, rd1, ok);
        cflv = cfl_tail_lean<LEAN>(s, recip_dx(a), recip_dy(a), q1[0], rd1[0], ok);
      }
    }
    else if (FUSE_DT)
    {
      double u2[2][4], q2[2][4];
      Recip  rd2[2];
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
      {
        u2[0][v] = unD[v];
        u2[1][v] = uP[v];
      }
      prim_lean_multi<LEAN, 2>(s, a.c, u2, q2, rd2, ok);
      cflv = cfl_tail_lean<LEAN>(s, recip_dx(a), recip_dy(a), q2[0], rd2[0], ok);
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
        qP[v] = q2[1][v];
      if (LEAN)
        ryP = rd2[1].y;
    }
    else
    {
      double u1[1][4], q1[1][4];
      Recip  rd1[1];
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
        u1[0][v] = uP[v];
      prim_lean_multi<LEAN, 1>(s, a.c, u1, q1, rd1, ok);
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
        qP[v] = q1[0][v];
      if (LEAN)
        ryP = rd1[0].y;
    }
  }

  // MATH == 1: the same phase with the fast arithmetic; fx, fy stay unscaled
  E2D_HD void
  compute_B_fast(const MarchArgs & a, const double xl[4], const double yl[4], const double fxE[4], const double uP[4],
                 double fx[4], double fy[4], double un[4], double qP[4], double & ryP, double & cflv) const
  {
    const Settings & s = a.s;
    if (SOLVER == 2)
    {
      fast::hllc(s, a.c, xl[ID], xl[IP], xl[IU], xl[IV], xmin[ID], xmin[IP], xmin[IU], xmin[IV], fx[ID], fx[IP], fx[IU],
                 fx[IV]);
      fast::hllc(s, a.c, yl[ID], yl[IP], yl[IV], yl[IU], ymin[ID], ymin[IP], ymin[IV], ymin[IU], fy[ID], fy[IP], fy[IV],
                 fy[IU]);
    }
    else
    {
      riemann<SOLVER>(s, xl[ID], xl[IP], xl[IU], xl[IV], xmin[ID], xmin[IP], xmin[IU], xmin[IV], fx[ID], fx[IP], fx[IU],
                      fx[IV]);
      riemann<SOLVER>(s, yl[ID], yl[IP], yl[IV], yl[IU], ymin[ID], ymin[IP], ymin[IV], ymin[IU], fy[ID], fy[IP], fy[IV],
                      fy[IU]);
    }
    // complete row r-1: U + Fx(i) - Fx(i+1) + Fy(j) - Fy(j+1) (HydroRunFunctors.h:695-713), pend = U + Fx(i)
    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      const double dty = SQUARE ? dtdx : dtdy;
      un[v] = fast::fmadd(-fy[v], dty, fast::fmadd(fyP[v], dty, fast::fmadd(-fxE[v], dtdx, pend[v])));
    }
    cflv = 0.0;
    if (FUSE_DT)
    {
      double qD[4], ryD;
      fast::prim(s, a.c, unD, qD, ryD);
      cflv = fast::cfl_tail(s, a.rdx_y, a.rdy_y, qD, ryD);
    }
    fast::prim(s, a.c, uP, qP, ryP);
  }

  // PART 0: a phase B as described at the top of this file — every phase B of the plain march.  The PEELED march (strict
  // arithmetic, short uniform segments: k_fused_step<.., PEEL>) treats the first and the last row of a segment apart,
  // because three of the Riemann solves of a segment serve nobody:
  // PART 1: the first phase B (r = j0-1) — nothing to solve, only the ring is advanced: row r+2 converted, row r+3
  //         fetched.  Phase B(j0) then "completes" row j0-1 from the start-up values of init (pending sum = row j0's
  //         state, zero east and south fluxes): a healthy state, discarded like before.
  // PART 2: the last one (r = j1) — only the south face, the update of row j1-1 and the deferred CFL integrand; nothing
  //         is fetched or converted any more.
  template <int PART = 0>
  E2D_HD void
  phaseB(const MarchArgs & a, MarchSmem<BX> & sm, int r)
  {
    const int sS = (m3 == 0) ? 2 : m3 - 1;
    double    xl[4], yl[4], fxE[4], uC[4], uP[4], fx[4], fy[4], un[4], qP[4], ryP = 0.0, cflv;
    if (PART == 1)
    {
      wait_prefetch(); // row r+2, in flight since init
      ld4<PACK>(sm.U[sS], t, uP);
      prefetch_row(a, sm, (r + 3 < a.jsize) ? r + 3 : a.jsize - 1, m3);
      convert_into(a, sm, uP, sS);
      m3 = (m3 == 2) ? 0 : m3 + 1;
      return;
    }
#if E2D_BULK
    if (MATH == 0)
    {
      bulk_wait_row(sm, r + 2); // row r+2, in flight since the previous phase B
      ld4<PACK>(sm.XMAX[r & 1], tm, xl);
      ld4<PACK>(sm.YMAX[(r - 1) & 1], t, yl);
      ld4<PACK>(sm.FX[r & 1], tp, fxE);
      ldb4(sm.UB[r & 3], t, uC);
      ldb4(sm.UB[(r + 2) & 3], t, uP);
      // row r+3 -> slot (r+3) & 3 = (r-1) & 3, last read (as uC) in phase B(r-1): every thread has passed barrier r since
      if (t == 0)
        bulk_fetch_row(a, sm, (r + 3 < a.jsize) ? r + 3 : a.jsize - 1, r + 3);
    }
    else
#endif
    if (PART == 0)
    {
    wait_prefetch(); // row r+2, in flight since phase A
    ld4<PACK>(sm.XMAX[r & 1], tm, xl);
    ld4<PACK>(sm.YMAX[(r - 1) & 1], t, yl);
    ld4<PACK>(sm.FX[r & 1], tp, fxE);
    ld4<PACK>(sm.U[m3], t, uC);
    ld4<PACK>(sm.U[sS], t, uP);
    // row r+3 (clamped: the last fetches of the topmost segment are harmless repeats) -> the U ring slot of row r,
    // just read; it is consumed by phase B(r+1), a whole B and A phase from here (own column only: no hazard)
    prefetch_row(a, sm, (r + 3 < a.jsize) ? r + 3 : a.jsize - 1, m3);
    }
    else
    {
      wait_prefetch(); // nothing may stay in flight when the block ends
      ld4<PACK>(sm.YMAX[(r - 1) & 1], t, yl);
      ld4<PACK>(sm.FX[r & 1], tp, fxE);
    }
    if (MATH == 1)
      compute_B_fast(a, xl, yl, fxE, uP, fx, fy, un, qP, ryP, cflv);
    else
    {
      bool ok = a.c.lean_ok != 0;
      compute_B<true, PART>(a, xl, yl, fxE, uP, fx, fy, un, qP, ryP, cflv, ok);
      if (!ok)
        compute_B<false, PART>(a, xl, yl, fxE, uP, fx, fy, un, qP, ryP, cflv, ok);
    }

    if (PART == 0)
    {
      st4<PACK>(sm.FX[(r + 1) & 1], t, fx);
      st4<PACK>(sm.Q[sS], t, qP); // row r+2 -> primitive ring, slot of row r-1 (last read in A(r))
      sm.RY[sS][t] = ryP;
    }

    // the CFL integrand just computed belongs to row r-2 (completed by the previous phase B)
    // invDt = fmax(invDt, v) (HydroRunFunctors.h:72) as compare + select: a NaN v compares false and is dropped like
    // fmax drops it, and the running maximum itself (0, then accepted values) is never a NaN
    if (FUSE_DT && store && r >= j0 + 2 && cflv > invdt)
      invdt = cflv;
    if (store && r >= j0 + 1)
    {
      const int jr = r - 1;
      double *  po = a.Uout + (jr * a.isize + i);
      E2D_UNROLL
      for (int v = 0; v < 4; ++v)
        po[v * plane] = un[v];
    }

    E2D_UNROLL
    for (int v = 0; v < 4; ++v)
    {
      if (PART == 0)
      {
        pend[v] = (MATH == 1) ? fast::fmadd(fx[v], dtdx, uC[v]) : uC[v] + fx[v];
        fyP[v] = fy[v];
      }
      unD[v] = un[v];
    }
    m3 = (m3 == 2) ? 0 : m3 + 1;
  }

  // phase B of row r in the peeled march, whichever part of the segment it is (host emulation of that march)
  E2D_HD void
  phaseB_peeled(const MarchArgs & a, MarchSmem<BX> & sm, int r)
  {
    if (r == j0 - 1)
      phaseB<1>(a, sm, r);
    else if (r == j1)
      phaseB<2>(a, sm, r);
    else
      phaseB<0>(a, sm, r);
  }

  // after the last phase B: the CFL integrand of the last completed row (j1-1)
  E2D_HD void
  finish(const MarchArgs & a)
  {
    if (!FUSE_DT)
      return;
    bool   ok = a.c.lean_ok != 0;
    double q[4];
    Recip  rd;
    if (MATH == 1)
    {
      fast::prim(a.s, a.c, unD, q, rd.y);
      if (store)
        invdt = fmax(invdt, fast::cfl_tail(a.s, a.rdx_y, a.rdy_y, q, rd.y));
      return;
    }
    prim_lean<true>(a.s, a.c, unD, q, rd, ok);
    double v = cfl_tail_lean<true>(a.s, recip_dx(a), recip_dy(a), q, rd, ok);
    if (!ok)
    {
      prim_lean<false>(a.s, a.c, unD, q, rd, ok);
      v = cfl_tail_lean<false>(a.s, recip_dx(a), recip_dy(a), q, rd, ok);
    }
    if (store && v > invdt)
      invdt = v;
  }
};

} // namespace e2d

#endif // E2D_MARCH_CUH
