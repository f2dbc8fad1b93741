// TEST HARNESS (host) — runs the product's marching-kernel body (euler2d_kokkos_b200/csrc/e2d_march.cuh) and
// per-cell math (e2d_math.cuh) on the CPU, one emulated thread after the other with the barrier between the
// phases, so that the kernel's indexing / pipelining logic and formulas can be checked against the oracle on a
// machine without a GPU.  Compiled by tests/test_host_emulation.py with g++ -ffp-contract=off.  Never part of
// the product.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/euler2d_b200.h"
#include "../../euler2d_kokkos_b200/csrc/e2d_march.cuh"

using namespace e2d;

static Settings
settings_of(const e2d_params & p)
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

static int g_peel = 0; // emul_set_peel: drive the peeled march (phaseB<1>, <0>..., <2>) instead of the plain one

extern "C" void
emul_set_peel(int on)
{
  g_peel = on;
}

template <int BX, int SOLVER, int MATH = 0>
static double
run_blocks(const e2d_params & p, const double * Uin, double * Uout, int jsize_loc, double dt, int seg_rows)
{
  MarchArgs a;
  a.Uin = Uin;
  a.Uout = Uout;
  a.isize = p.isize;
  a.jsize = jsize_loc;
  a.seg_rows = seg_rows;
  a.s = settings_of(p);
  a.c = make_step_consts(a.s);
  a.dt = dt;
  a.d_dt = nullptr;
  a.invdt_bits = nullptr;
  a.rdx_y = 1.0 / a.s.dx; // only the fast arithmetic reads these on the host (the strict host path divides)
  a.rdy_y = 1.0 / a.s.dy;
  const int nx = p.nx, ny = jsize_loc - 4;
  const int nbx = (nx + (BX - 4) - 1) / (BX - 4);
  const int nseg = (ny + seg_rows - 1) / seg_rows;
  double    invdt = 0.0;
  using Thread = MarchThread<BX, SOLVER, true, MATH>;
  std::vector<Thread> th(BX);
  MarchSmem<BX> *     sm = new MarchSmem<BX>();
  for (int seg = 0; seg < nseg; ++seg)
    for (int bx = 0; bx < nbx; ++bx)
    {
      std::memset(sm, 0xff, sizeof(*sm)); // poison: NaNs if something is read before it is written
      bool active = true;
      for (int t = 0; t < BX; ++t)
        active = th[t].init(a, *sm, t, bx, seg, a.d_dt ? *a.d_dt : a.dt) && active;
      if (!active)
        continue;
      for (int r = th[0].j0 - 1; r <= th[0].j1; ++r)
      {
        for (int t = 0; t < BX; ++t)
          th[t].phaseA(a, *sm, r);
        // __syncthreads()
        for (int t = BX - 1; t >= 0; --t) // reverse order: phase B must not depend on intra-phase ordering
        {
          if (g_peel && MATH == 0)
            th[t].phaseB_peeled(a, *sm, r);
          else
            th[t].phaseB(a, *sm, r);
        }
      }
      for (int t = 0; t < BX; ++t)
      {
        th[t].finish(a);
        invdt = std::fmax(invdt, th[t].invdt);
      }
    }
  delete sm;
  return invdt;
}

extern "C" int
emul_fused_step(const e2d_params * p, const double * Uin, double * Uout, int jsize_loc, double dt, int seg_rows,
                int bx_threads, int solver, double * invdt_out)
{
  double inv = -1.0;
#define CASE(BX, SOL)                                                          \
  if (bx_threads == BX && solver == SOL)                                       \
    inv = run_blocks<BX, SOL>(*p, Uin, Uout, jsize_loc, dt, seg_rows);
  // `[other] arithmetic=fast`: the same state machine over e2d_fast.cuh (host: fma(), 1.0 / d)
  if (p->arithmetic == 1 && solver == 2 && bx_threads == 32)
    inv = run_blocks<32, 2, 1>(*p, Uin, Uout, jsize_loc, dt, seg_rows);
  else if (p->arithmetic == 1 && solver == 2 && bx_threads == 128)
    inv = run_blocks<128, 2, 1>(*p, Uin, Uout, jsize_loc, dt, seg_rows);
  else
  CASE(128, 2)
  CASE(32, 2)
  CASE(16, 2)
  CASE(32, 0)
  CASE(32, 1)
#undef CASE
  if (inv < 0)
    return 1;
  if (invdt_out)
    *invdt_out = inv;
  return 0;
}

// function-level evaluation of the product's formulas on the host (same record formats as e2d_k_eval_host)
extern "C" int
emul_eval(const e2d_params * p, const char * func, const double * in, double * out, long n)
{
  const Settings    s = settings_of(*p);
  const std::string f = func;
  for (long r = 0; r < n; ++r)
  {
    if (f == "prim")
    {
      const double * a = in + 4 * r;
      double *       o = out + 5 * r;
      compute_primitives(s, a[ID], a[IP], a[IU], a[IV], o[ID], o[IP], o[IU], o[IV], o[4]);
    }
    else if (f == "slope")
    {
      const double * a = in + 20 * r;
      double *       o = out + 8 * r;
      slopes_dir(s, a, a + 4, a + 8, o);
      slopes_dir(s, a, a + 12, a + 16, o + 4);
    }
    else if (f == "trace")
    {
      const double * a = in + 14 * r;
      double *       o = out + 16 * r;
      double         s0[4];
      trace_sources(s, a, a + 4, a + 8, s0);
      trace_face<-1>(s, a, a + 4, s0, a[12], o);
      trace_face<+1>(s, a, a + 4, s0, a[12], o + 4);
      trace_face<-1>(s, a, a + 8, s0, a[13], o + 8);
      trace_face<+1>(s, a, a + 8, s0, a[13], o + 12);
    }
    else if (f == "hllc" || f == "hll")
    {
      const double * a = in + 8 * r;
      double *       o = out + 4 * r;
      if (f == "hllc")
        riemann_hllc(s, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP], o[IU],
                     o[IV]);
      else
        riemann_hll(s, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP], o[IU],
                    o[IV]);
    }
    else if (f == "approx")
    {
      const double * a = in + 8 * r;
      double *       o = out + 8 * r;
      riemann_approx(s, a[ID], a[IP], a[IU], a[IV], a[4 + ID], a[4 + IP], a[4 + IU], a[4 + IV], o[ID], o[IP], o[IU],
                     o[IV], o[4 + ID], o[4 + IP], o[4 + IU], o[4 + IV]);
    }
    else if (f == "cmpflx")
    {
      const double * a = in + 4 * r;
      double *       o = out + 4 * r;
      cmpflx(s, a[ID], a[IP], a[IU], a[IV], o[ID], o[IP], o[IU], o[IV]);
    }
    else
      return 1;
  }
  return 0;
}
