// Per-cell hydrodynamics of the unsplit MUSCL-Hancock Godunov step, as __device__ functions.
//
// These follow the arithmetic of the reference's HydroBaseFunctor (src/HydroBaseFunctor.h)
// operation for operation and in the same association, so that a build without FMA contraction
// (-fmad=false) is bit-identical to the reference's x86 build: IEEE-754 double add/mul/div/sqrt
// are correctly rounded on both sides and fmax/fmin agree on non-NaN input.
//
// The header is also compilable by a host compiler (E2D_HD expands to `inline`); the test-suite
// uses that to check the *product's* formulas on machines without a GPU (tests/host_emulation).
// No product code path runs these on the CPU.
#ifndef E2D_MATH_CUH
#define E2D_MATH_CUH

#include <math.h>

#if defined(__CUDACC__)
#  define E2D_HD __host__ __device__ __forceinline__
#else
#  define E2D_HD inline
#endif

namespace e2d
{

enum { ID = 0, IP = 1, IE = 1, IU = 2, IV = 3, NBVAR = 4 };

// HydroSettings (src/HydroParams.h:107-132) + the mesh spacing: everything the kernels need.
struct Settings
{
  double gamma0, gamma6, cfl, slope_type, smallr, smallc, smallp, smallpp;
  double dx, dy;
};

// computePrimitives, src/HydroBaseFunctor.h:76-102 (the EOS :56-66 inlined, as the reference does)
E2D_HD void
compute_primitives(const Settings & s, double u_d, double u_e, double u_mx, double u_my, double & d,
                   double & p, double & ux, double & uy, double & c)
{
  d = fmax(u_d, s.smallr);
  ux = u_mx / d;
  uy = u_my / d;
  double eken = 0.5 * (ux * ux + uy * uy);
  double e = u_e / d - eken;
  p = fmax((s.gamma0 - 1.0) * d * e, d * s.smallp);
  c = sqrt(s.gamma0 * p / d);
}

// same without the sound speed (ConvertToPrimitivesFunctor discards it, src/HydroRunFunctors.h:130)
E2D_HD void
compute_primitives_noc(const Settings & s, double u_d, double u_e, double u_mx, double u_my, double & d,
                       double & p, double & ux, double & uy)
{
  d = fmax(u_d, s.smallr);
  ux = u_mx / d;
  uy = u_my / d;
  double eken = 0.5 * (ux * ux + uy * uy);
  double e = u_e / d - eken;
  p = fmax((s.gamma0 - 1.0) * d * e, d * s.smallp);
}

// the CFL integrand of ComputeDtFunctor, src/HydroRunFunctors.h:56-72
E2D_HD double
cfl_inv_dt(const Settings & s, double u_d, double u_e, double u_mx, double u_my)
{
  double d, p, ux, uy, c;
  compute_primitives(s, u_d, u_e, u_mx, u_my, d, p, ux, uy, c);
  double vx = c + fabs(ux);
  double vy = c + fabs(uy);
  return vx / s.dx + vy / s.dy;
}

// slope_unsplit_hydro_2d_scalar for ONE direction, src/HydroBaseFunctor.h:433-442.
// slope_type 0 yields exactly 0 through the same formula only for finite inputs, so the caller
// handles slope_type == 0 explicitly like :486-500.
E2D_HD double
slope_scalar(double slope_type, double q, double qPlus, double qMinus)
{
  double dlft = slope_type * (q - qMinus);
  double drgt = slope_type * (qPlus - q);
  double dcen = 0.5 * (qPlus - qMinus);
  double dsgn = (dcen >= 0.0) ? 1.0 : -1.0;
  double slop = fmin(fabs(dlft), fabs(drgt));
  double dlim = slop;
  if ((dlft * drgt) <= 0.0)
    dlim = 0.0;
  return dsgn * fmin(dlim, fabs(dcen));
}

// slope_unsplit_hydro_2d along one direction for the 4 variables, src/HydroBaseFunctor.h:473-516.
// slope_type outside {0,1,2} is undefined behaviour in the reference (dq left uninitialised); we
// return zeros there.
E2D_HD void
slopes_dir(const Settings & s, const double q[4], const double qPlus[4], const double qMinus[4], double dq[4])
{
  if (s.slope_type == 1.0 || s.slope_type == 2.0)
  {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      dq[v] = slope_scalar(s.slope_type, q[v], qPlus[v], qMinus[v]);
  }
  else
  {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      dq[v] = 0.0;
  }
}

// Source terms shared by the four face reconstructions, src/HydroBaseFunctor.h:245-249.
E2D_HD void
trace_sources(const Settings & s, const double q[4], const double dqX[4], const double dqY[4], double s0[4])
{
  const double r = q[ID], p = q[IP], u = q[IU], v = q[IV];
  const double drx = dqX[ID], dpx = dqX[IP], dux = dqX[IU], dvx = dqX[IV];
  const double dry = dqY[ID], dpy = dqY[IP], duy = dqY[IU], dvy = dqY[IV];
  s0[ID] = -u * drx - v * dry - (dux + dvy) * r;
  s0[IP] = -u * dpx - v * dpy - (dux + dvy) * s.gamma0 * p;
  s0[IU] = -u * dux - v * duy - (dpx) / r;
  s0[IV] = -u * dvx - v * dvy - (dpy) / r;
}

// One face of trace_unsplit_2d_along_dir, src/HydroBaseFunctor.h:251-289:
//   qface = q -/+ 0.5*dq + s0*dtdir*0.5 ; rho floored at smallr.   sign = -1 for the MIN face, +1 for MAX.
template <int sign>
E2D_HD void
trace_face(const Settings & s, const double q[4], const double dq[4], const double s0[4], double dtdir,
           double qface[4])
{
#pragma unroll
  for (int v = 0; v < 4; ++v)
  {
    if (sign < 0)
      qface[v] = q[v] - 0.5 * dq[v] + s0[v] * dtdir * 0.5;
    else
      qface[v] = q[v] + 0.5 * dq[v] + s0[v] * dtdir * 0.5;
  }
  qface[ID] = fmax(s.smallr, qface[ID]);
}

// riemann_hllc, src/HydroBaseFunctor.h:704-809.  Written on (rho, p, un, ut) = density, pressure,
// normal and transverse velocity so that the IU<->IV swap of the y sweep
// (src/HydroRunFunctors.h:628-632) is a matter of argument order.  Outputs the four flux
// components (mass, energy, normal momentum, transverse momentum).
E2D_HD void
riemann_hllc(const Settings & s, double rl_in, double pl_in, double ul, double vl, double rr_in, double pr_in,
             double ur, double vr, double & f_d, double & f_e, double & f_n, double & f_t)
{
  const double entho = 1.0 / (s.gamma0 - 1.0);

  // Left variables
  double rl = fmax(rl_in, s.smallr);
  double pl = fmax(pl_in, rl * s.smallp);
  double ecinl = 0.5 * rl * ul * ul;
  ecinl += 0.5 * rl * vl * vl;
  double etotl = pl * entho + ecinl;

  // Right variables
  double rr = fmax(rr_in, s.smallr);
  double pr = fmax(pr_in, rr * s.smallp);
  double ecinr = 0.5 * rr * ur * ur;
  ecinr += 0.5 * rr * vr * vr;
  double etotr = pr * entho + ecinr;

  // largest eigenvalues normal to the interface
  double cfastl = sqrt(fmax(s.gamma0 * pl / rl, s.smallc * s.smallc));
  double cfastr = sqrt(fmax(s.gamma0 * pr / rr, s.smallc * s.smallc));

  // HLL wave speeds
  double SL = fmin(ul, ur) - fmax(cfastl, cfastr);
  double SR = fmax(ul, ur) + fmax(cfastl, cfastr);

  // lagrangian sound speeds
  double rcl = rl * (ul - SL);
  double rcr = rr * (SR - ur);

  // acoustic star state
  double ustar = (rcr * ur + rcl * ul + (pl - pr)) / (rcr + rcl);
  double ptotstar = (rcr * pl + rcl * pr + rcl * rcr * (ul - ur)) / (rcr + rcl);

  // star regions
  double rstarl = rl * (SL - ul) / (SL - ustar);
  double etotstarl = ((SL - ul) * etotl - pl * ul + ptotstar * ustar) / (SL - ustar);
  double rstarr = rr * (SR - ur) / (SR - ustar);
  double etotstarr = ((SR - ur) * etotr - pr * ur + ptotstar * ustar) / (SR - ustar);

  // sample at x/t = 0
  double ro, uo, ptoto, etoto;
  if (SL > 0.0)
  {
    ro = rl;
    uo = ul;
    ptoto = pl;
    etoto = etotl;
  }
  else if (ustar > 0.0)
  {
    ro = rstarl;
    uo = ustar;
    ptoto = ptotstar;
    etoto = etotstarl;
  }
  else if (SR > 0.0)
  {
    ro = rstarr;
    uo = ustar;
    ptoto = ptotstar;
    etoto = etotstarr;
  }
  else
  {
    ro = rr;
    uo = ur;
    ptoto = pr;
    etoto = etotr;
  }

  // Godunov flux
  f_d = ro * uo;
  f_n = ro * uo * uo + ptoto;
  f_e = (etoto + ptoto) * uo;
  if (f_d > 0.0)
    f_t = f_d * vl;
  else
    f_t = f_d * vr;
}

// cmpflx, src/HydroBaseFunctor.h:523-547, on (rho, p, un, ut)
E2D_HD void
cmpflx(const Settings & s, double g_d, double g_p, double g_n, double g_t, double & f_d, double & f_e,
       double & f_n, double & f_t)
{
  f_d = g_d * g_n;
  f_n = f_d * g_n + g_p;
  f_t = f_d * g_t;
  double entho = 1.0 / (s.gamma0 - 1.0);
  double ekin = 0.5 * g_d * (g_n * g_n + g_t * g_t);
  double etot = g_p * entho + ekin;
  f_e = g_n * (etot + g_p);
}

// riemann_approx, src/HydroBaseFunctor.h:558-693 (two-shock iterative solver of RAMSES).  Dead
// code in the reference (no kernel calls it), provided for the opt-in `honourRiemannSolver` path.
// The iteration cap (10) and tolerance (1e-6) are the literals of :592; `2.0f` at :596-597 is exact.
E2D_HD void
riemann_approx(const Settings & s, double rl_in, double pl_in, double ul, double vl, double rr_in, double pr_in,
               double ur, double vr, double & g_d, double & g_p, double & g_n, double & g_t, double & f_d,
               double & f_e, double & f_n, double & f_t)
{
  const double gamma0 = s.gamma0, gamma6 = s.gamma6, smallr = s.smallr, smallc = s.smallc;
  const double smallp = s.smallp, smallpp = s.smallpp;

  double rl = fmax(rl_in, smallr);
  double pl = fmax(pl_in, rl * smallp);
  double rr = fmax(rr_in, smallr);
  double pr = fmax(pr_in, rr * smallp);

  // Lagrangian sound speed
  double cl = gamma0 * pl * rl;
  double cr = gamma0 * pr * rr;

  // first guess
  double wl = sqrt(cl);
  double wr = sqrt(cr);
  double pstar = fmax(((wr * pl + wl * pr) + wl * wr * (ul - ur)) / (wl + wr), 0.0);
  double pold = pstar;
  double conv = 1.0;

  // Newton-Raphson on pstar
  for (int iter = 0; (iter < 10) && (conv > 1e-6); ++iter)
  {
    double wwl = sqrt(cl * (1.0 + gamma6 * (pold - pl) / pl));
    double wwr = sqrt(cr * (1.0 + gamma6 * (pold - pr) / pr));
    double ql = 2.0 * wwl * wwl * wwl / (wwl * wwl + cl);
    double qr = 2.0 * wwr * wwr * wwr / (wwr * wwr + cr);
    double usl = ul - (pold - pl) / wwl;
    double usr = ur + (pold - pr) / wwr;
    double delp = fmax(qr * ql / (qr + ql) * (usl - usr), -pold);
    pold = pold + delp;
    conv = fabs(delp / (pold + smallpp));
  }

  // star region pressure and velocity for a two-shock problem
  pstar = pold;
  wl = sqrt(cl * (1.0 + gamma6 * (pstar - pl) / pl));
  wr = sqrt(cr * (1.0 + gamma6 * (pstar - pr) / pr));
  double ustar = 0.5 * (ul + (pl - pstar) / wl + ur - (pr - pstar) / wr);

  // left- or right-going contact
  double sgnm = copysign(1.0, ustar);
  double ro, uo, po, wo;
  if (sgnm > 0.0)
  {
    ro = rl;
    uo = ul;
    po = pl;
    wo = wl;
  }
  else
  {
    ro = rr;
    uo = ur;
    po = pr;
    wo = wr;
  }
  double co = fmax(smallc, sqrt(fabs(gamma0 * po / ro)));
  double rstar = fmax(ro / (1.0 + ro * (po - pstar) / (wo * wo)), smallr);
  double cstar = fmax(smallc, sqrt(fabs(gamma0 * pstar / rstar)));

  // rarefaction head / tail and shock speeds
  double spout = co - sgnm * uo;
  double spin = cstar - sgnm * ustar;
  double ushock = wo / ro - sgnm * uo;
  if (pstar >= po)
  {
    spin = ushock;
    spout = ushock;
  }

  // sample at x/t = 0
  double scr = fmax(spout - spin, smallc + fabs(spout + spin));
  double frac = 0.5 * (1.0 + (spout + spin) / scr);
  if (frac != frac)
    frac = 0.0;
  else
    frac = frac >= 1.0 ? 1.0 : frac <= 0.0 ? 0.0 : frac;

  g_d = frac * rstar + (1.0 - frac) * ro;
  g_n = frac * ustar + (1.0 - frac) * uo;
  g_p = frac * pstar + (1.0 - frac) * po;
  if (spout < 0.0)
  {
    g_d = ro;
    g_n = uo;
    g_p = po;
  }
  if (spin > 0.0)
  {
    g_d = rstar;
    g_n = ustar;
    g_p = pstar;
  }
  g_t = (sgnm > 0.0) ? vl : vr;

  cmpflx(s, g_d, g_p, g_n, g_t, f_d, f_e, f_n, f_t);
}

// HLL (two-wave) solver.  NOT in the reference (RIEMANN_HLL is an enum value nothing implements,
// src/HydroParams.h:46-51) — an extension behind `honourRiemannSolver`, using the
// same wave-speed estimates as riemann_hllc.  Parity unpinned by construction; tested for
// consistency F(q,q) = F(q) and against HLLC on supersonic states.
E2D_HD void
riemann_hll(const Settings & s, double rl_in, double pl_in, double ul, double vl, double rr_in, double pr_in,
            double ur, double vr, double & f_d, double & f_e, double & f_n, double & f_t)
{
  const double entho = 1.0 / (s.gamma0 - 1.0);
  double       rl = fmax(rl_in, s.smallr);
  double       pl = fmax(pl_in, rl * s.smallp);
  double       rr = fmax(rr_in, s.smallr);
  double       pr = fmax(pr_in, rr * s.smallp);
  double       etotl = pl * entho + 0.5 * rl * (ul * ul + vl * vl);
  double       etotr = pr * entho + 0.5 * rr * (ur * ur + vr * vr);
  double       cfastl = sqrt(fmax(s.gamma0 * pl / rl, s.smallc * s.smallc));
  double       cfastr = sqrt(fmax(s.gamma0 * pr / rr, s.smallc * s.smallc));
  double       SL = fmin(ul, ur) - fmax(cfastl, cfastr);
  double       SR = fmax(ul, ur) + fmax(cfastl, cfastr);
  // physical fluxes and conserved states on both sides
  double fl_d = rl * ul, fl_n = rl * ul * ul + pl, fl_t = rl * ul * vl, fl_e = (etotl + pl) * ul;
  double fr_d = rr * ur, fr_n = rr * ur * ur + pr, fr_t = rr * ur * vr, fr_e = (etotr + pr) * ur;
  if (SL >= 0.0)
  {
    f_d = fl_d;
    f_n = fl_n;
    f_t = fl_t;
    f_e = fl_e;
  }
  else if (SR <= 0.0)
  {
    f_d = fr_d;
    f_n = fr_n;
    f_t = fr_t;
    f_e = fr_e;
  }
  else
  {
    double inv = 1.0 / (SR - SL);
    f_d = (SR * fl_d - SL * fr_d + SL * SR * (rr - rl)) * inv;
    f_n = (SR * fl_n - SL * fr_n + SL * SR * (rr * ur - rl * ul)) * inv;
    f_t = (SR * fl_t - SL * fr_t + SL * SR * (rr * vr - rl * vl)) * inv;
    f_e = (SR * fl_e - SL * fr_e + SL * SR * (etotr - etotl)) * inv;
  }
}

// Rusanov (local Lax-Friedrichs) flux.  NOT in the reference either (north_star names "approximate (Rusanov/HLL)"
// solvers) — an extension behind `honourRiemannSolver` + `riemann=rusanov`.  Parity unpinned by construction; the
// oracle carries a CPU restatement of the same formula (e2do_riemann_rusanov) that this is bit-identical to.
//   F = (FL + FR)/2 - smax (UR - UL)/2,  smax = max(|ul| + cl, |ur| + cr), floors and sound speeds as in riemann_hllc
E2D_HD void
riemann_rusanov(const Settings & s, double rl_in, double pl_in, double ul, double vl, double rr_in, double pr_in,
                double ur, double vr, double & f_d, double & f_e, double & f_n, double & f_t)
{
  const double entho = 1.0 / (s.gamma0 - 1.0);
  double       rl = fmax(rl_in, s.smallr);
  double       pl = fmax(pl_in, rl * s.smallp);
  double       rr = fmax(rr_in, s.smallr);
  double       pr = fmax(pr_in, rr * s.smallp);
  double       etotl = pl * entho + 0.5 * rl * (ul * ul + vl * vl);
  double       etotr = pr * entho + 0.5 * rr * (ur * ur + vr * vr);
  double       cl = sqrt(fmax(s.gamma0 * pl / rl, s.smallc * s.smallc));
  double       cr = sqrt(fmax(s.gamma0 * pr / rr, s.smallc * s.smallc));
  double       smax = fmax(fabs(ul) + cl, fabs(ur) + cr);
  double       fl_d = rl * ul, fl_n = rl * ul * ul + pl, fl_t = rl * ul * vl, fl_e = (etotl + pl) * ul;
  double       fr_d = rr * ur, fr_n = rr * ur * ur + pr, fr_t = rr * ur * vr, fr_e = (etotr + pr) * ur;
  f_d = 0.5 * (fl_d + fr_d) - 0.5 * smax * (rr - rl);
  f_n = 0.5 * (fl_n + fr_n) - 0.5 * smax * (rr * ur - rl * ul);
  f_t = 0.5 * (fl_t + fr_t) - 0.5 * smax * (rr * vr - rl * vl);
  f_e = 0.5 * (fl_e + fr_e) - 0.5 * smax * (etotr - etotl);
}

// Riemann solve at a face, selected at compile time.  solver: 0 approx, 1 hll, 2 hllc, 3 rusanov.
template <int solver>
E2D_HD void
riemann(const Settings & s, double rl, double pl, double ul, double vl, double rr, double pr, double ur,
        double vr, double & f_d, double & f_e, double & f_n, double & f_t)
{
  if (solver == 2)
    riemann_hllc(s, rl, pl, ul, vl, rr, pr, ur, vr, f_d, f_e, f_n, f_t);
  else if (solver == 1)
    riemann_hll(s, rl, pl, ul, vl, rr, pr, ur, vr, f_d, f_e, f_n, f_t);
  else if (solver == 3)
    riemann_rusanov(s, rl, pl, ul, vl, rr, pr, ur, vr, f_d, f_e, f_n, f_t);
  else
  {
    double g_d, g_p, g_n, g_t;
    riemann_approx(s, rl, pl, ul, vl, rr, pr, ur, vr, g_d, g_p, g_n, g_t, f_d, f_e, f_n, f_t);
  }
}

} // namespace e2d

#endif // E2D_MATH_CUH
