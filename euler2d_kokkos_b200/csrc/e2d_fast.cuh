// "fast" arithmetic for the fused marching kernel: the opt-in `[other] arithmetic=fast`.
//
// The strict build (e2d_math.cuh / e2d_lean.cuh, -fmad=false) reproduces the reference's x86 arithmetic bit for bit
// and is FP64-pipe bound at ~500 FP64 instructions per cell.  north_star's parity bar for the state is
// a tolerance (relative L1/Linf <= 1e-12 per conserved variable), not bit equality, and SURVEY.md Appendix C measured
// that FMA contraction alone stays below 5e-13 on every deck.  This header spends that tolerance where it buys
// instructions — the same formulas of src/HydroBaseFunctor.h, evaluated with
//   * fused multiply-adds, written out explicitly (the library is compiled with -fmad=false, so nothing else changes);
//   * a/d as a * (1/d) with a Newton-refined reciprocal (error <= ~1 ulp; 1 FP64 instruction per quotient instead of 3,
//     no fast-path guards and no IEEE slow path: denominators are densities >= smallr and wave-speed differences,
//     far from the subnormal range the strict build still handles), sqrt likewise without the final correction;
//   * the minmod limiter on the raw differences (the sign test (dlft*drgt <= 0) as a sign-bit xor, one product with
//     slope_type after the selection);
//   * the conservative update as fma chains on the unscaled fluxes.
// ~300 FP64 instructions per cell.  Every function cites the reference formula it evaluates; tests/test_gpu_fast.py
// holds the tolerance (<= 1e-12 against the compiled reference after N steps, identical step count).
//
// Host-compilable like the other math headers (tests/host_emulation); on the host the reciprocal is 1.0/d.
#ifndef E2D_FAST_CUH
#define E2D_FAST_CUH

#include "e2d_lean.cuh"

namespace e2d
{
namespace fast
{

E2D_HD double
fmadd(double a, double b, double c)
{
#if E2D_LEAN_DEVICE
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}

// fmax(x, f) for a floor f > 0 (smallr, rho*smallp, smallc: guards of 1e-10 .. 1e-20) as ONE integer instruction: the
// signed maximum of the two high words (a double's bit pattern orders like a signed integer for non-negative values,
// and every negative x has a negative high word), low word of x kept.  x >= f comes back untouched; below the floor
// the result is f up to a relative 2^-20 — an absolute difference of 1e-16 at most, in a value that is an arbitrary
// guard to begin with.
E2D_HD double
floor_at(double x, double f)
{
#if E2D_LEAN_DEVICE
  return __hiloint2double(max(__double2hiint(x), __double2hiint(f)), __double2loint(x));
#else
  return x > f ? x : f;
#endif
}

// x > 0 on the high word (a subnormal x with an all-zero high word counts as zero)
E2D_HD bool
is_pos(double x)
{
#if E2D_LEAN_DEVICE
  return __double2hiint(x) > 0;
#else
  return x > 0.0;
#endif
}

// 1/d to ~1 ulp: MUFU.RCP64H seed (relative error e0 ~ 2^-20), one cubic step y(1 + e + e^2) -> e0^3, then rounding
E2D_HD double
rcp(double d)
{
#if E2D_LEAN_DEVICE
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = __fma_rn(-d, y, 1.0);
  e = __fma_rn(e, e, e);
  return __fma_rn(y, e, y);
#else
  return 1.0 / d;
#endif
}

// sqrt(x), x > 0 normal: MUFU.RSQ64H seed, one cubic step on 1/sqrt, one product
E2D_HD double
sqrt_pos(double x)
{
#if E2D_LEAN_DEVICE
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double g = __dmul_rn(y0, y0);
  const double e = __fma_rn(x, -g, 1.0);
  const double p = __fma_rn(e, 0.375, 0.5);
  const double h = __dmul_rn(y0, e);
  const double y1 = __fma_rn(p, h, y0);
  return __dmul_rn(x, y1);
#else
  return sqrt(x);
#endif
}

// 1/sqrt(x), x > 0 normal, to ~1 ulp
E2D_HD double
rsqrt_pos(double x)
{
#if E2D_LEAN_DEVICE
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double g = __dmul_rn(y0, y0);
  const double e = __fma_rn(x, -g, 1.0);
  const double p = __fma_rn(e, 0.375, 0.5);
  const double h = __dmul_rn(y0, e);
  return __fma_rn(p, h, y0);
#else
  return 1.0 / sqrt(x);
#endif
}

// computePrimitives without the sound speed (src/HydroBaseFunctor.h:76-99); ry = 1/d for the trace and the CFL tail.
// p = (gamma-1)*d*e with e = u_E/d - eken  is evaluated as (gamma-1)*(u_E - d*eken): one rounding less, no division.
E2D_HD void
prim(const Settings & s, const StepConsts & c, const double u[4], double q[4], double & ry)
{
  const double d = floor_at(u[ID], s.smallr);
  const double y = rcp(d);
  const double ux = u[IU] * y;
  const double uy = u[IV] * y;
  const double k2 = fmadd(ux, ux, uy * uy); // 2 * eken
  const double ei = fmadd(-0.5 * d, k2, u[IP]);
  q[ID] = d;
  q[IP] = floor_at(c.gm1 * ei, d * s.smallp);
  q[IU] = ux;
  q[IV] = uy;
  ry = y;
}

// the CFL integrand after the primitive conversion (src/HydroRunFunctors.h:60-72); idx = 1/dx, idy = 1/dy
E2D_HD double
cfl_tail(const Settings & s, double idx, double idy, const double q[4], double ry)
{
  const double cs = sqrt_pos(s.gamma0 * q[IP] * ry);
  const double vx = cs + fabs(q[IU]);
  const double vy = cs + fabs(q[IV]);
  return fmadd(vx, idx, vy * idy);
}

// slope_unsplit_hydro_2d_scalar (src/HydroBaseFunctor.h:433-442): dq = minmod(st*a, st*b, dcen), a = q - qMinus,
// b = qPlus - q.  (dlft*drgt <= 0) <=> a, b differ in sign or one is zero; a zero a or b is selected as the operand
// of least magnitude by itself, so only the sign test remains.  When a and b agree in sign, dcen has that sign too
// (or is zero), so the selected operand already carries the result's sign.
E2D_HD double
slope(double slope_type, double q, double qPlus, double qMinus)
{
  const double a = q - qMinus;
  const double b = qPlus - q;
  const double dcen = 0.5 * (qPlus - qMinus);
#if E2D_LEAN_DEVICE
  // flat -> multiply by 0 instead of slope_type (1.0 or 2.0: low word 0, so the choice is one 32-bit select); a zero
  // `sel` is then the operand of least magnitude and comes out as the (zero) slope
  const bool   flat = (__double2hiint(a) ^ __double2hiint(b)) < 0;
  const double st = __hiloint2double(flat ? 0 : __double2hiint(slope_type), 0);
#else
  const double st = (signbit(a) != signbit(b)) ? 0.0 : slope_type;
#endif
  const double sel = st * ((fabs(b) < fabs(a)) ? b : a);
  return (fabs(dcen) < fabs(sel)) ? dcen : sel;
}

// slope_type outside {1, 2}: zero slopes (src/HydroBaseFunctor.h:486-500), by multiplying with 0 — `st_or_zero` is
// slope_type when the limiter applies, else 0: every candidate then loses against sel = +-0, no select needed
E2D_HD void
slopes(double st_or_zero, const double q[4], const double qPlus[4], const double qMinus[4], double dq[4])
{
#pragma unroll
  for (int v = 0; v < 4; ++v)
    dq[v] = slope(st_or_zero, q[v], qPlus[v], qMinus[v]);
}

// trace_unsplit_2d_along_dir (src/HydroBaseFunctor.h:245-289), all four faces of one cell.  n0 = -s0 (the source
// terms with the sign pulled out); face = q -+ dq/2 + s0*dtdir/2, evaluated as fma(+-0.5, dq, fma(-n0, dtdir/2, q)).
// SQUARE: dx == dy is known at compile time (MarchThread<.., TYP = 1>); else tested at run time
template <bool SQUARE = false>
E2D_HD void
trace(const Settings & s, const double q[4], double ry, const double dqX[4], const double dqY[4], double hdtdx,
      double hdtdy, double xmin[4], double xmax[4], double ymin[4], double ymax[4])
{
  const double r = q[ID], p = q[IP], u = q[IU], v = q[IV];
  const double dv = dqX[IU] + dqY[IV];
  double       n0[4];
  n0[ID] = fmadd(u, dqX[ID], fmadd(v, dqY[ID], dv * r));
  n0[IP] = fmadd(u, dqX[IP], fmadd(v, dqY[IP], dv * (s.gamma0 * p)));
  n0[IU] = fmadd(u, dqX[IU], fmadd(v, dqY[IU], dqX[IP] * ry));
  n0[IV] = fmadd(u, dqX[IV], fmadd(v, dqY[IV], dqY[IP] * ry));
  double cx[4], cy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    cx[k] = cy[k] = fmadd(-n0[k], hdtdx, q[k]);
  if (!SQUARE && hdtdx != hdtdy) // square cells: the half-step predictor is shared by x and y
  {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      cy[k] = fmadd(-n0[k], hdtdy, q[k]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    xmin[k] = fmadd(-0.5, dqX[k], cx[k]);
    xmax[k] = fmadd(0.5, dqX[k], cx[k]);
    ymin[k] = fmadd(-0.5, dqY[k], cy[k]);
    ymax[k] = fmadd(0.5, dqY[k], cy[k]);
  }
  xmin[ID] = floor_at(xmin[ID], s.smallr);
  xmax[ID] = floor_at(xmax[ID], s.smallr);
  ymin[ID] = floor_at(ymin[ID], s.smallr);
  ymax[ID] = floor_at(ymax[ID], s.smallr);
}

// riemann_hllc (src/HydroBaseFunctor.h:704-809) on (rho, p, un, ut) -> flux (mass, energy, normal, transverse).
// Same one-sided sampling as hllc_lean (only the star state that can be sampled is evaluated).
E2D_HD void
hllc(const Settings & s, const StepConsts & c, double rl_in, double pl_in, double ul, double vl, double rr_in,
     double pr_in, double ur, double vr, double & f_d, double & f_e, double & f_n, double & f_t)
{
  // rl, rr: the inputs are traced face states, already floored at smallr (src/HydroBaseFunctor.h:279-289), so the
  // reference's fmax(rl, smallr) (:714,:723) is the identity here
  const double rl = rl_in;
  const double pl = floor_at(pl_in, rl * s.smallp);
  const double rr = rr_in;
  const double pr = floor_at(pr_in, rr * s.smallp);

  // fmax(cfastl, cfastr) (:732-737) = fmax(sqrt(gamma * fmax(pl/rl, pr/rr)), smallc); the larger ratio is found by
  // cross-multiplication, and sqrt(gamma p / r) = gamma p / sqrt(gamma p r) costs no reciprocal
  const bool   big_l = pl * rr > pr * rl;
  const double gp = s.gamma0 * (big_l ? pl : pr);
  const double cmax = floor_at(gp * rsqrt_pos(gp * (big_l ? rl : rr)), s.smallc);

  const double SL = min_nn(ul, ur) - cmax;
  const double SR = max_nn(ul, ur) + cmax;
  const double dl = ul - SL;
  const double dr = SR - ur;
  const double rcl = rl * dl;
  const double rcr = rr * dr;

  const double ys = rcp(rcr + rcl);
  const double ustar = fmadd(rcr, ur, fmadd(rcl, ul, pl - pr)) * ys;

  const bool   sup_l = is_pos(SL);
  const bool   side_l = sup_l || is_pos(ustar);
  const bool   star = !sup_l && (side_l || is_pos(SR));
  const double Sk = side_l ? SL : SR;
  const double rk = side_l ? rl : rr;
  const double pk = side_l ? pl : pr;
  const double uk = side_l ? ul : ur;
  const double vk = side_l ? vl : vr;
  const double dk = Sk - uk;  // SL - ul = -dl,  SR - ur = dr
  const double rck = rk * dk; // -rcl, rcr
  // total energy (:716-721, :725-730) of the side that is sampled only
  const double ek = fmadd(pk, c.entho, (0.5 * rk) * fmadd(uk, uk, vk * vk));
  // ptotstar (:752-753) is the reference's symmetric form of  p_k + rho_k (S_k - u_k)(ustar - u_k)  (either side
  // gives the same number when ustar satisfies :749-750); the one-sided form costs 2 instructions instead of 6.
  // The star-state formulas (:755-768) with ustar replaced by uk return the side state itself (pk, rk dk / dk,
  // ek dk / dk): the supersonic branches of the sampling (:770-797) need no selects of their own.
  const double uo = star ? ustar : uk;
  const double ptoto = fmadd(rck, uo - uk, pk);
  const double yk = rcp(Sk - uo);
  const double ro = rck * yk;
  const double etoto = fmadd(ptoto, uo, fmadd(dk, ek, -(pk * uk))) * yk;

  f_d = ro * uo;
  f_n = fmadd(f_d, uo, ptoto);
  f_e = (etoto + ptoto) * uo;
  f_t = f_d * (is_pos(f_d) ? vl : vr);
}

} // namespace fast
} // namespace e2d

#endif // E2D_FAST_CUH
