/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's unsplit
 * MUSCL-Hancock Godunov step.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this.  The product (euler2d_kokkos_b200/) never does.
 *
 * Parity status: PINNED.  oracle/_ref/ is the reference's own source compiled here
 * (oracle/Makefile); tests/test_oracle_pins.py checks this restatement bit-for-bit against it
 * and against the committed golden fixtures generated from it (tests/golden/).
 *
 * Array layout everywhere: SoA planes, double, off = i + isize*(j + jsize*var),
 * var order ID=0 (rho) IP=IE=1 (E or p) IU=2 IV=3  (src/HydroParams.h:27-34,
 * = Kokkos LayoutLeft of src/kokkos_shared.h:21).
 */
#ifndef EULER2D_ORACLE_H
#define EULER2D_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { E2DO_ID = 0, E2DO_IP = 1, E2DO_IE = 1, E2DO_IU = 2, E2DO_IV = 3, E2DO_NBVAR = 4 };
enum { E2DO_FACE_XMIN = 0, E2DO_FACE_XMAX = 1, E2DO_FACE_YMIN = 2, E2DO_FACE_YMAX = 3 };
enum { E2DO_BC_UNDEFINED = 0, E2DO_BC_DIRICHLET = 1, E2DO_BC_NEUMANN = 2, E2DO_BC_PERIODIC = 3, E2DO_BC_COPY = 4 };
enum { E2DO_PROBLEM_IMPLODE = 0, E2DO_PROBLEM_BLAST, E2DO_PROBLEM_FOUR_QUADRANT, E2DO_PROBLEM_DISCONTINUITY,
       E2DO_PROBLEM_SHOCKED_BUBBLE };
enum { E2DO_RIEMANN_APPROX = 0, E2DO_RIEMANN_HLL = 1, E2DO_RIEMANN_HLLC = 2, E2DO_RIEMANN_RUSANOV = 3 /* extension */ };

/* field-for-field restatement of HydroParams + HydroSettings + ShockedBubbleParams
 * (src/HydroParams.h:107-265) */
typedef struct e2do_params
{
  int    nStepmax;
  double tEnd;
  int    nOutput;
  int    enableOutput;
  int    nx, ny, ghostWidth, imin, imax, jmin, jmax, isize, jsize;
  double xmin, xmax, ymin, ymax, dx, dy;
  int    boundary_type_xmin, boundary_type_xmax, boundary_type_ymin, boundary_type_ymax;
  int    ioVTK, ioHDF5;
  double gamma0, gamma6, cfl, slope_type, smallr, smallc, smallp, smallpp;
  int    niter_riemann, riemannSolverType, problemType;
  double blast_radius, blast_center_x, blast_center_y, blast_density_in, blast_density_out;
  double blast_pressure_in, blast_pressure_out, blast_total_energy_inside;
  int    blast_nbins;
  double bubble_radius, bubble_center_x, bubble_center_y, bubble_density, bubble_pressure;
  double preshock_density, preshock_pressure, postshock_density, postshock_pressure, postshock_velocity;
  double shock_loc;
  int    implementationVersion;
} e2do_params;

/* HydroParams::setup + init (src/HydroParams.cpp:43-190) on top of the inih/ConfigMap
 * semantics (config/inih/ini.cpp:65-150, config/ConfigMap.cpp:32-40). 0 = ok, -1 = cannot open
 * (the reference then silently runs with defaults; we fill the defaults and still return -1). */
int e2do_params_from_ini(const char * path, e2do_params * out);

/* ---- per-cell functions (src/HydroBaseFunctor.h) ---- */
void e2do_compute_primitives(const e2do_params * p, const double u[4], double * c, double q[4]);
void e2do_slope_unsplit_hydro_2d(const e2do_params * p, const double q[4], const double qPlusX[4],
                                 const double qMinusX[4], const double qPlusY[4], const double qMinusY[4],
                                 double dqX[4], double dqY[4]);
void e2do_trace_unsplit_2d_along_dir(const e2do_params * p, const double q[4], const double dqX[4],
                                     const double dqY[4], double dtdx, double dtdy, int faceId, double qface[4]);
void e2do_riemann_hllc(const e2do_params * p, const double qleft[4], const double qright[4], double flux[4]);
/* extension solvers (not in the reference: parity unpinned, see euler2d_oracle.c) and the switch that makes the array
 * operators / e2do_run use riemann_approx + cmpflx, HLL or Rusanov instead of the reference's hard-wired riemann_hllc */
void e2do_riemann_hll(const e2do_params * p, const double qleft[4], const double qright[4], double flux[4]);
void e2do_riemann_rusanov(const e2do_params * p, const double qleft[4], const double qright[4], double flux[4]);
void e2do_set_flux_solver(int solver);
int  e2do_get_flux_solver(void);
void e2do_riemann_approx(const e2do_params * p, const double qleft[4], const double qright[4], double qgdnv[4],
                         double flux[4]);
void e2do_cmpflx(const e2do_params * p, const double qgdnv[4], double flux[4]);

/* ---- array-level operators on a y-slab of the global grid ----
 * A slab is isize x jsize_loc cells (2 ghost rows each side), local row j <-> global row j + j_off.
 * The whole domain is the slab jsize_loc = p->jsize, j_off = 0. */
void   e2do_init_slab(const e2do_params * p, double * U, int jsize_loc, int j_off);
/* x faces always; y faces only when do_ymin / do_ymax (physical boundary owned by this slab) */
void   e2do_make_boundaries_slab(const e2do_params * p, double * U, int jsize_loc, int do_ymin, int do_ymax);
double e2do_compute_invdt_slab(const e2do_params * p, const double * U, int jsize_loc);
void   e2do_convert_to_primitives_slab(const e2do_params * p, const double * U, double * Q, int jsize_loc);
void   e2do_compute_and_store_fluxes_slab(const e2do_params * p, const double * Q, double * Fx, double * Fy,
                                          double dtdx, double dtdy, int jsize_loc);
void   e2do_update_slab(const e2do_params * p, double * U, const double * Fx, const double * Fy, int jsize_loc);
/* godunov_unsplit_impl (src/HydroRun.h:281-364) minus its make_boundaries call: out=in; Q; fluxes; update.
 * work = 3 arrays (Q,Fx,Fy) of isize*jsize_loc*4 doubles. */
void   e2do_godunov_slab(const e2do_params * p, const double * Uin, double * Uout, double * work, double dt,
                         int jsize_loc);

/* ---- Sedov post-processing: ComputeRadialProfileFunctor (src/ComputeRadialProfileFunctor.h:86-167) ----
 * Every cell of the array INCLUDING ghost cells (:106) is binned by the distance of its centre from the box centre,
 * bin = (int)(distance / max_radial_distance * nbins); counts[bin] += 1, sums[bin] += rho.  The reference writes past
 * the end of its bin arrays for the corner ghost cells (distance > max_radial_distance, no test at :160-165): those
 * samples are dropped here.  Serial, i outer / j inner (the order of the compiled reference on one thread).
 * rows [j_lo, j_hi) of the slab only (whole domain: 0, jsize).  distances[k] = (k + 0.5) * max_radial_distance / nbins.
 * The profile the reference saves is sums[k] / counts[k] (NaN for an empty bin, :130). */
void e2do_radial_profile_slab(const e2do_params * p, const double * U, int jsize_loc, int j_off, int j_lo, int j_hi,
                              int nbins, double * distances, double * sums, int * counts);

/* ---- whole-domain driver (src/main.cpp:86-143) ----
 * U, U2: isize*jsize*4 doubles each.  Runs until t >= tEnd or nStep >= max_steps (max_steps < 0: p->nStepmax).
 * dt_seq (may be NULL) receives dt of main.cpp:87 followed by the dt of every step (capacity dt_cap).
 * Returns the number of steps taken; the final state is in (nStep % 2 == 0 ? U : U2). */
int e2do_run(const e2do_params * p, double * U, double * U2, long max_steps, double * dt_seq, long dt_cap,
             double * t_out);

#ifdef __cplusplus
}
#endif
#endif
