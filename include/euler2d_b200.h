/* euler2d_b200 — C ABI of the B200-native unsplit MUSCL-Hancock Godunov step.
 *
 * This is the drop-in boundary for the hot path of pkestene/euler2d_kokkos.  The reference has no
 * FFI; its seam is the C++ class euler2d::HydroRun<device_t> (src/HydroRun.h:44-134) as driven by
 * src/main.cpp:76-143, plus the static XxxFunctor::apply() operators underneath it
 * (src/HydroRunFunctors.h).  Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only; all arrays are fp64, SoA planes  off = i + isize*(j + jsize*var)
 *     (= Kokkos LayoutLeft of DataArray, src/kokkos_shared.h:21), var: 0 rho, 1 E (or p), 2 rho*u (u), 3 rho*v (v)
 *   - every function returns an e2d_status (0 = ok); nothing throws, nothing calls exit()
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream)
 *   - kernel-level entry points (e2d_k_*) take DEVICE pointers and only enqueue work
 *   - there is no CPU fallback: without a CUDA device every compute call returns E2D_ERR_CUDA
 *   - a "slab" is a horizontal strip of the global grid owned by one GPU: isize x jsize_loc cells
 *     including 2 ghost rows on each side; local row j is global row j + j_off.  The whole domain is
 *     the slab jsize_loc = jsize, j_off = 0.
 */
#ifndef EULER2D_B200_H
#define EULER2D_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum e2d_status
{
  E2D_OK = 0,
  E2D_ERR_INVALID = 1, /* bad argument */
  E2D_ERR_IO = 2,      /* cannot open / write a file */
  E2D_ERR_CUDA = 3,    /* CUDA runtime error or no device (see e2d_last_error) */
  E2D_ERR_ALLOC = 4,
  E2D_ERR_UNSUPPORTED = 5
} e2d_status;

/* component / face / boundary / problem ids: src/HydroParams.h:27-104 */
enum { E2D_ID = 0, E2D_IP = 1, E2D_IE = 1, E2D_IU = 2, E2D_IV = 3, E2D_NBVAR = 4 };
enum { E2D_FACE_XMIN = 0, E2D_FACE_XMAX = 1, E2D_FACE_YMIN = 2, E2D_FACE_YMAX = 3 };
enum { E2D_BC_UNDEFINED = 0, E2D_BC_DIRICHLET = 1, E2D_BC_NEUMANN = 2, E2D_BC_PERIODIC = 3, E2D_BC_COPY = 4 };
enum { E2D_PROBLEM_IMPLODE = 0, E2D_PROBLEM_BLAST = 1, E2D_PROBLEM_FOUR_QUADRANT = 2,
       E2D_PROBLEM_DISCONTINUITY = 3, E2D_PROBLEM_SHOCKED_BUBBLE = 4 };
enum { E2D_RIEMANN_APPROX = 0, E2D_RIEMANN_HLL = 1, E2D_RIEMANN_HLLC = 2,
       E2D_RIEMANN_RUSANOV = 3 /* extension: `riemann=rusanov`, unknown to the reference's parser (HydroParams.cpp:86-105) */ };
enum { E2D_ARITH_STRICT = 0, E2D_ARITH_FAST = 1 };
/* bit mask of faces for e2d_k_make_boundaries */
enum { E2D_FACES_X = 3, E2D_FACES_YMIN = 4, E2D_FACES_YMAX = 8, E2D_FACES_ALL = 15 };
/* host array layouts for upload / download */
enum { E2D_LAYOUT_SOA = 0,      /* [var][j][i]   (device layout, Kokkos LayoutLeft)          */
       E2D_LAYOUT_KOKKOS_OMP = 1 /* (i*jsize+j)*4+var (Kokkos LayoutRight = the OpenMP build) */ };

/* Replaces: struct HydroParams + HydroSettings + ShockedBubbleParams (src/HydroParams.h:107-265). */
typedef struct e2d_params
{
  int    nStepmax;
  double tEnd;
  int    nOutput;
  int    enableOutput;
  int    nx, ny, ghostWidth, imin, imax, jmin, jmax, isize, jsize;
  double xmin, xmax, ymin, ymax, dx, dy;
  int    boundary_type_xmin, boundary_type_xmax, boundary_type_ymin, boundary_type_ymax;
  int    ioVTK, ioHDF5;
  double gamma0, gamma6, cfl, slope_type, smallr, smallc, smallp, smallpp; /* HydroSettings */
  int    niter_riemann, riemannSolverType, problemType;
  double blast_radius, blast_center_x, blast_center_y, blast_density_in, blast_density_out;
  double blast_pressure_in, blast_pressure_out, blast_total_energy_inside;
  int    blast_nbins;
  double bubble_radius, bubble_center_x, bubble_center_y, bubble_density, bubble_pressure;
  double preshock_density, preshock_pressure, postshock_density, postshock_pressure, postshock_velocity;
  double shock_loc;
  int    implementationVersion;
  /* [output] outputDir / outputPrefix, read by HydroRun::saveVTK (src/HydroRun.h:526-527) */
  char   outputDir[256];
  char   outputPrefix[256];
  /* extension (not in the reference, where `riemann=` is parsed but never used — SURVEY.md §0.4):
   * 0 = reference behaviour (every kernel solves HLLC), 1 = honour riemannSolverType. */
  int    honourRiemannSolver;
  /* extension: `[output] vtk_appended=yes` makes saveData write the .vti with raw appended binary data (full
   * precision, interior gathered on the device, D2H overlapped with the file writes) instead of the reference's
   * 6-digit ascii; 0 = reference behaviour. */
  int    vtkAppended;
  /* extension: `[other] arithmetic=strict|fast`.  0 = strict (default): IEEE double without FMA contraction,
   * bit-identical to the reference's Kokkos/OpenMP x86 build.  1 = fast: the fused step evaluates the same
   * formulas with fused multiply-adds and reciprocal-multiply division (csrc/e2d_fast.cuh) — NOT bit-identical.
   * What the tests enforce (tests/test_gpu_fast.py; metric: euler2d_kokkos_b200/parity.py): same step count and, per
   * conserved variable against the reference, relative L1 <= 1e-14 always; relative Linf <= 1e-12 (north_star's
   * tolerance) through 200 steps on every deck up to 4096^2; Linf <= 1e-11 at 300 steps.  The Linf bound loosens with
   * the step count because last-bit differences are amplified where the flow is unstable (four_quadrant's corner
   * interaction: 3.4e-12 at step 300, whichever approximation is made exact: profiles/r2n_fast_exactness_variants.txt),
   * which is why strict is the default and the bench headline.  Only the fused step kernel (e2d_godunov_unsplit unless
   * `unfusedKernels`, e2d_run, e2d_step_host*) with the HLLC solver has a fast form; everything else ignores it. */
  int    arithmetic;
  /* extension: `[other] unfusedKernels=yes` makes e2d_godunov_unsplit run implementationVersion 0 / 1 as the
   * reference's literal kernel sequence (deep_copy, ConvertToPrimitives, ComputeAndStoreFluxes + Update, or the
   * slopes / trace / update-per-direction trio; Q, flux and slope arrays allocated).  Default 0: the fused step kernel
   * followed by a copy of the ghost frame computes the very same output array — every bit, ghost cells included
   * (implementations 0 and 1 are bit-identical in the reference) — at 64 instead of 384 bytes of HBM
   * traffic per cell, so the reference's decks (all implementationVersion 0) run at full speed.  The operator-level
   * kernels stay available either way as e2d_k_*. */
  int    unfusedKernels;
} e2d_params;

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                    */
/* ------------------------------------------------------------------------------------------ */
const char * e2d_version(void);
const char * e2d_status_string(int status);
/* message of the last failing CUDA/IO call on this thread ("" if none) */
const char * e2d_last_error(void);
/* number of CUDA devices visible (0 without a driver/GPU); never fails */
int          e2d_device_count(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long e2d_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------ */
/* parameters: ConfigMap + HydroParams::setup + HydroParams::init                             */
/*   replaces config/ConfigMap.{h,cpp}, config/inih/, src/HydroParams.cpp:43-190               */
/*   (all reals go through strtof, exactly like ConfigMap::getFloat)                          */
/* ------------------------------------------------------------------------------------------ */
int e2d_params_from_ini(const char * path, e2d_params * out);
/* same, from an in-memory .ini text (used by tests and by callers that build decks on the fly) */
int e2d_params_from_string(const char * ini_text, e2d_params * out);
/* ConfigMap (config/ConfigMap.h:26-46 over config/inih/INIReader.h): the parsed .ini as a key/value map with the
 * reference's typed getters — getFloat goes through strtof (ConfigMap.cpp:32-40), integers through strtol base 0,
 * booleans accept 1/yes/true/on and 0/no/false/off; keys are "section.name", lower-cased; the last assignment wins.
 * e2d_config_open returns E2D_ERR_IO for a missing file but still hands out a valid (empty) map, because the
 * reference never checks ParseError() (src/main.cpp:76). */
typedef struct e2d_config e2d_config;
int   e2d_config_open(const char * path, e2d_config ** out);
int   e2d_config_from_string(const char * ini_text, e2d_config ** out);
void  e2d_config_close(e2d_config * c);
int   e2d_config_parse_error(const e2d_config * c); /* INIReader::ParseError(): 0, or -1 when the file did not open */
float e2d_config_get_float(const e2d_config * c, const char * section, const char * name, float default_value);
long  e2d_config_get_integer(const e2d_config * c, const char * section, const char * name, long default_value);
int   e2d_config_get_bool(const e2d_config * c, const char * section, const char * name, int default_value);
int   e2d_config_get_string(const e2d_config * c, const char * section, const char * name, const char * default_value,
                            char * buf, size_t cap);
/* ConfigMap::setFloat / setBool and friends: values are stored as text, like the reference's map */
int   e2d_config_set_string(e2d_config * c, const char * section, const char * name, const char * value);
/* HydroParams::setup(ConfigMap &) (src/HydroParams.cpp:43-155), including init() */
int   e2d_params_setup(e2d_params * out, const e2d_config * c);
/* recompute the derived fields after editing nx/ny/xmin/... (HydroParams::init, :161-190) */
int e2d_params_init(e2d_params * p);
/* HydroParams::print (src/HydroParams.cpp:196-224), same text, to stdout */
int e2d_params_print(const e2d_params * p);

/* ------------------------------------------------------------------------------------------ */
/* kernel-level operators (device pointers; replace the XxxFunctor::apply statics)            */
/* ------------------------------------------------------------------------------------------ */
/* Init{Implode,Blast,FourQuadrant,Discontinuity,ShockedBubble}Functor::apply
 * (src/HydroRunFunctors.h:1347-1827) on a slab; fills ghosts too, like the reference. */
int e2d_k_init_problem(const e2d_params * p, double * U, int jsize_loc, int j_off, void * stream);

/* MakeBoundariesFunctor<face>::apply x4 in the order of HydroRun::make_boundaries
 * (src/HydroRunFunctors.h:1832-2030, src/HydroRun.h:390-399), as ONE launch.  faces = bit mask. */
int e2d_k_make_boundaries(const e2d_params * p, double * U, int jsize_loc, int faces, void * stream);

/* ComputeDtFunctor::apply (src/HydroRunFunctors.h:17-79): d_invdt[0] = max(d_invdt[0], max over the
 * slab's interior cells of (c+|u|)/dx + (c+|v|)/dy).  d_invdt is a DEVICE double the caller has set
 * to 0 (non-negative values only: the reduction is an integer atomicMax on the bit pattern). */
int e2d_k_reduce_invdt(const e2d_params * p, const double * U, int jsize_loc, double * d_invdt, void * stream);

/* ConvertToPrimitivesFunctor::apply (src/HydroRunFunctors.h:84-143) */
int e2d_k_convert_to_primitives(const e2d_params * p, const double * U, double * Q, int jsize_loc, void * stream);

/* ComputeAndStoreFluxesFunctor::apply (src/HydroRunFunctors.h:412-651): Fx, Fy = flux*dt/dx, flux*dt/dy */
int e2d_k_compute_and_store_fluxes(const e2d_params * p, const double * Q, double * Fx, double * Fy,
                                   double dtdx, double dtdy, int jsize_loc, void * stream);

/* UpdateFunctor::apply (src/HydroRunFunctors.h:656-723) */
int e2d_k_update(const e2d_params * p, double * U, const double * Fx, const double * Fy, int jsize_loc,
                 void * stream);

/* implementation-1 trio (src/HydroRunFunctors.h:986-1342): dir = 1 (XDIR) or 2 (YDIR) */
int e2d_k_compute_slopes(const e2d_params * p, const double * Q, double * Sx, double * Sy, int jsize_loc,
                         void * stream);
int e2d_k_compute_trace_and_fluxes(const e2d_params * p, const double * Q, const double * Sx, const double * Sy,
                                   double * F, double dtdx, double dtdy, int dir, int jsize_loc, void * stream);
int e2d_k_update_dir(const e2d_params * p, double * U, const double * F, int dir, int jsize_loc, void * stream);

/* The fused single-kernel step — replaces ConvertToPrimitives + ComputeFluxesAndUpdateFunctor::apply
 * (src/HydroRunFunctors.h:728-980, "implementationVersion 2") *and* the deep_copy of
 * src/HydroRun.h:302, deterministically (no atomics on the state): reads Uin (ghosts filled),
 * writes the interior of Uout, bit-identical to the unfused implementation 0.
 *   dt         time step (used when d_dt == NULL)
 *   d_dt       optional DEVICE pointer to dt (device-resident time loop)
 *   d_invdt    optional DEVICE double (set to 0 by the caller): receives the CFL reduction
 *              (ComputeDtFunctor) of the NEW state, fused in the epilogue
 *   d_skip     optional DEVICE int: when *d_skip != 0 the launch does nothing (lets a device-resident
 *              loop run past its own end without a host round trip)
 */
int e2d_k_fused_step(const e2d_params * p, const double * Uin, double * Uout, int jsize_loc, double dt,
                     const double * d_dt, double * d_invdt, const int * d_skip, void * stream);

/* per-cell device functions of HydroBaseFunctor (src/HydroBaseFunctor.h) evaluated on the GPU over
 * n records of HOST doubles — function-level known-answer tests.
 *   func     in (doubles per record)                          out
 *   "prim"   u[4]                                             q[4], c              computePrimitives :76-102
 *   "slope"  q, qPlusX, qMinusX, qPlusY, qMinusY (20)         dqX[4], dqY[4]       slope_unsplit_hydro_2d :473-516
 *   "trace"  q, dqX, dqY, dtdx, dtdy (14)                     XMIN,XMAX,YMIN,YMAX  trace_unsplit_2d_along_dir :214-291
 *   "hllc"   qleft, qright (8)                                flux[4]              riemann_hllc :704-809
 *   "approx" qleft, qright (8)                                qgdnv[4], flux[4]    riemann_approx :558-693
 *   "cmpflx" qgdnv (4)                                        flux[4]              cmpflx :523-547
 *   "hll"    qleft, qright (8)                                flux[4]              (extension: not in the reference)
 * and the lean-but-exact forms the fused kernel uses (csrc/e2d_lean.cuh), each with its fast-path guard (1 = the
 * fast path was accepted, 0 = the value comes from the plain-operator fallback):
 *   "hllc_lean"   qleft, qright (8)                           flux[4], guard
 *   "cell_lean"   u[4]                                        q[4], CFL integrand (c+|u|)/dx+(c+|v|)/dy, guard
 *   "trace_lean"  q, qPlusX, qMinusX, qPlusY, qMinusY, dtdx, dtdy (22)   dqX, dqY, XMIN, XMAX, YMIN, YMAX (24), guard
 *   "div"         a, d                                        shared-reciprocal a/d (zero numerators allowed), guard,
 *                                                             a/d by the `/` operator, same without zeros, guard
 *   "sqrt"        x                                           fast-path sqrt, guard, sqrt(x)
 */
int e2d_k_eval_host(const e2d_params * p, const char * func, const double * in, double * out, long n);

/* ------------------------------------------------------------------------------------------ */
/* HydroRun handle (replaces class euler2d::HydroRun<device_t>, src/HydroRun.h:44-399)        */
/* ------------------------------------------------------------------------------------------ */
typedef struct e2d_handle e2d_handle;

/* which array */
enum { E2D_U = 0, E2D_U2 = 1, E2D_Q = 2 };

/* y-slab description for multi-GPU runs; NULL in e2d_create = whole domain on the current device */
typedef struct e2d_slab
{
  int rank, nranks;
  int ny_loc; /* interior rows owned by this rank */
  int j_off;  /* global row index (ghosts included) of local row 0 */
} e2d_slab;

/* HydroRun::HydroRun (src/HydroRun.h:143-216): allocate U, U2 (+ scratch), run the problem
 * initialiser, U2 = U.  Uses the CURRENT CUDA device and creates its own stream unless `stream`
 * is given.  U_ext/U2_ext: optional caller-owned device buffers (isize*jsize_loc*4 doubles). */
int e2d_create(const e2d_params * p, const e2d_slab * slab, double * U_ext, double * U2_ext, void * stream,
               e2d_handle ** out);
int e2d_destroy(e2d_handle * h);

/* Sedov blast (problem=blast with total_energy_inside > 0) on y-slabs.  InitBlastFunctor::apply
 * (src/HydroRunFunctors.h:1445-1463) sets the energy inside the disc to E_tot / volume_inside, where volume_inside is
 * a Kokkos::Sum over the WHOLE grid.  e2d_create on a slab counts the disc cells of the rows the rank owns and leaves
 * the handle "pending": every compute entry point refuses it until e2d_blast_renormalise is given the count summed
 * over all ranks (an integer all-reduce, hence exact and identical to the single-domain value).  PeerSlabRun does that
 * with torch.distributed; e2d_peer_connect_local does it for the handles of one process.  No-ops otherwise. */
int e2d_blast_inside_count(e2d_handle * h, unsigned long long * n_local, int * pending /* may be NULL */);
int e2d_blast_renormalise(e2d_handle * h, unsigned long long n_inside_global);

/* HydroRun::compute_dt(useU) (src/HydroRun.h:229-251): *dt = cfl / max(invDt). Synchronous.
 * On a slab, *invdt_local (may be NULL) returns this rank's partial max so the caller can
 * allreduce(max) it and form dt = cfl/invDt itself.
 * When the array was last written by e2d_godunov_unsplit with implementationVersion 2, the reduction has already
 * been folded into that step's kernel (same integrand per cell, same bits) and only 8 bytes are fetched.  The
 * shortcut is never taken for caller-owned arrays (U_ext/U2_ext) nor once e2d_device_ptr has handed a pointer out. */
int e2d_compute_dt(e2d_handle * h, int useU, double * dt, double * invdt_local);
/* HydroRun::make_boundaries(Udata) (src/HydroRun.h:390-399); which = E2D_U | E2D_U2.
 * On a slab only the faces this rank owns are filled (x always; ymin on rank 0; ymax on the last). */
int e2d_make_boundaries(e2d_handle * h, int which);
/* HydroRun::godunov_unsplit(nStep, dt) (src/HydroRun.h:259-364), honouring implementationVersion:
 * 0 = store fluxes then update, 1 = slopes array + per-direction trace/flux/update, 2 = fused kernel. */
int e2d_godunov_unsplit(e2d_handle * h, int nStep, double dt);
/* as above but skipping the make_boundaries(data_in) of src/HydroRun.h:296 — for slab runs where the
 * caller has exchanged halos and filled boundaries itself */
int e2d_godunov_unsplit_nobc(e2d_handle * h, int nStep, double dt);

typedef struct e2d_run_stats
{
  int    nStep;        /* steps taken so far (total, across calls) */
  double t;            /* simulation time reached */
  double dt_last;      /* dt of the last step */
  double seconds;      /* device time of this call's steps (CUDA events on the handle's stream) */
  long long launches;  /* kernels launched by this call */
  double seconds_step_kernel; /* device time spent inside the fused step kernel alone (sum over its launches,
                                 one CUDA-event pair per launch on the handle's stream); 0 unless
                                 e2d_enable_timers(h, 1) */
} e2d_run_stats;

/* The main loop of src/main.cpp:100-143 with IO off, device-resident: dt, t and nStep live in
 * device memory, one fused kernel per step (boundary fill, step, next-step CFL reduction), no
 * host synchronisation inside.  Continues from the handle's current (t, nStep); stops when
 * t >= tEnd or nStep >= max_steps (max_steps < 0: params.nStepmax).  Whole-domain handles, or slabs whose peers
 * are connected (below). */
int e2d_run(e2d_handle * h, long max_steps, e2d_run_stats * stats);
/* ---- multi-GPU: e2d_run on y-slabs, halo rows and CFL partials as plain stores into peer memory over NVLink ----
 * Each rank creates its slab handle (e2d_create with an e2d_slab; the library cudaMalloc's U/U2), the ranks connect
 * to each other ONCE, then every rank calls e2d_run with the same max_steps: per step the ranks exchange the two
 * boundary rows and their invDt partial by direct stores + system-scope flags (no NCCL call, no host round trip),
 * and the result is bitwise identical to the single-GPU run (csrc/e2d_slab.cu).
 *   one process per GPU (torchrun):  e2d_ipc_export -> all-gather the blobs by any means -> e2d_ipc_connect
 *   one process, several GPUs:       e2d_peer_connect_local (then drive each handle from its own host thread:
 *                                    e2d_run blocks while the ranks wait for each other on the device) */
#define E2D_IPC_HANDLE_BYTES 64
typedef struct e2d_ipc_blob
{
  unsigned char U[E2D_IPC_HANDLE_BYTES], U2[E2D_IPC_HANDLE_BYTES], comm[E2D_IPC_HANDLE_BYTES]; /* cudaIpcMemHandle_t */
  int           rank, nranks, ny_loc, device;
} e2d_ipc_blob;
int e2d_ipc_export(e2d_handle * h, e2d_ipc_blob * out);
int e2d_ipc_connect(e2d_handle * h, const e2d_ipc_blob * blobs /* nranks entries, rank order */, int nranks);
int e2d_peer_connect_local(e2d_handle ** handles /* nranks entries, rank order */, int nranks);

/* dt used by each step taken through e2d_run so far (n_cap entries max); returns count in *n */
int e2d_get_dt_history(e2d_handle * h, double * dts, long n_cap, long * n);
/* reset (t, nStep) bookkeeping of e2d_run, e.g. after e2d_upload */
int e2d_set_time(e2d_handle * h, double t, int nStep);

/* Kokkos::deep_copy(Uhost, Udata) + raw access (src/HydroRun.h:522) */
int e2d_download(e2d_handle * h, int which, double * host, int layout);
int e2d_upload(e2d_handle * h, int which, const double * host, int layout);
double * e2d_device_ptr(e2d_handle * h, int which);
void *   e2d_stream(e2d_handle * h);
int      e2d_synchronize(e2d_handle * h);
int      e2d_get_params(e2d_handle * h, e2d_params * out);

/* End-to-end entry for callers whose state lives in HOST memory (bench.py's e2e leg):
 * H2D of U_host_in (SoA, whole slab incl. ghosts) -> make_boundaries -> compute_dt -> one
 * godunov step -> D2H into U_host_out.  *dt_out receives the dt used.  Pinned host buffers make the
 * copies asynchronous; pageable ones work too.  The ghost cells of the result are filled from the NEW interior (a
 * self-consistent array); the reference's godunov_unsplit leaves the input's ghosts there (HydroRun.h:302), which
 * e2d_godunov_unsplit reproduces — this convenience entry point has no counterpart in the reference. */
int e2d_step_host(e2d_handle * h, const double * U_host_in, double * U_host_out, double * dt_out);

/* The same step for host-resident state, STREAMED: rows go host -> device in chunks of `chunk_rows` (<= 0: ny/32),
 * every chunk is advanced as soon as the rows above it have landed and travels back while later chunks are still
 * arriving — H2D, the fused step and D2H overlap.  That requires dt up front, which is the reference's own call
 * structure  dt = compute_dt(nStep % 2); godunov_unsplit(nStep, dt)  (src/main.cpp:128,139, src/HydroRun.h:259):
 *   dt_in > 0   the step uses it (pass the previous call's *dt_next; multi-GPU: the min over the ranks' values);
 *   dt_in <= 0  dt is computed from the uploaded state first (only the way back then overlaps the compute).
 * *dt_used receives the dt of this step, *dt_next = cfl / max invDt of the state just written = what compute_dt
 * returns for it (the reduction rides on the step).  Ghost cells: physical faces are filled on the device (input
 * and result); on a slab handle the ghost rows at slab interfaces are taken from U_host_in as they are and not
 * written to U_host_out — the caller exchanges them. */
int e2d_step_host_streamed(e2d_handle * h, const double * U_host_in, double * U_host_out, double dt_in, int chunk_rows,
                           double * dt_used, double * dt_next);

/* A time march whose state lives in HOST memory (the loop of src/main.cpp:100-143 for a caller that keeps its arrays on
 * the host): nsteps steps, the state ping-ponging between two host buffers of e2d layout [var][j][i] — step s reads
 * buf_a and writes buf_b when s is even, the other way round when odd, so the result is in buf_a for an even nsteps
 * and in buf_b for an odd one.  Every step moves the whole state host -> device and back, chunked by rows like
 * e2d_step_host_streamed, and the steps are PIPELINED: the upload of step s+1 follows the upload of step s without a
 * gap, each chunk as soon as the same rows of step s have landed in the host buffer, so both PCIe directions stay
 * busy across steps (a per-step call pays the head and the tail of its pipeline every time).  dt = cfl / max invDt is
 * formed on the device between the steps (HydroRun.h:246) and never visits the host; dts (may be NULL) receives the dt of
 * every step afterwards, *t_io (may be NULL) is advanced by their sum.  params.tEnd is not looked at: the caller chooses
 * nsteps.  Pinned host buffers are needed for the copies to be asynchronous.  Whole-domain handles without a periodic y
 * direction only (E2D_ERR_UNSUPPORTED otherwise: use e2d_step_host_streamed).  Results are bit-identical to e2d_run.
 * The handle's own device arrays are staging space here (as for e2d_step_host*): afterwards they do not hold the
 * result — e2d_upload it before continuing with the device-resident entry points. */
int e2d_march_host(e2d_handle * h, double * buf_a, double * buf_b, long nsteps, int chunk_rows, double * dts,
                   double * t_io);

/* HydroRun::saveData -> saveVTK (src/HydroRun.h:486-609): ascii .vti, ghosts stripped,
 * <outputDir>/<outputPrefix>_<%07d iStep>.vti, default ostream precision (6 significant digits): byte-identical to
 * the reference program's files (tests/test_gpu_refmain.py).  Whole-domain handles only: on a y-slab handle both VTK
 * writers return E2D_ERR_UNSUPPORTED (every rank would write its piece under the same name) — gather the interior on
 * one rank instead (e2d_save_raw writes a slab's own rows and is allowed). */
int e2d_save_vtk(e2d_handle * h, int which, int iStep);

/* Fast output — same file name, names (rho, E, mx, my: src/HydroParams.cpp:17), extents,
 * origin/spacing and cell order (ghosts stripped, i fastest) as HydroRun::saveVTK (src/HydroRun.h:507-609), but
 * the four arrays are Float64 `format="appended"` raw binary (header_type UInt64): bit-exact values instead of 6
 * significant digits.  The interior is gathered on the device into dense row blocks which travel D2H into pinned
 * double buffers while the previous block is being written to the file.  e2d_save_vtk dispatches here when
 * params.vtkAppended is set. */
int e2d_save_vtk_appended(e2d_handle * h, int which, int iStep);
/* raw snapshot of the interior: nx*ny doubles per variable, [var][j][i], no header (same pipeline) */
int e2d_save_raw(e2d_handle * h, int which, const char * path);

/* Sedov post-processing — replaces ComputeRadialProfileFunctor::apply (src/ComputeRadialProfileFunctor.h:86-140,
 * called by src/main.cpp:175-179 with hydro->U).  Every cell of the array `which` incl. ghost cells is binned by
 * the distance of its centre from the box centre; nbins <= 0: params.blast_nbins.  Outputs (host, nbins entries
 * each, any may be NULL): distances[k] = (k + 0.5) * max_radial_distance / nbins, sums[k] = sum of rho,
 * counts[k] = number of cells; the profile the reference saves is sums[k] / counts[k].  Deterministic (no
 * atomics); samples the reference bins out of bounds (corner ghost cells) are dropped.  On a slab handle the
 * sums / counts cover the rows this rank owns (interior rows + the ghost rows of the physical y faces): add them
 * over the ranks. */
int e2d_compute_radial_profile(e2d_handle * h, int which, int nbins, double * distances, double * sums, int * counts);
/* apply() including its output: <dir>/sedov_blast_radial_distances.npy and <dir>/sedov_blast_density_profile.npy
 * (cnpy::npy_save, src/cnpy/cnpy_io.h:36-63); dir NULL or "": the current directory, like the reference. */
int e2d_save_radial_profile(e2d_handle * h, int which, const char * dir);
/* a 1-D float64 array as NumPy .npy version 1.0 (what cnpy::npy_save writes for a rank-1 double view) */
int e2d_save_npy(const char * path, const double * data, long n);

/* the five public timers of HydroRun (src/HydroRun.h:74-75), seconds accumulated by the
 * e2d_godunov_unsplit path when timing is enabled: [boundaries, godunov, primitive, fluxes, update] */
int e2d_enable_timers(e2d_handle * h, int on);
int e2d_get_timers(e2d_handle * h, double out[5]);

/* Profiling regions: Kokkos::Profiling::pushRegion / popRegion (src/HydroRun.h:242-360, src/main.cpp:93,111) as
 * named NVTX ranges.  Off by default; on with e2d_profile_enable(1) or the environment variable E2D_PROFILE=1 (read
 * once).  In profile mode the library wraps its own phases with the reference's region names — compute_dt,
 * make_boundaries, compute_primitives, hydro_impl0 (compute_fluxes, update_hydro nested inside), hydro_impl1,
 * hydro_impl2 — and a host program can add its own (main_loop, output) through push / pop.  e2d_profile_stats
 * reports how many ranges were opened and the current nesting depth (tests; a balanced program ends at depth 0). */
int  e2d_profile_enable(int on);
void e2d_profile_push(const char * name);
void e2d_profile_pop(void);
int  e2d_profile_stats(unsigned long long * ranges_opened, int * depth);

#ifdef __cplusplus
}
#endif
#endif /* EULER2D_B200_H */
