// C ABI of euler2d_b200 (include/euler2d_b200.h): the HydroRun handle, the device-resident time
// loop and thin wrappers over the kernel launchers.  No torch types, no exceptions across the ABI.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <fstream>
#include <new>
#include <sstream>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges cost nothing unless a tool is attached

#include "e2d_internal.h"

namespace
{

thread_local std::string               g_last_error;
std::atomic<unsigned long long>        g_launches{ 0 };

int
fail_cuda(cudaError_t e, const char * what)
{
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return E2D_ERR_CUDA;
}

#define E2D_CUDA(call)                      \
  do                                        \
  {                                         \
    cudaError_t e2d_err_ = (call);          \
    if (e2d_err_ != cudaSuccess)            \
      return fail_cuda(e2d_err_, #call);    \
  } while (0)

int
fail(int status, const std::string & msg)
{
  g_last_error = msg;
  return status;
}

} // namespace

namespace e2d
{
void
count_launch(int n)
{
  g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed);
}
} // namespace e2d

using namespace e2d;

// ---- profiling regions (NVTX): Kokkos::Profiling::pushRegion / popRegion of the reference ----
namespace
{
std::atomic<int>                g_profile{ -1 }; // -1: not decided yet (E2D_PROFILE is read on first use)
std::atomic<unsigned long long> g_profile_ranges{ 0 };
thread_local int                t_profile_depth = 0;

bool
profiling()
{
  int v = g_profile.load(std::memory_order_relaxed);
  if (v < 0)
  {
    const char * e = std::getenv("E2D_PROFILE");
    v = (e && *e && std::strcmp(e, "0") != 0) ? 1 : 0;
    g_profile.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

struct ProfileRegion
{
  bool on;
  explicit ProfileRegion(const char * name)
    : on(profiling())
  {
    if (on)
    {
      nvtxRangePushA(name);
      ++t_profile_depth;
      g_profile_ranges.fetch_add(1, std::memory_order_relaxed);
    }
  }
  ~ProfileRegion()
  {
    if (on)
    {
      nvtxRangePop();
      --t_profile_depth;
    }
  }
};
} // namespace

struct e2d_handle
{
  e2d_params   p;
  e2d_slab     slab;
  bool         whole;
  Geom         g;
  size_t       n; // doubles per array
  cudaStream_t stream = nullptr;
  bool         own_stream = false;
  double *     U = nullptr;
  double *     U2 = nullptr;
  bool         own_U = false, own_U2 = false;
  double *     Q = nullptr;  // impl 0/1
  double *     Fx = nullptr; // impl 0/1
  double *     Fy = nullptr; // impl 0
  double *     Sx = nullptr; // impl 1
  double *     Sy = nullptr; // impl 1
  // scalars
  unsigned long long * d_bits = nullptr; // scratch for compute_dt
  // compute_dt cache: godunov_unsplit (implementationVersion 2) folds the CFL reduction of the state it writes into
  // the fused kernel (the same per-cell integrand, bit for bit), so the compute_dt that follows — the reference's
  // call pattern, main.cpp:128-139 — only fetches 8 bytes instead of re-reading the array.  Valid only while the
  // library is the sole writer: every entry point that writes an array clears it; it is never used for caller-owned
  // arrays (U_ext / U2_ext) nor after e2d_device_ptr has handed a pointer out.
  unsigned long long * d_cfl = nullptr; // [2]: invDt bit patterns of U, U2
  bool                 cfl_valid[2] = { false, false };
  bool                 cfl_cache_ok = true;
  LoopState *          d_loop = nullptr;
  LoopState *          h_loop = nullptr; // pinned mirror
  double *             d_hist = nullptr;
  long                 hist_cap = 0;
  bool                 loop_primed = false; // d_loop->invdt_cur valid for the current state
  // the slab loop (e2d_slab.cu)
  int                  device = 0;
  SlabComm *           d_comm = nullptr;
  SlabState *          d_state = nullptr;
  SlabState *          h_state = nullptr; // pinned mirror
  unsigned long long   seq = 0;           // steps issued through the slab loop so far (flags carry seq)
  bool                 seq_poisoned = false; // a wait for a peer timed out: flags and state are no longer trustworthy
  // Sedov init on a slab: disc cells of the rows this rank owns; the state is unusable until e2d_blast_renormalise
  unsigned long long   blast_inside_local = 0;
  bool                 blast_pending = false;
  struct
  {
    bool       connected = false;
    int        lower = -1, upper = -1; // neighbour ranks
    double *   lowerU[2] = { nullptr, nullptr }; // the neighbours' U / U2 as peer pointers
    double *   upperU[2] = { nullptr, nullptr };
    int        lower_jsize = 0, upper_jsize = 0;
    SlabComm * comm[kMaxRanks] = {};
    std::vector<void *> ipc_opened;
  } peers;
  double               t = 0.0;
  int                  nStep = 0;
  double               dt_last = 0.0;
  // timers: boundaries, godunov, primitive, fluxes, update
  bool        timing = false;
  double      timers[5] = { 0, 0, 0, 0, 0 };
  cudaEvent_t ev[2] = { nullptr, nullptr };
  cudaEvent_t ev_t[5][2] = {}; // one event pair per timer: the godunov timer nests the others
  std::vector<cudaEvent_t> ev_step; // e2d_run profiling: one pair per step of a batch
  // e2d_step_host_streamed: copy streams and an event pool (created on first use)
  cudaStream_t             s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  // fast output: dense interior blocks, device + pinned double buffers
  double *    out_dev[2] = { nullptr, nullptr };
  double *    out_host[2] = { nullptr, nullptr };
  cudaEvent_t out_ev[2] = { nullptr, nullptr };
  size_t      out_cap = 0;          // doubles per buffer
  size_t      out_prefix_bytes = 0; // bytes in front of each variable block in the file
};

namespace
{

int
faces_for(const e2d_handle * h)
{
  if (h->whole)
    return E2D_FACES_ALL;
  int f = E2D_FACES_X;
  // a periodic y direction split over several ranks is closed by the halo exchange, not by a fill
  if (h->slab.rank == 0 && h->p.boundary_type_ymin != E2D_BC_PERIODIC)
    f |= E2D_FACES_YMIN;
  if (h->slab.rank == h->slab.nranks - 1 && h->p.boundary_type_ymax != E2D_BC_PERIODIC)
    f |= E2D_FACES_YMAX;
  if (h->slab.nranks == 1)
    f = E2D_FACES_ALL;
  return f;
}

double *
array_of(e2d_handle * h, int which)
{
  switch (which)
  {
    case E2D_U:
      return h->U;
    case E2D_U2:
      return h->U2;
    case E2D_Q:
      return h->Q;
  }
  return nullptr;
}

struct PhaseTimer
{
  e2d_handle * h;
  int          slot;
  PhaseTimer(e2d_handle * hh, int s)
    : h(hh)
    , slot(s)
  {
    if (h->timing)
      cudaEventRecord(h->ev_t[slot][0], h->stream);
  }
  ~PhaseTimer()
  {
    if (h->timing)
    { // like the reference's CudaTimer::stop (src/CudaTimer.h:54-62): record + synchronize
      cudaEventRecord(h->ev_t[slot][1], h->stream);
      cudaEventSynchronize(h->ev_t[slot][1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, h->ev_t[slot][0], h->ev_t[slot][1]);
      h->timers[slot] += ms * 1e-3;
    }
  }
};

int
ensure_scratch(e2d_handle * h, int impl)
{
  const size_t bytes = h->n * sizeof(double);
  if ((impl == 0 || impl == 1) && !h->Q)
  {
    E2D_CUDA(cudaMalloc(&h->Q, bytes));
    E2D_CUDA(cudaMemsetAsync(h->Q, 0, bytes, h->stream));
  }
  if ((impl == 0 || impl == 1) && !h->Fx)
  {
    E2D_CUDA(cudaMalloc(&h->Fx, bytes));
    E2D_CUDA(cudaMemsetAsync(h->Fx, 0, bytes, h->stream));
  }
  if (impl == 0 && !h->Fy)
  {
    E2D_CUDA(cudaMalloc(&h->Fy, bytes));
    E2D_CUDA(cudaMemsetAsync(h->Fy, 0, bytes, h->stream));
  }
  if (impl == 1 && !h->Sx)
  {
    E2D_CUDA(cudaMalloc(&h->Sx, bytes));
    E2D_CUDA(cudaMalloc(&h->Sy, bytes));
    E2D_CUDA(cudaMemsetAsync(h->Sx, 0, bytes, h->stream));
    E2D_CUDA(cudaMemsetAsync(h->Sy, 0, bytes, h->stream));
  }
  return E2D_OK;
}

// godunov_unsplit_impl (src/HydroRun.h:281-364)
int
godunov_impl(e2d_handle * h, double * in, double * out, double dt, bool do_bc)
{
  const e2d_params & p = h->p;
  const double       dtdx = dt / p.dx; // :290-291
  const double       dtdy = dt / p.dy;
  cudaStream_t       st = h->stream;

  if (do_bc)
  {
    ProfileRegion pr("make_boundaries"); // :294
    PhaseTimer    tb(h, 0);
    E2D_CUDA(launch_make_boundaries(p, h->g, in, faces_for(h), nullptr, st)); // :296
  }
  const int  impl = p.implementationVersion;
  const bool fused = impl == 2 || !p.unfusedKernels;
  if (!fused)
    if (int rc = ensure_scratch(h, impl))
      return rc;

  PhaseTimer tg(h, 1);
  const int w_out = (out == h->U) ? 0 : 1;
  h->cfl_valid[w_out] = false;
  if (fused)
  {
    // fused: no deep_copy, no Q array (the reference's impl 2 keeps both, :302,:309,:359); the CFL reduction of the
    // new state rides along for the next compute_dt (see e2d_handle::d_cfl).
    // Implementations 0 and 1 (bit-identical to each other in the reference) produce this very interior and leave
    // in's ghost cells in out (deep_copy, :302): the fused step + a copy of the ghost frame is the same array.
    unsigned long long * cfl = h->cfl_cache_ok ? h->d_cfl + w_out : nullptr;
    if (cfl)
      E2D_CUDA(cudaMemsetAsync(cfl, 0, sizeof(unsigned long long), st));
    ProfileRegion pi(impl == 0 ? "hydro_impl0" : (impl == 1 ? "hydro_impl1" : "hydro_impl2")); // :315,:335,:357
    if (impl == 0)
    {
      ProfileRegion pf("compute_fluxes"); // :317 (the fused kernel: primitives, fluxes and update in one launch)
      PhaseTimer    tf(h, 3); // the reference times its flux kernel for implementation 0 only (:316-320)
      E2D_CUDA(launch_fused_step(p, h->g, in, out, dt, nullptr, cfl, nullptr, st));
    }
    else
      E2D_CUDA(launch_fused_step(p, h->g, in, out, dt, nullptr, cfl, nullptr, st));
    if (impl != 2)
      E2D_CUDA(launch_copy_ghost_frame(h->g, in, out, st));
    h->cfl_valid[w_out] = cfl != nullptr;
    return E2D_OK;
  }
  E2D_CUDA(cudaMemcpyAsync(out, in, h->n * sizeof(double), cudaMemcpyDeviceToDevice, st)); // :302
  {
    ProfileRegion pp("compute_primitives"); // :308
    PhaseTimer    tp(h, 2);
    E2D_CUDA(launch_convert_to_primitives(p, h->g, in, h->Q, st)); // :309
  }
  if (impl == 0)
  {
    ProfileRegion pi("hydro_impl0"); // :315
    {
      ProfileRegion pf("compute_fluxes"); // :317
      PhaseTimer    tf(h, 3);
      E2D_CUDA(launch_compute_and_store_fluxes(p, h->g, h->Q, h->Fx, h->Fy, dtdx, dtdy, st)); // :319
    }
    {
      ProfileRegion pu("update_hydro"); // :324
      PhaseTimer    tu(h, 4);
      E2D_CUDA(launch_update(p, h->g, out, h->Fx, h->Fy, st)); // :326
    }
  }
  else
  { // :338-352
    ProfileRegion pi("hydro_impl1"); // :335
    E2D_CUDA(launch_compute_slopes(p, h->g, h->Q, h->Sx, h->Sy, st));
    E2D_CUDA(launch_trace_and_fluxes(p, h->g, h->Q, h->Sx, h->Sy, h->Fx, dtdx, dtdy, 1, st));
    E2D_CUDA(launch_update_dir(p, h->g, out, h->Fx, 1, st));
    E2D_CUDA(launch_trace_and_fluxes(p, h->g, h->Q, h->Sx, h->Sy, h->Fx, dtdx, dtdy, 2, st));
    E2D_CUDA(launch_update_dir(p, h->g, out, h->Fx, 2, st));
  }
  return E2D_OK;
}

void
soa_from_kokkos_omp(const double * src, double * dst, int isize, int jsize)
{
  for (int v = 0; v < 4; ++v)
    for (int j = 0; j < jsize; ++j)
      for (int i = 0; i < isize; ++i)
        dst[(size_t)i + (size_t)isize * ((size_t)j + (size_t)jsize * v)] = src[((size_t)i * jsize + j) * 4 + v];
}

void
soa_to_kokkos_omp(const double * src, double * dst, int isize, int jsize)
{
  for (int v = 0; v < 4; ++v)
    for (int j = 0; j < jsize; ++j)
      for (int i = 0; i < isize; ++i)
        dst[((size_t)i * jsize + j) * 4 + v] = src[(size_t)i + (size_t)isize * ((size_t)j + (size_t)jsize * v)];
}

int
check_slab_args(const e2d_params * p, int jsize_loc)
{
  if (!p)
    return fail(E2D_ERR_INVALID, "params is NULL");
  // nx, ny >= 2: with a single interior cell the reference's boundary passes read ghost cells written by
  // an earlier pass (src/HydroRunFunctors.h:1895,1933), which the single-launch fill does not reproduce
  if (p->nx < 2 || p->isize != p->nx + 4 || jsize_loc < 6 || p->ghostWidth != 2)
    return fail(E2D_ERR_INVALID, "inconsistent geometry (need nx>=2, isize=nx+4, jsize_loc>=6, ghostWidth=2)");
  // the kernels index one variable plane with 32-bit offsets (the reference does too: `int ijsize`, HydroRun.h:511)
  if ((long long)p->isize * (long long)jsize_loc >= (1ll << 31))
    return fail(E2D_ERR_UNSUPPORTED, "slab too large: isize*jsize_loc must stay below 2^31 cells per device");
  return E2D_OK;
}

} // namespace

// ==========================================================================================
extern "C"
{

  const char *
  e2d_version(void)
  {
    return "euler2d_b200 0.1 (sm_100a, strict fp64)";
  }

  const char *
  e2d_status_string(int status)
  {
    switch (status)
    {
      case E2D_OK:
        return "ok";
      case E2D_ERR_INVALID:
        return "invalid argument";
      case E2D_ERR_IO:
        return "i/o error";
      case E2D_ERR_CUDA:
        return "CUDA error";
      case E2D_ERR_ALLOC:
        return "allocation failed";
      case E2D_ERR_UNSUPPORTED:
        return "unsupported";
    }
    return "unknown status";
  }

  const char *
  e2d_last_error(void)
  {
    return g_last_error.c_str();
  }

  int
  e2d_device_count(void)
  {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
    return n;
  }

  unsigned long long
  e2d_kernel_launch_count(void)
  {
    return g_launches.load();
  }

  // ---------------------------------------------------------------- kernel-level entry points
  int
  e2d_k_init_problem(const e2d_params * p, double * U, int jsize_loc, int j_off, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    if (p->problemType == E2D_PROBLEM_BLAST && p->blast_total_energy_inside > 0 &&
        (jsize_loc != p->jsize || j_off != 0))
      return fail(E2D_ERR_UNSUPPORTED, "energy-renormalised blast init needs the whole domain on one device");
    // Sedov (blast with total_energy_inside): the energy inside the disc depends on a count over the WHOLE grid
    // (src/HydroRunFunctors.h:1445-1463); a slab cannot know it from its own rows.  Handles do it with
    // e2d_blast_inside_count / e2d_blast_renormalise; this stateless entry point refuses.
    if (p->problemType == E2D_PROBLEM_BLAST && p->blast_total_energy_inside > 0 && (jsize_loc != p->jsize || j_off != 0))
      return fail(E2D_ERR_UNSUPPORTED, "e2d_k_init_problem: the energy-renormalised blast init needs the whole domain; on "
                                       "slabs use e2d_create + e2d_blast_inside_count / e2d_blast_renormalise");
    E2D_CUDA(launch_init_problem(*p, make_geom(*p, jsize_loc, j_off), U, (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_make_boundaries(const e2d_params * p, double * U, int jsize_loc, int faces, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    E2D_CUDA(launch_make_boundaries(*p, make_geom(*p, jsize_loc, 0), U, faces, nullptr, (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_reduce_invdt(const e2d_params * p, const double * U, int jsize_loc, double * d_invdt, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    E2D_CUDA(launch_reduce_invdt(*p, make_geom(*p, jsize_loc, 0), U, reinterpret_cast<unsigned long long *>(d_invdt),
                                 (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_convert_to_primitives(const e2d_params * p, const double * U, double * Q, int jsize_loc, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    E2D_CUDA(launch_convert_to_primitives(*p, make_geom(*p, jsize_loc, 0), U, Q, (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_compute_and_store_fluxes(const e2d_params * p, const double * Q, double * Fx, double * Fy, double dtdx,
                                 double dtdy, int jsize_loc, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    E2D_CUDA(launch_compute_and_store_fluxes(*p, make_geom(*p, jsize_loc, 0), Q, Fx, Fy, dtdx, dtdy,
                                             (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_update(const e2d_params * p, double * U, const double * Fx, const double * Fy, int jsize_loc, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    E2D_CUDA(launch_update(*p, make_geom(*p, jsize_loc, 0), U, Fx, Fy, (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_compute_slopes(const e2d_params * p, const double * Q, double * Sx, double * Sy, int jsize_loc,
                       void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    E2D_CUDA(launch_compute_slopes(*p, make_geom(*p, jsize_loc, 0), Q, Sx, Sy, (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_compute_trace_and_fluxes(const e2d_params * p, const double * Q, const double * Sx, const double * Sy,
                                 double * F, double dtdx, double dtdy, int dir, int jsize_loc, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    if (dir != 1 && dir != 2)
      return fail(E2D_ERR_INVALID, "dir must be 1 (XDIR) or 2 (YDIR)");
    E2D_CUDA(launch_trace_and_fluxes(*p, make_geom(*p, jsize_loc, 0), Q, Sx, Sy, F, dtdx, dtdy, dir,
                                     (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_update_dir(const e2d_params * p, double * U, const double * F, int dir, int jsize_loc, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    if (dir != 1 && dir != 2)
      return fail(E2D_ERR_INVALID, "dir must be 1 (XDIR) or 2 (YDIR)");
    E2D_CUDA(launch_update_dir(*p, make_geom(*p, jsize_loc, 0), U, F, dir, (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_fused_step(const e2d_params * p, const double * Uin, double * Uout, int jsize_loc, double dt,
                   const double * d_dt, double * d_invdt, const int * d_skip, void * stream)
  {
    if (int rc = check_slab_args(p, jsize_loc))
      return rc;
    if (Uin == Uout)
      return fail(E2D_ERR_INVALID, "the fused step is out of place: Uin and Uout must differ");
    E2D_CUDA(launch_fused_step(*p, make_geom(*p, jsize_loc, 0), Uin, Uout, dt, d_dt,
                               reinterpret_cast<unsigned long long *>(d_invdt), d_skip, (cudaStream_t)stream));
    return E2D_OK;
  }

  int
  e2d_k_eval_host(const e2d_params * p, const char * func, const double * in, double * out, long n)
  {
    if (!p || !func || !in || !out || n < 0)
      return fail(E2D_ERR_INVALID, "bad argument");
    static const struct
    {
      const char * name;
      int          id, nin, nout;
    } table[] = { { "prim", 0, 4, 5 },   { "slope", 1, 20, 8 }, { "trace", 2, 14, 16 }, { "hllc", 3, 8, 4 },
                  { "approx", 4, 8, 8 }, { "cmpflx", 5, 4, 4 }, { "hll", 6, 8, 4 },     { "hllc_lean", 7, 8, 5 },
                  { "cell_lean", 8, 4, 6 }, { "trace_lean", 9, 22, 25 }, { "div", 10, 2, 5 }, { "sqrt", 11, 1, 3 },
                  { "fast_div", 12, 2, 2 }, { "fast_sqrt", 13, 1, 2 }, { "fast_hllc", 14, 8, 4 },
                  { "fast_cell", 15, 4, 5 }, { "fast_slope", 16, 20, 8 }, { "fast_trace", 17, 14, 16 },
                  { "rusanov", 18, 8, 4 } };
    int id = -1, nin = 0, nout = 0;
    for (const auto & e : table)
      if (!std::strcmp(e.name, func))
        id = e.id, nin = e.nin, nout = e.nout;
    if (id < 0)
      return fail(E2D_ERR_INVALID, std::string("unknown function ") + func);
    if (n == 0)
      return E2D_OK;
    double *d_in = nullptr, *d_out = nullptr;
    E2D_CUDA(cudaMalloc(&d_in, sizeof(double) * nin * n));
    E2D_CUDA(cudaMalloc(&d_out, sizeof(double) * nout * n));
    E2D_CUDA(cudaMemcpy(d_in, in, sizeof(double) * nin * n, cudaMemcpyHostToDevice));
    cudaError_t e = launch_eval(*p, id, d_in, d_out, n, nullptr);
    if (e == cudaSuccess)
      e = cudaMemcpy(out, d_out, sizeof(double) * nout * n, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess)
      return fail_cuda(e, "e2d_k_eval_host");
    return E2D_OK;
  }

  // ---------------------------------------------------------------- HydroRun handle
  int
  e2d_create(const e2d_params * p, const e2d_slab * slab, double * U_ext, double * U2_ext, void * stream,
             e2d_handle ** out)
  {
    if (!p || !out)
      return fail(E2D_ERR_INVALID, "bad argument");
    *out = nullptr;
    if (e2d_device_count() < 1)
      return fail(E2D_ERR_CUDA, "no CUDA device: euler2d_b200 has no CPU fallback");
    e2d_handle * h = new (std::nothrow) e2d_handle();
    if (!h)
      return fail(E2D_ERR_ALLOC, "out of host memory");
    h->p = *p;
    if (slab)
    {
      h->slab = *slab;
      h->whole = (slab->nranks == 1);
    }
    else
    {
      h->slab.rank = 0;
      h->slab.nranks = 1;
      h->slab.ny_loc = p->ny;
      h->slab.j_off = 0;
      h->whole = true;
    }
    const int jsize_loc = h->slab.ny_loc + 2 * p->ghostWidth;
    if (int rc = check_slab_args(p, jsize_loc))
    {
      delete h;
      return rc;
    }
    h->g = make_geom(*p, jsize_loc, h->slab.j_off);
    h->n = (size_t)p->isize * jsize_loc * 4;
    int rc = E2D_OK;
    do
    {
#define E2D_TRY(call)                        \
  {                                          \
    cudaError_t e_ = (call);                 \
    if (e_ != cudaSuccess)                   \
    {                                        \
      rc = fail_cuda(e_, #call);             \
      break;                                 \
    }                                        \
  }
      if (stream)
        h->stream = (cudaStream_t)stream;
      else
      {
        E2D_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
      }
      if (U_ext)
        h->U = U_ext;
      else
      {
        E2D_TRY(cudaMalloc(&h->U, h->n * sizeof(double)));
        h->own_U = true;
      }
      if (U2_ext)
        h->U2 = U2_ext;
      else
      {
        E2D_TRY(cudaMalloc(&h->U2, h->n * sizeof(double)));
        h->own_U2 = true;
      }
      E2D_TRY(cudaMalloc(&h->d_bits, sizeof(unsigned long long)));
      E2D_TRY(cudaMalloc(&h->d_cfl, 2 * sizeof(unsigned long long)));
      h->cfl_cache_ok = !U_ext && !U2_ext;
      E2D_TRY(cudaMalloc(&h->d_loop, sizeof(LoopState)));
      E2D_TRY(cudaMallocHost(&h->h_loop, sizeof(LoopState)));
      E2D_TRY(cudaGetDevice(&h->device));
      E2D_TRY(cudaMalloc(&h->d_comm, sizeof(SlabComm)));
      E2D_TRY(cudaMemset(h->d_comm, 0, sizeof(SlabComm)));
      E2D_TRY(cudaMalloc(&h->d_state, sizeof(SlabState)));
      E2D_TRY(cudaMemset(h->d_state, 0, sizeof(SlabState)));
      E2D_TRY(cudaMallocHost(&h->h_state, sizeof(SlabState)));
      h->peers.comm[h->slab.rank < kMaxRanks ? h->slab.rank : 0] = h->d_comm;
      E2D_TRY(cudaEventCreate(&h->ev[0]));
      E2D_TRY(cudaEventCreate(&h->ev[1]));
      for (int k = 0; k < 5 && rc == E2D_OK; ++k)
        for (int e = 0; e < 2; ++e)
          if (cudaEventCreate(&h->ev_t[k][e]) != cudaSuccess)
            rc = fail(E2D_ERR_CUDA, "cudaEventCreate");
      if (rc != E2D_OK)
        break;
      // A periodic y direction split over several ranks is closed by the halo exchange between the first and the
      // last rank: that needs BOTH faces periodic (the reference reads each face on its own and would wrap one side
      // only, src/HydroRunFunctors.h:1968-2019 — a combination no deck uses and no rank pair could serve).
      if (!h->whole && (p->boundary_type_ymin == E2D_BC_PERIODIC) != (p->boundary_type_ymax == E2D_BC_PERIODIC))
      {
        rc = fail(E2D_ERR_UNSUPPORTED, "y-slabs need boundary_type_ymin and boundary_type_ymax both periodic or neither");
        break;
      }
      (void)refined_reciprocal(p->dx); // warm the cache: launches inside the loops must never synchronise
      (void)refined_reciprocal(p->dy);
      // HydroRun.h:185-214: initial condition, then U2 = U
      if (p->problemType == E2D_PROBLEM_BLAST && p->blast_total_energy_inside > 0 && !h->whole)
      {
        // Sedov on slabs: the energy inside the disc is E_tot / (volume of ALL disc cells), a reduction over the whole
        // grid in the reference (src/HydroRunFunctors.h:1445-1463).  This rank counts the disc cells of the rows it
        // owns; the initialisation is completed by e2d_blast_renormalise with the sum over the ranks (an integer
        // all-reduce: PeerSlabRun does it, e2d_peer_connect_local does it for handles of one process).
        const int jlo = h->slab.rank == 0 ? 0 : 2;
        const int jhi = h->slab.rank == h->slab.nranks - 1 ? jsize_loc : jsize_loc - 2;
        E2D_TRY(launch_init_problem(*p, h->g, h->U, h->stream, &h->blast_inside_local, -1, jlo, jhi));
        h->blast_pending = true;
      }
      else
        E2D_TRY(launch_init_problem(*p, h->g, h->U, h->stream));
      E2D_TRY(cudaMemcpyAsync(h->U2, h->U, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      E2D_TRY(cudaStreamSynchronize(h->stream));
#undef E2D_TRY
    } while (0);
    if (rc != E2D_OK)
    {
      e2d_destroy(h);
      return rc;
    }
    *out = h;
    return E2D_OK;
  }

  int
  e2d_destroy(e2d_handle * h)
  {
    if (!h)
      return E2D_OK;
    cudaSetDevice(h->device);
    if (h->stream)
      cudaStreamSynchronize(h->stream);
    if (h->own_U)
      cudaFree(h->U);
    if (h->own_U2)
      cudaFree(h->U2);
    cudaFree(h->Q);
    cudaFree(h->Fx);
    cudaFree(h->Fy);
    cudaFree(h->Sx);
    cudaFree(h->Sy);
    cudaFree(h->d_bits);
    cudaFree(h->d_cfl);
    cudaFree(h->d_loop);
    for (void * q : h->peers.ipc_opened)
      cudaIpcCloseMemHandle(q);
    cudaFree(h->d_comm);
    cudaFree(h->d_state);
    if (h->h_state)
      cudaFreeHost(h->h_state);
    cudaFree(h->d_hist);
    if (h->h_loop)
      cudaFreeHost(h->h_loop);
    if (h->ev[0])
      cudaEventDestroy(h->ev[0]);
    if (h->ev[1])
      cudaEventDestroy(h->ev[1]);
    for (int k = 0; k < 5; ++k)
      for (int e = 0; e < 2; ++e)
        if (h->ev_t[k][e])
          cudaEventDestroy(h->ev_t[k][e]);
    for (auto & e : h->ev_step)
      if (e)
        cudaEventDestroy(e);
    for (auto & e : h->ev_pool)
      if (e)
        cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b)
    {
      cudaFree(h->out_dev[b]);
      if (h->out_host[b])
        cudaFreeHost(h->out_host[b]);
      if (h->out_ev[b])
        cudaEventDestroy(h->out_ev[b]);
    }
    if (h->s_in)
      cudaStreamDestroy(h->s_in);
    if (h->s_out)
      cudaStreamDestroy(h->s_out);
    if (h->own_stream && h->stream)
      cudaStreamDestroy(h->stream);
    delete h;
    return E2D_OK;
  }

  int
  e2d_blast_inside_count(e2d_handle * h, unsigned long long * n_local, int * pending)
  {
    if (!h || !n_local)
      return fail(E2D_ERR_INVALID, "bad argument");
    *n_local = h->blast_inside_local;
    if (pending)
      *pending = h->blast_pending ? 1 : 0;
    return E2D_OK;
  }

  int
  e2d_blast_renormalise(e2d_handle * h, unsigned long long n_inside_global)
  {
    if (!h)
      return fail(E2D_ERR_INVALID, "bad argument");
    if (!h->blast_pending)
      return E2D_OK; // nothing to complete (whole domain, or not the Sedov problem)
    if (n_inside_global < h->blast_inside_local)
      return fail(E2D_ERR_INVALID, "the global count of disc cells cannot be smaller than this slab's own");
    cudaSetDevice(h->device);
    E2D_CUDA(launch_init_problem(h->p, h->g, h->U, h->stream, nullptr, (long long)n_inside_global));
    E2D_CUDA(cudaMemcpyAsync(h->U2, h->U, h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    E2D_CUDA(cudaStreamSynchronize(h->stream));
    h->blast_pending = false;
    return E2D_OK;
  }

#define E2D_REFUSE_PENDING(h)                                                                                         \
  if ((h)->blast_pending)                                                                                             \
    return fail(E2D_ERR_INVALID, "Sedov initialisation of this slab is incomplete: call e2d_blast_renormalise with the " \
                                 "disc-cell count summed over all ranks (e2d_blast_inside_count)")

  int
  e2d_compute_dt(e2d_handle * h, int useU, double * dt, double * invdt_local)
  {
    if (!h || !dt)
      return fail(E2D_ERR_INVALID, "bad argument");
    E2D_REFUSE_PENDING(h);
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    ProfileRegion  pr("compute_dt");               // HydroRun.h:242
    const double * A = (useU == 0) ? h->U : h->U2; // HydroRun.h:237-240
    const int      w = (useU == 0) ? 0 : 1;
    double         invDt = 0.0;
    if (h->cfl_cache_ok && h->cfl_valid[w])
      E2D_CUDA(cudaMemcpyAsync(&invDt, h->d_cfl + w, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    else
    {
      E2D_CUDA(cudaMemsetAsync(h->d_bits, 0, sizeof(unsigned long long), h->stream));
      E2D_CUDA(launch_reduce_invdt(h->p, h->g, A, h->d_bits, h->stream));
      E2D_CUDA(cudaMemcpyAsync(&invDt, h->d_bits, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    E2D_CUDA(cudaStreamSynchronize(h->stream));
    if (invdt_local)
      *invdt_local = invDt;
    *dt = h->p.cfl / invDt; // HydroRun.h:246
    return E2D_OK;
  }

  int
  e2d_make_boundaries(e2d_handle * h, int which)
  {
    if (!h || (which != E2D_U && which != E2D_U2))
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    ProfileRegion pr("make_boundaries");
    E2D_CUDA(launch_make_boundaries(h->p, h->g, array_of(h, which), faces_for(h), nullptr, h->stream));
    return E2D_OK;
  }

  int
  e2d_godunov_unsplit(e2d_handle * h, int nStep, double dt)
  {
    if (!h)
      return fail(E2D_ERR_INVALID, "bad argument");
    E2D_REFUSE_PENDING(h);
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    h->loop_primed = false;
    if (nStep % 2 == 0) // HydroRun.h:263-270
      return godunov_impl(h, h->U, h->U2, dt, true);
    return godunov_impl(h, h->U2, h->U, dt, true);
  }

  int
  e2d_godunov_unsplit_nobc(e2d_handle * h, int nStep, double dt)
  {
    if (!h)
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    h->loop_primed = false;
    if (nStep % 2 == 0)
      return godunov_impl(h, h->U, h->U2, dt, false);
    return godunov_impl(h, h->U2, h->U, dt, false);
  }

  // dt history: device buffer of at most kHistMax entries (steps beyond it are not recorded: a run of 10^8 steps must
  // not allocate 800 MB of history), grown geometrically; never reallocated while a peer may be spinning
  static const long kHistMax = 1l << 20;
  static int
  ensure_history(e2d_handle * h, long want)
  {
    if (want > kHistMax)
      want = kHistMax;
    if (want < 1024)
      want = 1024;
    if (h->hist_cap >= want)
      return E2D_OK;
    long cap = h->hist_cap > 0 ? h->hist_cap : 1024;
    while (cap < want)
      cap *= 2;
    double * nh = nullptr;
    E2D_CUDA(cudaMalloc(&nh, sizeof(double) * cap));
    E2D_CUDA(cudaMemsetAsync(nh, 0, sizeof(double) * cap, h->stream));
    if (h->d_hist)
      E2D_CUDA(cudaMemcpyAsync(nh, h->d_hist, sizeof(double) * h->hist_cap, cudaMemcpyDeviceToDevice, h->stream));
    E2D_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_hist);
    h->d_hist = nh;
    h->hist_cap = cap;
    return E2D_OK;
  }

  int
  e2d_run(e2d_handle * h, long max_steps, e2d_run_stats * stats)
  {
    if (!h)
      return fail(E2D_ERR_INVALID, "bad argument");
    const int nranks = h->slab.nranks, rank = h->slab.rank;
    if (nranks > 1 && !h->peers.connected)
      return fail(E2D_ERR_UNSUPPORTED, "e2d_run on a slab needs its peers: call e2d_ipc_connect / e2d_peer_connect_local");
    E2D_REFUSE_PENDING(h);
    const e2d_params & p = h->p;
    cudaStream_t       st = h->stream;
    E2D_CUDA(cudaSetDevice(h->device));
    if (max_steps < 0)
      max_steps = p.nStepmax;
    if (max_steps > 2147483647l)
      max_steps = 2147483647l; // nStep is an int, like the reference's (main.cpp:61)
    h->cfl_valid[0] = h->cfl_valid[1] = false; // the loop rewrites both arrays
    const unsigned long long launches0 = g_launches.load();
    if (h->seq_poisoned)
      return fail(E2D_ERR_CUDA, "slab loop: an earlier e2d_run on this handle lost a peer; the state is invalid");

    // dt history buffer (slabs: sized at connect time, see prepare_slab_loop)
    if (nranks == 1)
      if (int rc = ensure_history(h, max_steps + 1))
        return rc;

    // prime the loop state: (t, nStep) from the handle, invDt partial of the current array (main.cpp:128)
    // E2D_FORCE_LINKED=1 (development aid): run the peer-publishing instantiation of the step kernel on a single
    // GPU — no neighbours, it only publishes its invDt partial to its own slot — so that it can be timed and profiled
    // without a second rank.  E2D_TWO_LAUNCH=1: the round-1 single-GPU loop (boundary kernel + step kernel).
    static const bool force_linked = std::getenv("E2D_FORCE_LINKED") != nullptr;
    static const bool two_launch = std::getenv("E2D_TWO_LAUNCH") != nullptr;
    const bool        linked = nranks > 1 || force_linked;
    const bool        solo = !linked && !two_launch;
    SlabState &       hs = *h->h_state;
    std::memset(&hs, 0, sizeof hs);
    hs.t = h->t;
    hs.dt = h->dt_last;
    hs.nStep = h->nStep;
    hs.done = !(h->t < p.tEnd && h->nStep < max_steps);
    hs.solo_T[h->nStep & 3] = h->t; // solo: the first step of the call reads ring slot nStep & 3
    E2D_CUDA(cudaMemcpyAsync(h->d_state, h->h_state, sizeof(SlabState), cudaMemcpyHostToDevice, st));
    {
      const double * cur = (h->nStep % 2 == 0) ? h->U : h->U2;
      E2D_CUDA(launch_reduce_invdt(p, h->g, cur, solo ? &h->d_state->solo_acc[h->nStep & 3] : &h->d_state->invdt_acc, st));
    }

    SlabStepArgs sa;
    sa.mine = h->d_comm;
    sa.st = h->d_state;
    sa.has_lower = h->peers.lower >= 0;
    sa.has_upper = h->peers.upper >= 0;
    sa.nranks = nranks;
    sa.cfl = p.cfl;
    sa.tEnd = p.tEnd;
    sa.max_steps = (int)max_steps;
    sa.dt_hist = h->d_hist;
    sa.hist_cap = h->hist_cap;
    {
      // how long a kernel may wait for a peer's halo rows / invDt partial before it gives up (seconds of SM clock at
      // 2 GHz).  Ranks may enter e2d_run at different times (PeerSlabRun.run puts a barrier in front, other callers
      // should too); 60 s covers first-call initialisation and rank-0-only work in between.
      static const double timeout_s = [] {
        const char * e = std::getenv("E2D_PEER_TIMEOUT_S");
        const double v = e ? std::atof(e) : 60.0;
        return v > 0.0 ? v : 60.0;
      }();
      sa.timeout_clocks = (long long)(timeout_s * 2.0e9);
    }
    SlabPushArgs pa;
    pa.isize = h->g.isize;
    pa.jsize = h->g.jsize;
    pa.lower_jsize = h->peers.lower_jsize;
    pa.upper_jsize = h->peers.upper_jsize;
    for (int k = 0; k < kMaxRanks; ++k)
      pa.comm[k] = h->peers.comm[k];
    pa.st = h->d_state;
    pa.nranks = nranks;
    pa.rank = rank;
    pa.lower = h->peers.lower;
    pa.upper = h->peers.upper;
    const int faces = faces_for(h);
    FusedLink lk{};
    if (linked)
    {
      lk.cnt = h->d_comm->fused_cnt;
      lk.flag_lo = h->peers.lower >= 0 ? &h->peers.comm[h->peers.lower]->halo_flag[1] : nullptr; // I am its upper
      lk.flag_hi = h->peers.upper >= 0 ? &h->peers.comm[h->peers.upper]->halo_flag[0] : nullptr; // I am its lower
      for (int k = 0; k < kMaxRanks; ++k)
        lk.comm[k] = h->peers.comm[k];
      if (nranks == 1)
        lk.comm[0] = h->d_comm;
      lk.nranks = nranks;
      lk.rank = rank;
    }
    SoloLoop so;
    so.st = h->d_state;
    so.cfl = p.cfl;
    so.tEnd = p.tEnd;
    so.max_steps = (int)max_steps;
    so.dt_hist = h->d_hist;
    so.hist_cap = h->hist_cap;
    so.bc_xmin = p.boundary_type_xmin;
    so.bc_xmax = p.boundary_type_xmax;
    so.bc_ymin = p.boundary_type_ymin;
    so.bc_ymax = p.boundary_type_ymax;

    ProfileRegion pr_loop("main_loop"); // src/main.cpp:93 — the whole loop is this one call here
    E2D_CUDA(cudaEventRecord(h->ev[0], st));
    int       n_host = h->nStep; // parity the host believes in; wrong only after `done`, when the step is a no-op
    const int batch = 64;
    double    step_kernel_ms = 0.0;
    if (h->timing && h->ev_step.empty())
    {
      h->ev_step.resize(2 * batch, nullptr);
      for (auto & e : h->ev_step)
        E2D_CUDA(cudaEventCreate(&e));
    }
    // Every rank must issue the same number of steps (the flags count them): `done` is computed from identical
    // data on every rank and looked at after the same batches, so all ranks stop together.
    // The flags of this call start one past anything an earlier call's last fused step may have published.
    h->seq += 1;
    // Single GPU, no per-step event records in between: consecutive kernels are launched as programmatic
    // dependents of each other, so their launch latency overlaps the predecessor's tail (it matters on small grids,
    // where a step is a few microseconds).  Peers: kept off — their kernels spin on remote flags.
    static const bool pdl_off = std::getenv("E2D_NO_PDL") != nullptr;
    const bool        pdl = nranks == 1 && !h->timing && !pdl_off;
    bool first = true;
    bool finished = hs.done != 0;
    if (solo && !finished)
    { // the only boundary fill of the call: every later step finds the ghost cells pushed by its predecessor
      double * cur = (h->nStep % 2 == 0) ? h->U : h->U2;
      E2D_CUDA(launch_make_boundaries(p, h->g, cur, faces, nullptr, st));
    }
    while (!finished)
    {
      long todo = max_steps - n_host;
      if (todo > batch)
        todo = batch;
      for (long k = 0; k < todo; ++k, ++n_host)
      {
        const int which = n_host % 2; // 0: U -> U2
        double *  in = which == 0 ? h->U : h->U2;
        double *  out = which == 0 ? h->U2 : h->U;
        if (solo)
        {
          so.step = n_host;
          if (h->timing)
            E2D_CUDA(cudaEventRecord(h->ev_step[2 * k], st));
          E2D_CUDA(launch_fused_step(p, h->g, in, out, 0.0, nullptr, &h->d_state->solo_acc[(n_host + 1) & 3], nullptr, st,
                                     nullptr, nullptr, 2, 0, pdl, &so));
          if (h->timing)
            E2D_CUDA(cudaEventRecord(h->ev_step[2 * k + 1], st));
          continue;
        }
        h->seq += 1;
        sa.seq = pa.seq = h->seq;
        sa.parity = pa.parity = (int)(h->seq & 1);
        if (nranks > 1 && first)
        { // the call's first step: halo rows of the current array + the primed invDt partial, by a push kernel;
          // every later step finds them already published by the previous fused step
          pa.A = in;
          pa.lowerA = h->peers.lower >= 0 ? h->peers.lowerU[which] : nullptr;
          pa.upperA = h->peers.upper >= 0 ? h->peers.upperU[which] : nullptr;
          E2D_CUDA(launch_slab_push(pa, st));
        }
        first = false;
        E2D_CUDA(launch_slab_boundaries(p, h->g, in, faces, sa, st, pdl));
        if (h->timing)
          E2D_CUDA(cudaEventRecord(h->ev_step[2 * k], st));
        if (linked)
        {
          MarchPeers mp; // the step writes `out`: the array the NEXT step reads, on the neighbours too
          mp.lo = h->peers.lower >= 0 ? h->peers.lowerU[1 - which] : nullptr;
          mp.hi = h->peers.upper >= 0 ? h->peers.upperU[1 - which] : nullptr;
          mp.lo_jsize = h->peers.lower_jsize;
          mp.hi_jsize = h->peers.upper_jsize;
          lk.seq_next = h->seq + 1;
          lk.parity_next = (int)((h->seq + 1) & 1);
          E2D_CUDA(launch_fused_step(p, h->g, in, out, 0.0, &h->d_state->dt, &h->d_state->invdt_acc,
                                     &h->d_state->done, st, &mp, &lk));
        }
        else
          E2D_CUDA(launch_fused_step(p, h->g, in, out, 0.0, &h->d_state->dt, &h->d_state->invdt_acc,
                                     &h->d_state->done, st, nullptr, nullptr, 2, 0, pdl));
        if (h->timing)
          E2D_CUDA(cudaEventRecord(h->ev_step[2 * k + 1], st));
      }
      if (!solo)
        E2D_CUDA(launch_slab_finish(sa, st)); // closes the last opened step (a no-op for the next batch's first step)
      E2D_CUDA(cudaMemcpyAsync(h->h_state, h->d_state, sizeof(SlabState), cudaMemcpyDeviceToHost, st));
      E2D_CUDA(cudaStreamSynchronize(st));
      if (h->timing)
        for (long k = 0; k < todo; ++k)
        {
          float ms_k = 0;
          E2D_CUDA(cudaEventElapsedTime(&ms_k, h->ev_step[2 * k], h->ev_step[2 * k + 1]));
          step_kernel_ms += ms_k;
        }
      if (hs.error)
      {
        h->seq_poisoned = true;
        return fail(E2D_ERR_CUDA,
                    "slab loop: timed out waiting for a peer GPU (did a rank die, or enter e2d_run much later than the "
                    "others?); the loop was stopped and the state of this handle is INVALID");
      }
      finished = hs.done != 0 || n_host >= max_steps;
    }
    E2D_CUDA(cudaEventRecord(h->ev[1], st));
    E2D_CUDA(cudaEventSynchronize(h->ev[1]));
    float ms = 0;
    E2D_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->t = hs.t;
    h->nStep = hs.nStep;
    h->dt_last = hs.dt;
    if (stats)
    {
      stats->nStep = h->nStep;
      stats->t = h->t;
      stats->dt_last = h->dt_last;
      stats->seconds = ms * 1e-3;
      stats->launches = (long long)(g_launches.load() - launches0);
      stats->seconds_step_kernel = step_kernel_ms * 1e-3;
    }
    return E2D_OK;
  }

  // ------------------------------------------------------------------ peers of a slab (NVLink peer memory)
  // everything e2d_run would otherwise do lazily (kernel loading, the dt-history allocation) and that could wait for
  // a peer's spinning kernel when several ranks share a device
  static int
  prepare_slab_loop(e2d_handle * h)
  {
    E2D_CUDA(preload_slab_kernels());
    E2D_CUDA(preload_step_kernels());
    if (int rc = ensure_history(h, (long)h->p.nStepmax + 128))
      return rc;
    return E2D_OK;
  }

  static void
  neighbours_of(const e2d_handle * h, int & lower, int & upper)
  {
    const int r = h->slab.rank, n = h->slab.nranks;
    // e2d_create refuses slabs with only ONE of the two y faces periodic, so both wraps exist or neither does
    lower = r > 0 ? r - 1 : (h->p.boundary_type_ymin == E2D_BC_PERIODIC ? n - 1 : -1);
    upper = r < n - 1 ? r + 1 : (h->p.boundary_type_ymax == E2D_BC_PERIODIC ? 0 : -1);
    if (n == 1)
      lower = upper = -1;
  }

  int
  e2d_ipc_export(e2d_handle * h, e2d_ipc_blob * out)
  {
    if (!h || !out)
      return fail(E2D_ERR_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) <= E2D_IPC_HANDLE_BYTES, "handle size");
    std::memset(out, 0, sizeof *out);
    E2D_CUDA(cudaSetDevice(h->device));
    cudaIpcMemHandle_t m;
    E2D_CUDA(cudaIpcGetMemHandle(&m, h->U));
    std::memcpy(out->U, &m, sizeof m);
    E2D_CUDA(cudaIpcGetMemHandle(&m, h->U2));
    std::memcpy(out->U2, &m, sizeof m);
    E2D_CUDA(cudaIpcGetMemHandle(&m, h->d_comm));
    std::memcpy(out->comm, &m, sizeof m);
    out->rank = h->slab.rank;
    out->nranks = h->slab.nranks;
    out->ny_loc = h->slab.ny_loc;
    out->device = h->device;
    return E2D_OK;
  }

  int
  e2d_ipc_connect(e2d_handle * h, const e2d_ipc_blob * blobs, int nranks)
  {
    if (!h || !blobs || nranks != h->slab.nranks || nranks > kMaxRanks)
      return fail(E2D_ERR_INVALID, "bad argument (nranks must match the slab and be <= 16)");
    E2D_CUDA(cudaSetDevice(h->device));
    int lower, upper;
    neighbours_of(h, lower, upper);
    auto open = [&](const unsigned char * raw, void ** ptr) -> cudaError_t {
      cudaIpcMemHandle_t m;
      std::memcpy(&m, raw, sizeof m);
      cudaError_t e = cudaIpcOpenMemHandle(ptr, m, cudaIpcMemLazyEnablePeerAccess);
      if (e == cudaSuccess)
        h->peers.ipc_opened.push_back(*ptr);
      return e;
    };
    const int me = h->slab.rank;
    for (int k = 0; k < nranks; ++k)
    {
      if (blobs[k].rank != k || blobs[k].nranks != nranks)
        return fail(E2D_ERR_INVALID, "blobs must be in rank order");
      if (k == me)
      {
        h->peers.comm[k] = h->d_comm;
        continue;
      }
      void * c = nullptr;
      E2D_CUDA(open(blobs[k].comm, &c));
      h->peers.comm[k] = static_cast<SlabComm *>(c);
    }
    void *lU = nullptr, *lU2 = nullptr;
    if (lower >= 0)
    {
      E2D_CUDA(open(blobs[lower].U, &lU));
      E2D_CUDA(open(blobs[lower].U2, &lU2));
      h->peers.lowerU[0] = static_cast<double *>(lU);
      h->peers.lowerU[1] = static_cast<double *>(lU2);
      h->peers.lower_jsize = blobs[lower].ny_loc + 2 * h->p.ghostWidth;
    }
    if (upper >= 0)
    {
      if (upper == lower)
      { // two ranks with a periodic wrap: one peer is both neighbours (an IPC handle can be opened only once)
        h->peers.upperU[0] = h->peers.lowerU[0];
        h->peers.upperU[1] = h->peers.lowerU[1];
      }
      else
      {
        void *uU = nullptr, *uU2 = nullptr;
        E2D_CUDA(open(blobs[upper].U, &uU));
        E2D_CUDA(open(blobs[upper].U2, &uU2));
        h->peers.upperU[0] = static_cast<double *>(uU);
        h->peers.upperU[1] = static_cast<double *>(uU2);
      }
      h->peers.upper_jsize = blobs[upper].ny_loc + 2 * h->p.ghostWidth;
    }
    h->peers.lower = lower;
    h->peers.upper = upper;
    h->peers.connected = true;
    return prepare_slab_loop(h);
  }

  int
  e2d_peer_connect_local(e2d_handle ** hs, int n)
  {
    if (!hs || n < 1 || n > kMaxRanks)
      return fail(E2D_ERR_INVALID, "bad argument");
    for (int r = 0; r < n; ++r)
      if (!hs[r] || hs[r]->slab.rank != r || hs[r]->slab.nranks != n)
        return fail(E2D_ERR_INVALID, "handles must be the n slabs of one run, in rank order");
    { // Sedov on slabs of one process: the integer "all-reduce" of the disc-cell counts is a loop
      unsigned long long total = 0;
      bool               pending = false;
      for (int r = 0; r < n; ++r)
      {
        total += hs[r]->blast_inside_local;
        pending = pending || hs[r]->blast_pending;
      }
      if (pending)
        for (int r = 0; r < n; ++r)
          if (int rc = e2d_blast_renormalise(hs[r], total))
            return rc;
    }
    for (int r = 0; r < n; ++r)
    {
      e2d_handle * h = hs[r];
      E2D_CUDA(cudaSetDevice(h->device));
      for (int k = 0; k < n; ++k)
      {
        h->peers.comm[k] = hs[k]->d_comm;
        if (k != r && hs[k]->device != h->device)
        {
          cudaError_t e = cudaDeviceEnablePeerAccess(hs[k]->device, 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled)
            cudaGetLastError();
          else if (e != cudaSuccess)
            return fail_cuda(e, "cudaDeviceEnablePeerAccess");
        }
      }
      int lower, upper;
      neighbours_of(h, lower, upper);
      if (lower >= 0)
      {
        h->peers.lowerU[0] = hs[lower]->U;
        h->peers.lowerU[1] = hs[lower]->U2;
        h->peers.lower_jsize = hs[lower]->g.jsize;
      }
      if (upper >= 0)
      {
        h->peers.upperU[0] = hs[upper]->U;
        h->peers.upperU[1] = hs[upper]->U2;
        h->peers.upper_jsize = hs[upper]->g.jsize;
      }
      h->peers.lower = lower;
      h->peers.upper = upper;
      h->peers.connected = true;
      if (int rc = prepare_slab_loop(h))
        return rc;
    }
    return E2D_OK;
  }

  int
  e2d_get_dt_history(e2d_handle * h, double * dts, long n_cap, long * n)
  {
    if (!h || !dts || !n)
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    long cnt = h->nStep < h->hist_cap ? h->nStep : h->hist_cap;
    if (cnt > n_cap)
      cnt = n_cap;
    if (cnt > 0 && h->d_hist)
    {
      E2D_CUDA(cudaMemcpyAsync(dts, h->d_hist, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream));
      E2D_CUDA(cudaStreamSynchronize(h->stream));
    }
    else
      cnt = 0;
    *n = cnt;
    return E2D_OK;
  }

  int
  e2d_set_time(e2d_handle * h, double t, int nStep)
  {
    if (!h || nStep < 0)
      return fail(E2D_ERR_INVALID, "bad argument");
    h->t = t;
    h->nStep = nStep;
    return E2D_OK;
  }

  int
  e2d_download(e2d_handle * h, int which, double * host, int layout)
  {
    double * A = h ? array_of(h, which) : nullptr;
    if (!A || !host)
      return fail(E2D_ERR_INVALID, "bad argument (array not allocated?)");
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    if (layout == E2D_LAYOUT_SOA)
    {
      E2D_CUDA(cudaMemcpyAsync(host, A, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      E2D_CUDA(cudaStreamSynchronize(h->stream));
      return E2D_OK;
    }
    if (layout != E2D_LAYOUT_KOKKOS_OMP)
      return fail(E2D_ERR_INVALID, "unknown layout");
    std::vector<double> tmp(h->n);
    E2D_CUDA(cudaMemcpyAsync(tmp.data(), A, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    E2D_CUDA(cudaStreamSynchronize(h->stream));
    soa_to_kokkos_omp(tmp.data(), host, h->g.isize, h->g.jsize);
    return E2D_OK;
  }

  int
  e2d_upload(e2d_handle * h, int which, const double * host, int layout)
  {
    double * A = h ? array_of(h, which) : nullptr;
    if (!A || !host)
      return fail(E2D_ERR_INVALID, "bad argument (array not allocated?)");
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    h->loop_primed = false;
    h->cfl_valid[0] = h->cfl_valid[1] = false;
    if (layout == E2D_LAYOUT_SOA)
    {
      E2D_CUDA(cudaMemcpyAsync(A, host, h->n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      E2D_CUDA(cudaStreamSynchronize(h->stream));
      return E2D_OK;
    }
    if (layout != E2D_LAYOUT_KOKKOS_OMP)
      return fail(E2D_ERR_INVALID, "unknown layout");
    std::vector<double> tmp(h->n);
    soa_from_kokkos_omp(host, tmp.data(), h->g.isize, h->g.jsize);
    E2D_CUDA(cudaMemcpyAsync(A, tmp.data(), h->n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    E2D_CUDA(cudaStreamSynchronize(h->stream));
    return E2D_OK;
  }

  double *
  e2d_device_ptr(e2d_handle * h, int which)
  {
    if (h)
      h->cfl_cache_ok = false; // the caller may write through the pointer: compute_dt re-reads the array from now on
    return h ? array_of(h, which) : nullptr;
  }

  void *
  e2d_stream(e2d_handle * h)
  {
    return h ? (void *)h->stream : nullptr;
  }

  int
  e2d_synchronize(e2d_handle * h)
  {
    if (!h)
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    E2D_CUDA(cudaStreamSynchronize(h->stream));
    return E2D_OK;
  }

  int
  e2d_get_params(e2d_handle * h, e2d_params * out)
  {
    if (!h || !out)
      return fail(E2D_ERR_INVALID, "bad argument");
    *out = h->p;
    return E2D_OK;
  }

  int
  e2d_step_host(e2d_handle * h, const double * U_host_in, double * U_host_out, double * dt_out)
  {
    if (!h || !U_host_in || !U_host_out)
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    const e2d_params & p = h->p;
    cudaStream_t       st = h->stream;
    const size_t       bytes = h->n * sizeof(double);
    E2D_CUDA(cudaMemcpyAsync(h->U, U_host_in, bytes, cudaMemcpyHostToDevice, st));
    E2D_CUDA(cudaMemsetAsync(h->d_bits, 0, sizeof(unsigned long long), st));
    E2D_CUDA(launch_reduce_invdt(p, h->g, h->U, h->d_bits, st));
    E2D_CUDA(launch_make_boundaries(p, h->g, h->U, faces_for(h), nullptr, st));
    // dt = cfl / invDt on the device (IEEE division, same value as HydroRun.h:246), no host round trip
    h->h_loop->t = 0.0;
    h->h_loop->dt = 0.0;
    h->h_loop->nStep = 0;
    h->h_loop->done = 0;
    h->h_loop->invdt_cur = 0;
    h->h_loop->invdt_next = 0;
    E2D_CUDA(cudaMemcpyAsync(h->d_loop, h->h_loop, sizeof(LoopState), cudaMemcpyHostToDevice, st));
    E2D_CUDA(cudaMemcpyAsync(&h->d_loop->invdt_cur, h->d_bits, sizeof(unsigned long long),
                             cudaMemcpyDeviceToDevice, st));
    E2D_CUDA(launch_loop_begin_step(h->d_loop, p.cfl, 1e300, st));
    E2D_CUDA(launch_fused_step(p, h->g, h->U, h->U2, 0.0, &h->d_loop->dt, nullptr, nullptr, st));
    // Ghost cells of the result: filled from the NEW interior (what the next make_boundaries would produce), so that
    // the array handed back is self-consistent.  This differs from the reference's godunov_unsplit, whose out array
    // keeps the INPUT's filled ghosts (deep_copy, HydroRun.h:302) — e2d_godunov_unsplit reproduces that; this
    // convenience entry point has no counterpart in the reference.
    E2D_CUDA(launch_make_boundaries(p, h->g, h->U2, faces_for(h), nullptr, st));
    E2D_CUDA(cudaMemcpyAsync(U_host_out, h->U2, bytes, cudaMemcpyDeviceToHost, st));
    E2D_CUDA(cudaMemcpyAsync(h->h_loop, h->d_loop, sizeof(LoopState), cudaMemcpyDeviceToHost, st));
    E2D_CUDA(cudaStreamSynchronize(st));
    if (dt_out)
      *dt_out = h->h_loop->dt;
    h->loop_primed = false;
    h->cfl_valid[0] = h->cfl_valid[1] = false;
    return E2D_OK;
  }

  // The step of a caller whose state lives in HOST memory, streamed: the rows travel host -> device in chunks, each
  // chunk is advanced as soon as the two rows above it have landed, and finished chunks travel back while later
  // ones are still arriving or being computed — H2D, the fused step and D2H overlap (full-duplex PCIe), so the call
  // costs about one direction's transfer instead of two.  This needs dt BEFORE the rows are all there, which is
  // exactly the reference's call structure: dt = compute_dt(); godunov_unsplit(nStep, dt) (src/main.cpp:128,139) —
  // and the CFL reduction of the NEW state rides on the step, so the caller gets the next dt with the result.
  int
  e2d_step_host_streamed(e2d_handle * h, const double * U_host_in, double * U_host_out, double dt_in, int chunk_rows,
                         double * dt_used, double * dt_next)
  {
    if (!h || !U_host_in || !U_host_out)
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device);
    const e2d_params & p = h->p;
    const Geom &       g = h->g;
    cudaStream_t       st = h->stream;
    const int          isize = g.isize, jsize = g.jsize, ny = g.ny;
    const size_t       plane = (size_t)isize * jsize;
    const int          faces = faces_for(h);
    if (!h->s_in)
    {
      E2D_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
      E2D_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    }
    // Chunk boundaries: jb[k] = first interior row of chunk k, jb[nchunk] = jsize - 2.  Default: ~32 chunks, with the
    // first two and the last one cut in four — the copy of the first chunks (nothing to overlap with yet) and the way
    // back of the last one (nothing left to overlap with) are the un-overlapped head and tail of the pipeline.
    const bool taper = chunk_rows <= 0;
    if (chunk_rows <= 0)
      chunk_rows = (ny + 31) / 32;
    if (chunk_rows < 16)
      chunk_rows = 16; // every chunk holds the source rows of the y faces next to it
    std::vector<int> jb;
    {
      const int small = chunk_rows / 4 >= 16 ? chunk_rows / 4 : 16;
      const int n_uniform = (ny + chunk_rows - 1) / chunk_rows;
      if (taper && n_uniform >= 8)
      { // between the tapered head and tail: half-size chunks (measured at 8192^2: 128 rows 52.1 ms, 256 rows 53.1,
        // 512 rows 65 — the way back of a chunk can only start once the NEXT chunk has landed)
        chunk_rows = chunk_rows / 2 >= 16 ? chunk_rows / 2 : 16;
        static const char * env = std::getenv("E2D_STREAM_MAIN_ROWS"); // development knob
        if (env && std::atoi(env) >= 16)
          chunk_rows = std::atoi(env);
      }
      int       j = 2;
      jb.push_back(j);
      if (taper && n_uniform >= 8)
      {
        const int tail_from = jsize - 2 - 4 * small;
        for (int k = 0; k < 8 && j + small < tail_from; ++k) // the first two chunks' worth of rows in 8 pieces
          jb.push_back(j += small);
        while (j + chunk_rows < tail_from)
          jb.push_back(j += chunk_rows);
        if (j < tail_from)
        {
          if (tail_from - j < 16 && jb.size() > 1)
            jb.back() = j = tail_from; // a remainder too short to stand alone extends the chunk before it
          else
            jb.push_back(j = tail_from);
        }
        while (j + small < jsize - 2)
          jb.push_back(j += small);
      }
      else
      {
        while (j + chunk_rows < jsize - 2)
          jb.push_back(j += chunk_rows);
      }
      jb.push_back(jsize - 2);
    }
    const int nchunk = (int)jb.size() - 1;
    while ((int)h->ev_pool.size() < 2 * nchunk + 4)
    {
      cudaEvent_t e;
      E2D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev_pool.push_back(e);
    }
    cudaEvent_t * ev_in = h->ev_pool.data();           // [nchunk] chunk landed
    cudaEvent_t * ev_cmp = h->ev_pool.data() + nchunk; // [nchunk] chunk advanced
    cudaEvent_t   ev_start = h->ev_pool[2 * nchunk], ev_ymin = h->ev_pool[2 * nchunk + 1],
                ev_fin = h->ev_pool[2 * nchunk + 2];
    auto first_row = [&](int k) { return jb[k]; };
    auto last_row = [&](int k) { return jb[k + 1]; };
    // rows [jlo, jhi) of all four planes
    // the rows of the four variable planes go as ONE 2-D copy (4 "lines" one plane apart): a quarter of the copy
    // calls, 52.3 instead of 54.0 ms per 8192^2 step.  E2D_COPY_PER_PLANE=1 restores one copy per plane.
    static const bool copy2d = std::getenv("E2D_COPY_PER_PLANE") == nullptr;
    auto copy_rows = [&](double * dst, const double * src, int jlo, int jhi, cudaMemcpyKind kind, cudaStream_t s) {
      if (copy2d)
      {
        const size_t o = (size_t)jlo * isize;
        return cudaMemcpy2DAsync(dst + o, plane * sizeof(double), src + o, plane * sizeof(double),
                                 (size_t)(jhi - jlo) * isize * sizeof(double), 4, kind, s);
      }
      for (int v = 0; v < 4; ++v)
      {
        const size_t    o = (size_t)jlo * isize + v * plane;
        const cudaError_t e = cudaMemcpyAsync(dst + o, src + o, (size_t)(jhi - jlo) * isize * sizeof(double), kind, s);
        if (e != cudaSuccess)
          return e;
      }
      return cudaSuccess;
    };

    E2D_CUDA(cudaMemsetAsync(h->d_bits, 0, sizeof(unsigned long long), st));
    E2D_CUDA(cudaEventRecord(ev_start, st));
    E2D_CUDA(cudaStreamWaitEvent(h->s_in, ev_start, 0));
    E2D_CUDA(cudaStreamWaitEvent(h->s_out, ev_start, 0));

    // ---- host -> device, in row order; a periodic YMIN face reads the LAST two interior rows: send those first
    const bool early_top = (faces & E2D_FACES_YMIN) && p.boundary_type_ymin == E2D_BC_PERIODIC && nchunk > 1;
    if (early_top)
      E2D_CUDA(copy_rows(h->U, U_host_in, jsize - 4, jsize - 2, cudaMemcpyHostToDevice, h->s_in));
    // (with early_top the last chunk rewrites those two rows, so it is enqueued only after the YMIN fill that reads
    //  them has been — see below)
    for (int k = 0; k < nchunk - (early_top ? 1 : 0); ++k)
    {
      const int jlo = k == 0 ? 0 : first_row(k), jhi = k == nchunk - 1 ? jsize : last_row(k);
      E2D_CUDA(copy_rows(h->U, U_host_in, jlo, jhi, cudaMemcpyHostToDevice, h->s_in));
      E2D_CUDA(cudaEventRecord(ev_in[k], h->s_in));
    }

    double dt = dt_in;
    if (!(dt > 0.0))
    {
      // no dt from the caller: compute_dt needs every row, so only the step and the way back overlap
      if (early_top)
      {
        const int k = nchunk - 1;
        E2D_CUDA(copy_rows(h->U, U_host_in, first_row(k), jsize, cudaMemcpyHostToDevice, h->s_in));
        E2D_CUDA(cudaEventRecord(ev_in[k], h->s_in));
      }
      E2D_CUDA(cudaStreamWaitEvent(st, ev_in[nchunk - 1], 0));
      E2D_CUDA(launch_reduce_invdt(p, g, h->U, h->d_bits, st));
      E2D_CUDA(cudaMemcpyAsync(&h->h_loop->invdt_cur, h->d_bits, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      E2D_CUDA(cudaMemsetAsync(h->d_bits, 0, sizeof(unsigned long long), st));
      E2D_CUDA(cudaStreamSynchronize(st));
      double inv;
      std::memcpy(&inv, &h->h_loop->invdt_cur, sizeof inv);
      dt = p.cfl / inv; // HydroRun.h:246
    }
    const bool top_pending = early_top && dt_in > 0.0; // the last chunk is still to be enqueued

    // ---- fill, advance, device -> host
    auto advance = [&](int m) -> int {
      const int ja = first_row(m), jb = last_row(m);
      E2D_CUDA(launch_fused_step(p, g, h->U, h->U2, dt, nullptr, h->d_bits, nullptr, st, nullptr, nullptr, ja, jb));
      E2D_CUDA(launch_bc_x_rows(p, g, h->U2, faces, ja, jb, st)); // ghost columns of the result
      E2D_CUDA(cudaEventRecord(ev_cmp[m], st));
      E2D_CUDA(cudaStreamWaitEvent(h->s_out, ev_cmp[m], 0));
      E2D_CUDA(copy_rows(U_host_out, h->U2, ja, jb, cudaMemcpyDeviceToHost, h->s_out));
      return E2D_OK;
    };
    for (int k = 0; k < nchunk; ++k)
    {
      E2D_CUDA(cudaStreamWaitEvent(st, ev_in[k], 0));
      const int jlo = k == 0 ? 0 : first_row(k), jhi = k == nchunk - 1 ? jsize : last_row(k);
      E2D_CUDA(launch_bc_x_rows(p, g, h->U, faces, jlo, jhi, st));
      if (k == 0)
      {
        if (early_top)
          E2D_CUDA(launch_bc_x_rows(p, g, h->U, faces, jsize - 4, jsize - 2, st));
        if (faces & E2D_FACES_YMIN)
          E2D_CUDA(launch_make_boundaries(p, g, h->U, E2D_FACES_YMIN, nullptr, st));
        E2D_CUDA(cudaEventRecord(ev_ymin, st));
        if (top_pending)
        {
          // now the last chunk may overwrite the two rows the YMIN fill has read
          const int kl = nchunk - 1;
          E2D_CUDA(cudaStreamWaitEvent(h->s_in, ev_ymin, 0));
          E2D_CUDA(copy_rows(h->U, U_host_in, first_row(kl), jsize, cudaMemcpyHostToDevice, h->s_in));
          E2D_CUDA(cudaEventRecord(ev_in[kl], h->s_in));
        }
      }
      if (k == nchunk - 1 && (faces & E2D_FACES_YMAX))
        E2D_CUDA(launch_make_boundaries(p, g, h->U, E2D_FACES_YMAX, nullptr, st));
      if (k >= 1)
        if (int rc = advance(k - 1))
          return rc;
      if (k == nchunk - 1)
        if (int rc = advance(k))
          return rc;
    }
    // y-ghost rows of the result (faces this slab owns), the CFL reduction of the new state
    if (faces & (E2D_FACES_YMIN | E2D_FACES_YMAX))
      E2D_CUDA(launch_make_boundaries(p, g, h->U2, faces & (E2D_FACES_YMIN | E2D_FACES_YMAX), nullptr, st));
    E2D_CUDA(cudaEventRecord(ev_fin, st));
    E2D_CUDA(cudaStreamWaitEvent(h->s_out, ev_fin, 0));
    if (faces & E2D_FACES_YMIN)
      E2D_CUDA(copy_rows(U_host_out, h->U2, 0, 2, cudaMemcpyDeviceToHost, h->s_out));
    if (faces & E2D_FACES_YMAX)
      E2D_CUDA(copy_rows(U_host_out, h->U2, jsize - 2, jsize, cudaMemcpyDeviceToHost, h->s_out));
    E2D_CUDA(cudaMemcpyAsync(&h->h_loop->invdt_next, h->d_bits, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             h->s_out));
    E2D_CUDA(cudaStreamSynchronize(h->s_out));
    E2D_CUDA(cudaStreamSynchronize(st));
    double inv_next;
    std::memcpy(&inv_next, &h->h_loop->invdt_next, sizeof inv_next);
    if (dt_used)
      *dt_used = dt;
    if (dt_next)
      *dt_next = p.cfl / inv_next; // = compute_dt of the state just written (this slab's rows)
    h->loop_primed = false;
    h->cfl_valid[0] = h->cfl_valid[1] = false;
    return E2D_OK;
  }

  // A time march whose state lives in HOST memory, the steps PIPELINED.  e2d_step_host_streamed overlaps the two copy
  // directions and the kernel inside one step but returns between steps, so every step pays the head of its pipeline
  // (nothing can come back before the first rows have arrived and been advanced) and its tail.  Here the host -> device
  // stream never stops: the rows of step s+1 follow the rows of step s as soon as (a) the way back of the same rows in
  // step s has landed in the host buffer they are read from and (b) the kernels of step s no longer read the device
  // rows they overwrite.  dt never visits the host: the CFL maximum rides on the chunk kernels, a one-thread kernel forms
  // dt = cfl / invDt between the steps (HydroRun.h:246).  tEnd is not looked at: the caller chooses nsteps.
  int
  e2d_march_host(e2d_handle * h, double * buf_a, double * buf_b, long nsteps, int chunk_rows, double * dts, double * t_io)
  {
    if (!h || !buf_a || !buf_b || nsteps < 0)
      return fail(E2D_ERR_INVALID, "bad argument");
    if (!h->whole)
      return fail(E2D_ERR_UNSUPPORTED, "e2d_march_host needs the whole domain on one device (slabs: e2d_step_host_streamed "
                                       "per step, ghost rows exchanged by the caller)");
    E2D_REFUSE_PENDING(h);
    const e2d_params & p = h->p;
    if (p.boundary_type_ymin == E2D_BC_PERIODIC || p.boundary_type_ymax == E2D_BC_PERIODIC)
      return fail(E2D_ERR_UNSUPPORTED, "e2d_march_host: a periodic y direction makes the first rows of a step depend on the "
                                       "last rows of the step before; use e2d_step_host_streamed per step");
    cudaSetDevice(h->device);
    const Geom & g = h->g;
    cudaStream_t st = h->stream;
    const int    isize = g.isize, jsize = g.jsize, ny = g.ny;
    const size_t plane = (size_t)isize * jsize;
    const int    faces = faces_for(h);
    if (nsteps == 0)
      return E2D_OK;
    if (!h->s_in)
    {
      E2D_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
      E2D_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    }
    // uniform chunks of interior rows (>= 16 rows: every chunk holds the source rows of the faces next to it)
    if (chunk_rows <= 0)
      chunk_rows = (ny + 63) / 64;
    if (chunk_rows < 16)
      chunk_rows = 16;
    std::vector<int> jb;
    for (int j = 2; j < jsize - 2; j += chunk_rows)
      jb.push_back(j);
    if (jb.size() > 1 && jsize - 2 - jb.back() < 16)
      jb.pop_back(); // a remainder too short to stand alone joins the chunk before it
    jb.push_back(jsize - 2);
    const int nchunk = (int)jb.size() - 1;
    // events: [parity][kind][chunk], kind 0 = rows landed on the device, 1 = chunk advanced, 2 = rows landed on the host
    while ((int)h->ev_pool.size() < 6 * nchunk + 4)
    {
      cudaEvent_t e;
      E2D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev_pool.push_back(e);
    }
    auto ev = [&](int parity, int kind, int k) { return h->ev_pool[(size_t)((parity * 3 + kind) * nchunk + k)]; };
    cudaEvent_t ev_start = h->ev_pool[6 * nchunk];
    auto copy_rows = [&](double * dst, const double * src, int jlo, int jhi, cudaMemcpyKind kind, cudaStream_t s) {
      const size_t o = (size_t)jlo * isize;
      return cudaMemcpy2DAsync(dst + o, plane * sizeof(double), src + o, plane * sizeof(double),
                               (size_t)(jhi - jlo) * isize * sizeof(double), 4, kind, s);
    };
    // dt of every step, on the device until the end
    double * d_dts = nullptr;
    E2D_CUDA(cudaMalloc(&d_dts, sizeof(double) * (size_t)nsteps));
    int rc = E2D_OK;
#define E2D_MARCH(call)                       \
  {                                           \
    cudaError_t e_ = (call);                  \
    if (e_ != cudaSuccess && rc == E2D_OK)    \
      rc = fail_cuda(e_, #call);              \
  }
    h->h_loop->t = t_io ? *t_io : 0.0;
    h->h_loop->dt = 0.0;
    h->h_loop->nStep = 0;
    h->h_loop->done = 0;
    h->h_loop->invdt_cur = 0;
    h->h_loop->invdt_next = 0;
    E2D_MARCH(cudaMemcpyAsync(h->d_loop, h->h_loop, sizeof(LoopState), cudaMemcpyHostToDevice, st));
    E2D_MARCH(cudaEventRecord(ev_start, st));
    E2D_MARCH(cudaStreamWaitEvent(h->s_in, ev_start, 0));
    E2D_MARCH(cudaStreamWaitEvent(h->s_out, ev_start, 0));

    for (long s = 0; s < nsteps && rc == E2D_OK; ++s)
    {
      const int      par = (int)(s & 1), prev = 1 - par;
      const double * in = par == 0 ? buf_a : buf_b;
      double *       out = par == 0 ? buf_b : buf_a;
      // ---- host -> device (interior rows only: the physical ghost cells are refilled on the device)
      for (int k = 0; k < nchunk; ++k)
      {
        if (s > 0)
        {
          E2D_MARCH(cudaStreamWaitEvent(h->s_in, ev(prev, 2, k), 0));                                   // (a)
          E2D_MARCH(cudaStreamWaitEvent(h->s_in, ev(prev, 1, k + 1 < nchunk ? k + 1 : nchunk - 1), 0)); // (b)
        }
        E2D_MARCH(copy_rows(h->U, in, jb[k], jb[k + 1], cudaMemcpyHostToDevice, h->s_in));
        E2D_MARCH(cudaEventRecord(ev(par, 0, k), h->s_in));
      }
      if (s == 0)
      { // the first dt needs the whole state: main.cpp:87,128
        E2D_MARCH(cudaStreamWaitEvent(st, ev(par, 0, nchunk - 1), 0));
        E2D_MARCH(launch_reduce_invdt(p, g, h->U, &h->d_loop->invdt_cur, st));
      }
      E2D_MARCH(launch_loop_begin_step(h->d_loop, p.cfl, 1e300, st));
      // ---- fill, advance, device -> host
      auto advance = [&](int m) {
        const int ja = jb[m], jz = jb[m + 1];
        if (s > 0)
          E2D_MARCH(cudaStreamWaitEvent(st, ev(prev, 2, m), 0)); // the previous result's rows have left U2
        E2D_MARCH(launch_fused_step(p, g, h->U, h->U2, 0.0, &h->d_loop->dt, &h->d_loop->invdt_next, nullptr, st, nullptr,
                                    nullptr, ja, jz));
        E2D_MARCH(launch_bc_x_rows(p, g, h->U2, faces, ja, jz, st)); // ghost columns of the result
        E2D_MARCH(cudaEventRecord(ev(par, 1, m), st));
        E2D_MARCH(cudaStreamWaitEvent(h->s_out, ev(par, 1, m), 0));
        E2D_MARCH(copy_rows(out, h->U2, ja, jz, cudaMemcpyDeviceToHost, h->s_out));
        E2D_MARCH(cudaEventRecord(ev(par, 2, m), h->s_out));
      };
      for (int k = 0; k < nchunk; ++k)
      {
        E2D_MARCH(cudaStreamWaitEvent(st, ev(par, 0, k), 0));
        E2D_MARCH(launch_bc_x_rows(p, g, h->U, faces, jb[k], jb[k + 1], st));
        if (k == 0 && (faces & E2D_FACES_YMIN))
          E2D_MARCH(launch_make_boundaries(p, g, h->U, E2D_FACES_YMIN, nullptr, st));
        if (k == nchunk - 1 && (faces & E2D_FACES_YMAX))
          E2D_MARCH(launch_make_boundaries(p, g, h->U, E2D_FACES_YMAX, nullptr, st));
        if (k >= 1)
          advance(k - 1);
        if (k == nchunk - 1)
          advance(k);
      }
      E2D_MARCH(launch_loop_end_step(h->d_loop, 1e300, 2147483647, d_dts, nsteps, st));
    }
    // ghost rows of the final state (the interior rows' ghost columns travelled with them)
    double * last = (nsteps & 1) ? buf_b : buf_a;
    if (rc == E2D_OK)
    {
      cudaEvent_t ev_fin = h->ev_pool[6 * nchunk + 1];
      E2D_MARCH(launch_make_boundaries(p, g, h->U2, faces & (E2D_FACES_YMIN | E2D_FACES_YMAX), nullptr, st));
      E2D_MARCH(cudaEventRecord(ev_fin, st));
      E2D_MARCH(cudaStreamWaitEvent(h->s_out, ev_fin, 0));
      E2D_MARCH(copy_rows(last, h->U2, 0, 2, cudaMemcpyDeviceToHost, h->s_out));
      E2D_MARCH(copy_rows(last, h->U2, jsize - 2, jsize, cudaMemcpyDeviceToHost, h->s_out));
      E2D_MARCH(cudaMemcpyAsync(h->h_loop, h->d_loop, sizeof(LoopState), cudaMemcpyDeviceToHost, st));
    }
    cudaStreamSynchronize(h->s_in);
    cudaStreamSynchronize(h->s_out);
    cudaStreamSynchronize(st);
    if (rc == E2D_OK && dts)
      E2D_MARCH(cudaMemcpy(dts, d_dts, sizeof(double) * (size_t)nsteps, cudaMemcpyDeviceToHost));
    cudaFree(d_dts);
#undef E2D_MARCH
    if (rc == E2D_OK && t_io)
      *t_io = h->h_loop->t;
    h->loop_primed = false;
    h->cfl_valid[0] = h->cfl_valid[1] = false;
    return rc;
  }

  int
  e2d_save_vtk(e2d_handle * h, int which, int iStep)
  {
    double * A = h ? array_of(h, which) : nullptr;
    if (!A)
      return fail(E2D_ERR_INVALID, "bad argument");
    if (!h->whole)
      return fail(E2D_ERR_UNSUPPORTED, "saveData on a y-slab handle: every rank would write its piece under the same file "
                                       "name; gather the interior (e2d_download / PeerSlabRun.gather_interior) and write it "
                                       "from one rank");
    if (h->p.vtkAppended)
      return e2d_save_vtk_appended(h, which, iStep);
    cudaSetDevice(h->device); // the handle may be driven from a thread whose current device differs
    const e2d_params &  p = h->p;
    std::vector<double> host(h->n);
    E2D_CUDA(cudaMemcpyAsync(host.data(), A, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    E2D_CUDA(cudaStreamSynchronize(h->stream));

    // HydroRun.h:537-544
    std::ostringstream stepNum;
    stepNum.width(7);
    stepNum.fill('0');
    stepNum << iStep;
    const std::string filename = std::string(p.outputDir) + "/" + p.outputPrefix + "_" + stepNum.str() + ".vti";
    std::fstream      f;
    f.open(filename.c_str(), std::ios_base::out);
    if (!f.is_open())
      return fail(E2D_ERR_IO, "cannot open " + filename);

    const int isize = h->g.isize, jsize = h->g.jsize, nx = p.nx, ny = h->g.ny;
    // HydroRun.h:551-570 (little-endian host)
    f << "<?xml version=\"1.0\"?>\n";
    f << "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
    f << "  <ImageData WholeExtent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << 0 << "\" "
      << "Origin=\"" << p.xmin << " " << p.ymin << " " << 0.0 << "\" "
      << "Spacing=\"" << p.dx << " " << p.dy << " " << 0.0 << "\">\n";
    f << "  <Piece Extent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << 0 << " "
      << "\">\n";
    f << "    <PointData>\n";
    f << "    </PointData>\n";
    f << "    <CellData>\n";
    static const char * varNames[4] = { "rho", "E", "mx", "my" }; // HydroParams.cpp:17
    for (int v = 0; v < 4; ++v)
    { // HydroRun.h:573-598: interior cells, i fastest
      f << "    <DataArray type=\"Float64\" Name=\"" << varNames[v] << "\" format=\"ascii\" >\n";
      for (int j = 2; j < jsize - 2; ++j)
        for (int i = 2; i < isize - 2; ++i)
          f << host[(size_t)i + (size_t)isize * ((size_t)j + (size_t)jsize * v)] << " ";
      f << "\n    </DataArray>\n";
    }
    f << "    </CellData>\n";
    f << "  </Piece>\n";
    f << "  </ImageData>\n";
    f << "</VTKFile>\n";
    f.close();
    return f.fail() ? fail(E2D_ERR_IO, "write failed: " + filename) : (int)E2D_OK;
  }

  // ---- fast output: interior gathered on the device, D2H into pinned double buffers overlapped with pwrite ----
  static int
  stream_interior_to_file(e2d_handle * h, const double * A, int fd, off_t base, const std::string & fname)
  {
    // file layout from `base`: 4 blocks, block v = [prefix bytes] nx*ny doubles of variable v (rows in order)
    const Geom &   g = h->g;
    const size_t   nx = (size_t)g.nx, ny = (size_t)g.ny;
    const size_t   var_bytes = nx * ny * sizeof(double);
    const off_t    prefix = (off_t)h->out_prefix_bytes;
    size_t         rows_per_chunk = ((size_t)32 << 20) / (4 * nx * sizeof(double)); // ~32 MB per chunk
    if (rows_per_chunk < 1)
      rows_per_chunk = 1;
    if (rows_per_chunk > ny)
      rows_per_chunk = ny;
    const size_t chunk_doubles = 4 * rows_per_chunk * nx;
    if (h->out_cap < chunk_doubles)
    {
      for (int b = 0; b < 2; ++b)
      {
        if (h->out_dev[b])
          cudaFree(h->out_dev[b]);
        if (h->out_host[b])
          cudaFreeHost(h->out_host[b]);
        h->out_dev[b] = nullptr;
        h->out_host[b] = nullptr;
      }
      h->out_cap = 0;
      for (int b = 0; b < 2; ++b)
      {
        E2D_CUDA(cudaMalloc(&h->out_dev[b], chunk_doubles * sizeof(double)));
        E2D_CUDA(cudaMallocHost(&h->out_host[b], chunk_doubles * sizeof(double)));
        if (!h->out_ev[b])
          E2D_CUDA(cudaEventCreateWithFlags(&h->out_ev[b], cudaEventDisableTiming));
      }
      h->out_cap = chunk_doubles;
    }
    const size_t nchunk = (ny + rows_per_chunk - 1) / rows_per_chunk;
    auto         issue = [&](size_t k) -> int {
      const int    b = (int)(k & 1);
      const size_t r0 = k * rows_per_chunk, nr = std::min(rows_per_chunk, ny - r0);
      E2D_CUDA(launch_gather_interior(g, A, h->out_dev[b], 2 + (int)r0, (int)nr, h->stream));
      E2D_CUDA(cudaMemcpyAsync(h->out_host[b], h->out_dev[b], 4 * nr * nx * sizeof(double), cudaMemcpyDeviceToHost,
                               h->stream));
      E2D_CUDA(cudaEventRecord(h->out_ev[b], h->stream));
      return E2D_OK;
    };
    if (int rc = issue(0))
      return rc;
    for (size_t k = 0; k < nchunk; ++k)
    {
      if (k + 1 < nchunk) // the next block travels while this one is written
        if (int rc = issue(k + 1))
          return rc;
      const int    b = (int)(k & 1);
      const size_t r0 = k * rows_per_chunk, nr = std::min(rows_per_chunk, ny - r0);
      E2D_CUDA(cudaEventSynchronize(h->out_ev[b]));
      for (int v = 0; v < 4; ++v)
      {
        const char * src = (const char *)(h->out_host[b] + (size_t)v * nr * nx);
        size_t       left = nr * nx * sizeof(double);
        off_t        off = base + (off_t)v * (prefix + (off_t)var_bytes) + prefix + (off_t)(r0 * nx * sizeof(double));
        while (left > 0)
        {
          const ssize_t w = pwrite(fd, src, left, off);
          if (w <= 0)
            return fail(E2D_ERR_IO, "write failed: " + fname);
          src += w;
          off += w;
          left -= (size_t)w;
        }
      }
    }
    return E2D_OK;
  }

  int
  e2d_save_raw(e2d_handle * h, int which, const char * path)
  {
    double * A = h ? array_of(h, which) : nullptr;
    if (!A || !path)
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device);
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0)
      return fail(E2D_ERR_IO, std::string("cannot open ") + path);
    h->out_prefix_bytes = 0;
    const int rc = stream_interior_to_file(h, A, fd, 0, path);
    return (close(fd) != 0 && rc == E2D_OK) ? fail(E2D_ERR_IO, std::string("close failed: ") + path) : rc;
  }

  int
  e2d_save_vtk_appended(e2d_handle * h, int which, int iStep)
  {
    double * A = h ? array_of(h, which) : nullptr;
    if (!A)
      return fail(E2D_ERR_INVALID, "bad argument");
    if (!h->whole)
      return fail(E2D_ERR_UNSUPPORTED, "saveData on a y-slab handle: gather the interior and write it from one rank");
    cudaSetDevice(h->device);
    const e2d_params & p = h->p;
    const int          nx = p.nx, ny = h->g.ny;
    std::ostringstream stepNum; // HydroRun.h:537-544
    stepNum.width(7);
    stepNum.fill('0');
    stepNum << iStep;
    const std::string filename = std::string(p.outputDir) + "/" + p.outputPrefix + "_" + stepNum.str() + ".vti";
    const unsigned long long var_bytes = (unsigned long long)nx * ny * sizeof(double);
    std::ostringstream       hd;
    hd << "<?xml version=\"1.0\"?>\n";
    hd << "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
    hd << "  <ImageData WholeExtent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << 0 << "\" "
       << "Origin=\"" << p.xmin << " " << p.ymin << " " << 0.0 << "\" "
       << "Spacing=\"" << p.dx << " " << p.dy << " " << 0.0 << "\">\n";
    hd << "  <Piece Extent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << 0 << " "
       << "\">\n";
    hd << "    <PointData>\n    </PointData>\n    <CellData>\n";
    static const char * varNames[4] = { "rho", "E", "mx", "my" }; // HydroParams.cpp:17
    for (int v = 0; v < 4; ++v)
      hd << "    <DataArray type=\"Float64\" Name=\"" << varNames[v] << "\" format=\"appended\" offset=\""
         << (unsigned long long)v * (8ull + var_bytes) << "\" />\n";
    hd << "    </CellData>\n  </Piece>\n  </ImageData>\n";
    hd << "  <AppendedData encoding=\"raw\">\n   _";
    const std::string head = hd.str();
    const std::string tail = "\n  </AppendedData>\n</VTKFile>\n";

    const int fd = open(filename.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0)
      return fail(E2D_ERR_IO, "cannot open " + filename);
    int rc = E2D_OK;
    if (pwrite(fd, head.data(), head.size(), 0) != (ssize_t)head.size())
      rc = fail(E2D_ERR_IO, "write failed: " + filename);
    const off_t base = (off_t)head.size();
    for (int v = 0; v < 4 && rc == E2D_OK; ++v) // each array is preceded by its byte count (UInt64)
      if (pwrite(fd, &var_bytes, 8, base + (off_t)v * (8 + (off_t)var_bytes)) != 8)
        rc = fail(E2D_ERR_IO, "write failed: " + filename);
    if (rc == E2D_OK)
    {
      h->out_prefix_bytes = 8;
      rc = stream_interior_to_file(h, A, fd, base, filename);
    }
    if (rc == E2D_OK &&
        pwrite(fd, tail.data(), tail.size(), base + 4 * (8 + (off_t)var_bytes)) != (ssize_t)tail.size())
      rc = fail(E2D_ERR_IO, "write failed: " + filename);
    if (close(fd) != 0 && rc == E2D_OK)
      rc = fail(E2D_ERR_IO, "close failed: " + filename);
    return rc;
  }

  // ---- Sedov post-processing ----
  int
  e2d_compute_radial_profile(e2d_handle * h, int which, int nbins, double * distances, double * sums, int * counts)
  {
    double * A = h ? array_of(h, which) : nullptr;
    if (!A)
      return fail(E2D_ERR_INVALID, "bad argument");
    cudaSetDevice(h->device);
    const e2d_params & p = h->p;
    if (nbins <= 0)
      nbins = p.blast_nbins;
    if (nbins <= 0)
      return fail(E2D_ERR_INVALID, "nbins must be positive");
    // rows this handle owns: the interior rows, plus the ghost rows of the physical y faces (the reference bins
    // every cell of the global array, ghost cells included)
    const bool first = h->whole || h->slab.rank == 0, last = h->whole || h->slab.rank == h->slab.nranks - 1;
    const int  j_lo = first ? 0 : 2, j_hi = last ? h->g.jsize : h->g.jsize - 2;
    const RadialArgs a = make_radial_args(p, nbins, j_lo, j_hi);
    const int        nseg = radial_segments(h->g, nbins, j_hi - j_lo);
    const size_t     n = (size_t)nbins * nseg * h->g.isize;
    double *         part_sum = nullptr;
    int *            part_cnt = nullptr;
    double *         d_sums = nullptr;
    int *            d_counts = nullptr;
    std::vector<double> h_sums(nbins);
    std::vector<int>    h_counts(nbins);
    cudaError_t         e = cudaMalloc(&part_sum, n * sizeof(double));
    if (e == cudaSuccess)
      e = cudaMalloc(&part_cnt, n * sizeof(int));
    if (e == cudaSuccess)
      e = cudaMalloc(&d_sums, nbins * sizeof(double));
    if (e == cudaSuccess)
      e = cudaMalloc(&d_counts, nbins * sizeof(int));
    if (e == cudaSuccess)
      e = launch_radial_profile(h->g, a, A, nseg, part_sum, part_cnt, d_sums, d_counts, h->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(h_sums.data(), d_sums, nbins * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(h_counts.data(), d_counts, nbins * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(h->stream);
    cudaFree(part_sum);
    cudaFree(part_cnt);
    cudaFree(d_sums);
    cudaFree(d_counts);
    E2D_CUDA(e);
    const double dr = a.rmax / nbins; // ComputeRadialProfileFunctor.h:125
    for (int k = 0; k < nbins; ++k)
    {
      if (distances)
        distances[k] = (k + 0.5) * dr;
      if (sums)
        sums[k] = h_sums[k];
      if (counts)
        counts[k] = h_counts[k];
    }
    return E2D_OK;
  }

  int
  e2d_save_npy(const char * path, const double * data, long n)
  {
    if (!path || (!data && n > 0) || n < 0)
      return fail(E2D_ERR_INVALID, "bad argument");
    // NumPy format 1.0: magic, version, little-endian uint16 header length, dict padded with spaces to a
    // multiple of 16 bytes and terminated by '\n' (cnpy::create_npy_header)
    std::string dict = "{'descr': '<f8', 'fortran_order': False, 'shape': (" + std::to_string(n) + ",), }";
    const size_t unpadded = 10 + dict.size() + 1;
    dict.append((16 - unpadded % 16) % 16, ' ');
    dict.push_back('\n');
    const unsigned short hl = (unsigned short)dict.size();
    FILE *               f = std::fopen(path, "wb");
    if (!f)
      return fail(E2D_ERR_IO, std::string("cannot open ") + path);
    const unsigned char magic[8] = { 0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0 };
    bool                ok = std::fwrite(magic, 1, 8, f) == 8;
    const unsigned char hlb[2] = { (unsigned char)(hl & 0xff), (unsigned char)(hl >> 8) };
    ok = ok && std::fwrite(hlb, 1, 2, f) == 2;
    ok = ok && std::fwrite(dict.data(), 1, dict.size(), f) == dict.size();
    ok = ok && (n == 0 || std::fwrite(data, sizeof(double), (size_t)n, f) == (size_t)n);
    ok = (std::fclose(f) == 0) && ok;
    return ok ? (int)E2D_OK : fail(E2D_ERR_IO, std::string("write failed: ") + path);
  }

  int
  e2d_save_radial_profile(e2d_handle * h, int which, const char * dir)
  {
    if (!h)
      return fail(E2D_ERR_INVALID, "bad argument");
    if (!h->whole && h->slab.nranks > 1)
      return fail(E2D_ERR_INVALID, "e2d_save_radial_profile: whole-domain handles only (add the slabs' partial "
                                   "sums from e2d_compute_radial_profile over the ranks)");
    const int           nbins = h->p.blast_nbins;
    std::vector<double> dist(nbins), sums(nbins);
    std::vector<int>    cnt(nbins);
    if (int rc = e2d_compute_radial_profile(h, which, nbins, dist.data(), sums.data(), cnt.data()))
      return rc;
    for (int k = 0; k < nbins; ++k)
      sums[k] /= cnt[k]; // ComputeRadialProfileFunctor.h:130 (0/0 = NaN for an empty bin, like the reference)
    const std::string d = (dir && *dir) ? std::string(dir) + "/" : std::string();
    if (int rc = e2d_save_npy((d + "sedov_blast_radial_distances.npy").c_str(), dist.data(), nbins)) // :134
      return rc;
    return e2d_save_npy((d + "sedov_blast_density_profile.npy").c_str(), sums.data(), nbins); // :135
  }

  int
  e2d_profile_enable(int on)
  {
    g_profile.store(on ? 1 : 0, std::memory_order_relaxed);
    return E2D_OK;
  }

  void
  e2d_profile_push(const char * name)
  {
    if (!profiling())
      return;
    nvtxRangePushA(name ? name : "");
    ++t_profile_depth;
    g_profile_ranges.fetch_add(1, std::memory_order_relaxed);
  }

  void
  e2d_profile_pop(void)
  {
    if (!profiling() || t_profile_depth <= 0)
      return;
    nvtxRangePop();
    --t_profile_depth;
  }

  int
  e2d_profile_stats(unsigned long long * ranges_opened, int * depth)
  {
    if (ranges_opened)
      *ranges_opened = g_profile_ranges.load(std::memory_order_relaxed);
    if (depth)
      *depth = t_profile_depth;
    return profiling() ? 1 : 0;
  }

  int
  e2d_enable_timers(e2d_handle * h, int on)
  {
    if (!h)
      return fail(E2D_ERR_INVALID, "bad argument");
    h->timing = on != 0;
    return E2D_OK;
  }

  int
  e2d_get_timers(e2d_handle * h, double out[5])
  {
    if (!h || !out)
      return fail(E2D_ERR_INVALID, "bad argument");
    // boundaries, godunov, primitive, fluxes, update  (HydroRun.h:74-75)
    for (int k = 0; k < 5; ++k)
      out[k] = h->timers[k];
    return E2D_OK;
  }

} // extern "C"
