// TEST INFRASTRUCTURE ONLY — drives the UNMODIFIED reference (euler2d::HydroRun from
// /root/reference/src/HydroRun.h) the way /root/reference/src/main.cpp:76-143 does, but
// dumps the conservative state raw (the reference's VTK writer keeps 6 digits only,
// SURVEY.md §8c trap 3) and can time the solver loop for the CPU baseline.
//
// usage: ref_dump <file.ini> [--out PREFIX] [--nstep N] [--warmup W] [--states K] [--radial]
//   --radial       after the loop, run the reference's Sedov post-processing exactly as main.cpp:175-179 does
//                  (ComputeRadialProfileFunctor::apply(params, hydro->U) — always U, whatever the step parity):
//                  writes sedov_blast_radial_distances.npy / sedov_blast_density_profile.npy into the cwd
//   --warmup W     the loop timer starts after W steps (bench.py's reference arm)
//   PREFIX.U.bin   final state, doubles, [var][j][i] (whole array incl. ghost cells)
//   PREFIX.dt.bin  dt used at every step (doubles, nStep entries) preceded by the dt of main.cpp:87
//   PREFIX.sN.bin  (with --states K) state after every K-th step, same layout
// stdout: one JSON line with nStep, t (hex), sizes, loop seconds and Mcell-updates/s.
//
// Builds against the real Kokkos or against oracle/kokkos_shim (same source).
#include <chrono>
#ifdef _OPENMP
#  include <omp.h>
#endif
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "HydroParams.h"
#include "HydroRun.h"
#include "ComputeRadialProfileFunctor.h"
#include "real_type.h"

using device = Kokkos::Device<Kokkos::DefaultExecutionSpace, Kokkos::DefaultExecutionSpace::memory_space>;
using real_t = euler2d::real_t;

static void
dump_state(euler2d::HydroRun<device> & h, bool useU2, const std::string & fname)
{
  Kokkos::deep_copy(h.Uhost, useU2 ? h.U2 : h.U);
  const int           isize = h.params.isize, jsize = h.params.jsize;
  std::vector<real_t> buf(static_cast<size_t>(isize) * jsize * 4);
  for (int v = 0; v < 4; ++v)
    for (int j = 0; j < jsize; ++j)
      for (int i = 0; i < isize; ++i)
        buf[(static_cast<size_t>(v) * jsize + j) * isize + i] = h.Uhost(i, j, v);
  FILE * f = fopen(fname.c_str(), "wb");
  if (!f)
  {
    fprintf(stderr, "ref_dump: cannot open %s\n", fname.c_str());
    exit(2);
  }
  fwrite(buf.data(), sizeof(real_t), buf.size(), f);
  fclose(f);
}

int
main(int argc, char * argv[])
{
  Kokkos::initialize(argc, argv);
  int rc = 0;
  {
    if (argc < 2)
    {
      fprintf(stderr, "usage: ref_dump <file.ini> [--out PREFIX] [--nstep N] [--states K]\n");
      return 2;
    }
    std::string ini = argv[1], out;
    long        nstep_override = -1;
    int         states_every = 0;
    int         warmup = 0;
    bool        radial = false;
    for (int a = 2; a < argc; ++a)
    {
      if (!strcmp(argv[a], "--out") && a + 1 < argc)
        out = argv[++a];
      else if (!strcmp(argv[a], "--nstep") && a + 1 < argc)
        nstep_override = atol(argv[++a]);
      else if (!strcmp(argv[a], "--states") && a + 1 < argc)
        states_every = atoi(argv[++a]);
      else if (!strcmp(argv[a], "--warmup") && a + 1 < argc)
        warmup = atoi(argv[++a]);
      else if (!strcmp(argv[a], "--radial"))
        radial = true;
    }

    ConfigMap            configMap(ini);
    euler2d::HydroParams params = euler2d::HydroParams();
    params.setup(configMap);
    if (nstep_override >= 0)
      params.nStepmax = static_cast<int>(nstep_override);

    euler2d::HydroRun<device> * hydro = new euler2d::HydroRun<device>(params, configMap);

    real_t              t = 0, dt = 0;
    int                 nStep = 0;
    std::vector<real_t> dts;

    dt = hydro->compute_dt(nStep % 2); // main.cpp:87
    dts.push_back(dt);
    hydro->make_boundaries(hydro->U); // main.cpp:90-91
    hydro->make_boundaries(hydro->U2);
    if (!out.empty() && states_every > 0)
      dump_state(*hydro, false, out + ".s0.bin");

    auto t0 = std::chrono::steady_clock::now();
    while (t < params.tEnd && nStep < params.nStepmax) // main.cpp:100
    {
      if (nStep == warmup)
        t0 = std::chrono::steady_clock::now();
      dt = hydro->compute_dt(nStep % 2); // main.cpp:128
      if (t + dt > params.tEnd)          // main.cpp:131-134
        dt = params.tEnd - t;
      hydro->godunov_unsplit(nStep, dt); // main.cpp:139
      nStep++;
      t += dt;
      dts.push_back(dt);
      if (!out.empty() && states_every > 0 && nStep % states_every == 0)
        dump_state(*hydro, nStep % 2 != 0, out + ".s" + std::to_string(nStep) + ".bin");
    }
    auto   t1 = std::chrono::steady_clock::now();
    double secs = std::chrono::duration<double>(t1 - t0).count();

    if (radial)
      euler2d::ComputeRadialProfileFunctor<device>::apply(params, hydro->U); // main.cpp:175-179
    if (!out.empty())
    {
      dump_state(*hydro, nStep % 2 != 0, out + ".U.bin");
      if (radial) // the array the profile was taken from
        dump_state(*hydro, false, out + ".Uradial.bin");
      FILE * f = fopen((out + ".dt.bin").c_str(), "wb");
      fwrite(dts.data(), sizeof(real_t), dts.size(), f);
      fclose(f);
    }
    const int    timed_steps = nStep - (warmup < nStep ? warmup : 0);
    const double cells_ghost = 1.0 * params.isize * params.jsize;
    const double cells = 1.0 * params.nx * params.ny;
    printf("{\"nstep\": %d, \"t_hex\": \"%a\", \"t\": %.17g, \"isize\": %d, \"jsize\": %d, "
           "\"nx\": %d, \"ny\": %d, \"loop_seconds\": %.6f, \"mcell_updates_per_s\": %.4f, "
           "\"mcell_updates_per_s_ref_style\": %.4f, \"dx_hex\": \"%a\", \"dy_hex\": \"%a\", "
           "\"gamma0_hex\": \"%a\", \"cfl_hex\": \"%a\", \"smallr_hex\": \"%a\", \"smallc_hex\": \"%a\", "
           "\"smallp_hex\": \"%a\", \"smallpp_hex\": \"%a\", \"gamma6_hex\": \"%a\", \"tend_hex\": \"%a\", "
           "\"impl\": %d, \"timed_steps\": %d, \"threads\": %d}\n",
           nStep,
           t,
           t,
           params.isize,
           params.jsize,
           params.nx,
           params.ny,
           secs,
           secs > 0 ? timed_steps * cells / secs * 1e-6 : 0.0,
           secs > 0 ? timed_steps * cells_ghost / secs * 1e-6 : 0.0,
           params.dx,
           params.dy,
           params.settings.gamma0,
           params.settings.cfl,
           params.settings.smallr,
           params.settings.smallc,
           params.settings.smallp,
           params.settings.smallpp,
           params.settings.gamma6,
           params.tEnd,
           params.implementationVersion,
           timed_steps,
#ifdef _OPENMP
           omp_get_max_threads()
#else
           1
#endif
    );
    delete hydro;
  }
  Kokkos::finalize();
  return rc;
}
