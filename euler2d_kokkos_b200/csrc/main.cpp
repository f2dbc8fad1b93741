// euler2d_b200 <file.ini> [--device-loop] — the program of the reference's src/main.cpp on the B200 path:
// same command line (one .ini argument, main.cpp:65-71), same loop (main.cpp:86-143), same report lines
// (main.cpp:182-204).  --device-loop runs the step loop device-resident (e2d_run) when output is off.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "HydroRun.hpp"

using namespace euler2d_b200;

int
main(int argc, char * argv[])
{
  bool        device_loop = false;
  const char * ini = nullptr;
  int          nfiles = 0;
  for (int a = 1; a < argc; ++a)
  {
    if (!std::strcmp(argv[a], "--device-loop"))
      device_loop = true;
    else
    {
      ini = argv[a];
      ++nfiles;
    }
  }
  if (nfiles != 1)
  {
    std::fprintf(stderr, "Error: wrong number of argument; input filename must be "
                         "the only parameter on the command line\n");
    return EXIT_FAILURE;
  }
  std::cout << "##########################\n";
  std::cout << e2d_version() << ", " << e2d_device_count() << " CUDA device(s)\n";
  std::cout << "##########################\n";

  real_t t = 0, dt = 0;
  int    nStep = 0;

  // the reference's sequence, verbatim (src/main.cpp:76-86)
  ConfigMap   configMap(ini);
  HydroParams params = HydroParams();
  params.setup(configMap);
  params.print();

  using device = void;
  HydroRun<device> * hydro = new HydroRun<device>(params, configMap, /*timers=*/!device_loop);
  dt = hydro->compute_dt(nStep % 2);
  hydro->make_boundaries(hydro->U);
  hydro->make_boundaries(hydro->U2);

  e2d_profile_push("main_loop"); // Kokkos::Profiling::pushRegion("main_loop"), main.cpp:93 (NVTX, with E2D_PROFILE=1)
  std::cout << "Start computation....\n";
  double     t_io = 0, t_dt = 0;
  const auto t0 = std::chrono::steady_clock::now();
  auto       secs_since = [](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
  };

  if (device_loop && !(params.enableOutput && params.nOutput > 0))
  {
    e2d_run_stats st = hydro->run();
    nStep = st.nStep;
    t = st.t;
    dt = st.dt_last;
  }
  else
  {
    while (t < params.tEnd && nStep < params.nStepmax)
    {
      if (nStep % 10 == 0)
        std::printf("time step=%7d (dt=% 10.8f t=% 10.8f)\n", nStep, dt, t);
      if (params.enableOutput)
      {
        if (params.nOutput > 0 && nStep % params.nOutput == 0)
        {
          std::cout << "Output results at time t=" << t << " step " << nStep << " dt=" << dt << std::endl;
          const auto a = std::chrono::steady_clock::now();
          hydro->saveData(nStep % 2 == 0 ? hydro->U : hydro->U2, nStep, "U");
          t_io += secs_since(a);
        }
      }
      const auto a = std::chrono::steady_clock::now();
      dt = hydro->compute_dt(nStep % 2);
      if (t + dt > params.tEnd)
        dt = params.tEnd - t;
      t_dt += secs_since(a);
      hydro->godunov_unsplit(nStep, dt);
      nStep++;
      t += dt;
    }
    if (params.enableOutput && params.nOutput > 0)
    {
      std::cout << "Output results at time t=" << t << " step " << nStep << " dt=" << dt << std::endl;
      const auto a = std::chrono::steady_clock::now();
      hydro->saveData(nStep % 2 == 0 ? hydro->U : hydro->U2, nStep, "U");
      t_io += secs_since(a);
    }
  }
  hydro->synchronize();
  const double t_tot = secs_since(t0);
  e2d_profile_pop();

  // post-processing for Sedov blast (src/main.cpp:175-179: always hydro->U, whatever the parity of nStep)
  if (params.problemType == E2D_PROBLEM_BLAST && params.blast_total_energy_inside > 0)
    euler2d_b200::ComputeRadialProfileFunctor<device>::apply(params, hydro->U);

  const double t_comp = hydro->godunov_timer.elapsed(), t_prim = hydro->compute_primitive_timer.elapsed();
  const double t_flux = hydro->comp_fluxes_timer.elapsed(), t_update = hydro->update_hydro_timer.elapsed();
  const double t_bound = hydro->boundaries_timer.elapsed();
  std::printf("total           time : %5.3f secondes\n", t_tot);
  std::printf("godunov         time : %5.3f secondes %5.2f%%\n", t_comp, 100 * t_comp / t_tot);
  std::printf("compute dt      time : %5.3f secondes %5.2f%%\n", t_dt, 100 * t_dt / t_tot);
  std::printf("primitive       time : %5.3f secondes %5.2f%%\n", t_prim, 100 * t_prim / t_tot);
  std::printf("compute fluxes  time : %5.3f secondes %5.2f%%\n", t_flux, 100 * t_flux / t_tot);
  std::printf("update hydro    time : %5.3f secondes %5.2f%%\n", t_update, 100 * t_update / t_tot);
  std::printf("boundaries      time : %5.3f secondes %5.2f%%\n", t_bound, 100 * t_bound / t_tot);
  std::printf("io              time : %5.3f secondes %5.2f%%\n", t_io, 100 * t_io / t_tot);
  std::printf("Perf                 : %10.2f number of Mcell-updates/s\n",
              1.0 * nStep * params.isize * params.jsize / t_tot * 1e-6);
  std::printf("final: nStep=%d t=%a\n", nStep, t);
  delete hydro;
  return EXIT_SUCCESS;
}
