// C++ host-side mirror of the reference's driver class, over the C ABI.
//
//   euler2d_b200::HydroParams  <->  euler2d::HydroParams              (src/HydroParams.h:155-265)
//   euler2d_b200::HydroRun     <->  euler2d::HydroRun<device_t>       (src/HydroRun.h:44-134)
//
// Same member names and argument meaning as the reference, so src/main.cpp ports by changing the
// namespace: `hydro->compute_dt(nStep % 2)`, `hydro->make_boundaries(hydro->U)`,
// `hydro->godunov_unsplit(nStep, dt)`, `hydro->saveData(hydro->U, nStep, "U")`, the five public timers
// with .elapsed().  Arrays are named by the handles U / U2 instead of Kokkos views.  Error behaviour:
// the reference calls exit(EXIT_FAILURE) on fatal errors (HydroParams.cpp:179-184); so does this shim.
#ifndef EULER2D_B200_HYDRORUN_HPP
#define EULER2D_B200_HYDRORUN_HPP

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/euler2d_b200.h"

namespace euler2d_b200
{

using real_t = double;

inline void
check(int status, const char * what)
{
  if (status != E2D_OK)
  {
    std::fprintf(stderr, "%s: %s (%s)\n", what, e2d_status_string(status), e2d_last_error());
    std::exit(EXIT_FAILURE);
  }
}

struct HydroParams : e2d_params
{
  HydroParams() { e2d_params_from_string("", this); }
  // HydroParams::setup(ConfigMap&): the ConfigMap is the .ini path here
  void
  setup(const std::string & ini_path)
  {
    e2d_params_from_ini(ini_path.c_str(), this); // a missing file leaves the defaults, like the reference
    std::printf("Using Euler implementation version %d\n", implementationVersion);
  }
  void
  init()
  {
    e2d_params_init(this);
  }
  void
  print()
  {
    e2d_params_print(this);
  }
};

// array handle standing in for DataArray_t in the method signatures
struct DataArray
{
  int which;
};

class HydroRun
{
public:
  struct Timer
  {
    const HydroRun * owner = nullptr;
    int              slot = 0;
    double
    elapsed() const
    {
      double t[5];
      e2d_get_timers(owner->h_, t);
      return t[slot];
    }
  };

  HydroParams & params;
  DataArray     U{ E2D_U }, U2{ E2D_U2 }, Q{ E2D_Q };
  Timer         boundaries_timer, godunov_timer, compute_primitive_timer, comp_fluxes_timer, update_hydro_timer;

  explicit HydroRun(HydroParams & p, bool timers = true)
    : params(p)
  {
    check(e2d_create(&p, nullptr, nullptr, nullptr, nullptr, &h_), "HydroRun");
    Timer * ts[5] = { &boundaries_timer, &godunov_timer, &compute_primitive_timer, &comp_fluxes_timer,
                      &update_hydro_timer };
    for (int k = 0; k < 5; ++k)
    {
      ts[k]->owner = this;
      ts[k]->slot = k;
    }
    e2d_enable_timers(h_, timers ? 1 : 0);
  }
  ~HydroRun() { e2d_destroy(h_); }
  HydroRun(const HydroRun &) = delete;
  HydroRun &
  operator=(const HydroRun &) = delete;

  real_t
  compute_dt(int useU)
  {
    double dt = 0;
    check(e2d_compute_dt(h_, useU, &dt, nullptr), "compute_dt");
    return dt;
  }
  void
  make_boundaries(DataArray Udata)
  {
    check(e2d_make_boundaries(h_, Udata.which), "make_boundaries");
  }
  void
  godunov_unsplit(int nStep, real_t dt)
  {
    check(e2d_godunov_unsplit(h_, nStep, dt), "godunov_unsplit");
  }
  void
  saveData(DataArray Udata, int iStep, const std::string & /*name*/)
  {
    if (params.ioVTK)
      check(e2d_save_vtk(h_, Udata.which, iStep), "saveData");
  }
  // the whole loop of main.cpp:100-143 on the device (IO off)
  e2d_run_stats
  run(long max_steps = -1)
  {
    e2d_run_stats st{};
    check(e2d_run(h_, max_steps, &st), "run");
    return st;
  }
  std::vector<real_t>
  download(DataArray Udata, int layout = E2D_LAYOUT_SOA)
  {
    std::vector<real_t> host(static_cast<size_t>(params.isize) * params.jsize * 4);
    check(e2d_download(h_, Udata.which, host.data(), layout), "download");
    return host;
  }
  void
  synchronize()
  {
    check(e2d_synchronize(h_), "synchronize");
  }
  e2d_handle *
  handle()
  {
    return h_;
  }

private:
  e2d_handle * h_ = nullptr;
};

// euler2d::ComputeRadialProfileFunctor<device_t> (src/ComputeRadialProfileFunctor.h): apply() bins the density of
// every cell by its distance from the box centre and writes sedov_blast_radial_distances.npy /
// sedov_blast_density_profile.npy into the current directory.  The reference passes (params, hydro->U); the array
// lives behind the HydroRun handle here, so apply takes the run.
struct ComputeRadialProfileFunctor
{
  static void
  apply(const HydroParams & /*params*/, HydroRun & hydro, DataArray Udata)
  {
    check(e2d_save_radial_profile(hydro.handle(), Udata.which, nullptr), "ComputeRadialProfileFunctor::apply");
  }
};

} // namespace euler2d_b200

#endif
