// C++ host-side mirror of the reference's driver class, over the C ABI.
//
//   euler2d_b200::HydroParams  <->  euler2d::HydroParams              (src/HydroParams.h:155-265)
//   euler2d_b200::HydroRun     <->  euler2d::HydroRun<device_t>       (src/HydroRun.h:44-134)
//
//   ConfigMap                  <->  ConfigMap                         (config/ConfigMap.h:26-46)
//
// Same class names, constructor arguments, member names and argument meaning as the reference:
// `ConfigMap configMap(file); params.setup(configMap); new HydroRun<device>(params, configMap)`,
// `hydro->compute_dt(nStep % 2)`, `hydro->make_boundaries(hydro->U)`, `hydro->godunov_unsplit(nStep, dt)`,
// `hydro->saveData(hydro->U, nStep, "U")`, the five public timers with .elapsed(),
// `ComputeRadialProfileFunctor<device>::apply(params, hydro->U)`.  With include/euler2d_compat/ on the include path
// (headers named like the reference's, `namespace euler2d = euler2d_b200`, a few Kokkos:: names) the reference's
// src/main.cpp compiles UNMODIFIED against this library — tests/test_gpu_refmain.py builds and runs exactly that.
// Arrays are named by the handles U / U2 instead of Kokkos views.  Error behaviour: the reference calls
// exit(EXIT_FAILURE) on fatal errors (HydroParams.cpp:179-184); so does this shim.
#ifndef EULER2D_B200_HYDRORUN_HPP
#define EULER2D_B200_HYDRORUN_HPP

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/euler2d_b200.h"

// ConfigMap (config/ConfigMap.h:26-46): global namespace in the reference, so also here.  Same constructor (a file
// name) and the same typed getters; a missing file gives an empty map (the reference never checks ParseError()).
#ifndef E2D_NO_CONFIGMAP
class ConfigMap
{
public:
  explicit ConfigMap(std::string filename)
  {
    e2d_config_open(filename.c_str(), &c_);
  }
  ~ConfigMap() { e2d_config_close(c_); }
  ConfigMap(const ConfigMap &) = delete;
  ConfigMap &
  operator=(const ConfigMap &) = delete;

  float
  getFloat(std::string section, std::string name, float default_value) const
  {
    return e2d_config_get_float(c_, section.c_str(), name.c_str(), default_value);
  }
  void
  setFloat(std::string section, std::string name, float value)
  {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.9g", (double)value);
    e2d_config_set_string(c_, section.c_str(), name.c_str(), buf);
  }
  bool
  getBool(std::string section, std::string name, bool default_value) const
  {
    return e2d_config_get_bool(c_, section.c_str(), name.c_str(), default_value ? 1 : 0) != 0;
  }
  void
  setBool(std::string section, std::string name, bool value)
  {
    e2d_config_set_string(c_, section.c_str(), name.c_str(), value ? "true" : "false");
  }
  // INIReader's interface (config/inih/INIReader.h)
  int
  ParseError() const
  {
    return e2d_config_parse_error(c_);
  }
  long
  getInteger(std::string section, std::string name, long default_value) const
  {
    return e2d_config_get_integer(c_, section.c_str(), name.c_str(), default_value);
  }
  void
  setInteger(std::string section, std::string name, long value)
  {
    e2d_config_set_string(c_, section.c_str(), name.c_str(), std::to_string(value).c_str());
  }
  std::string
  getString(std::string section, std::string name, std::string default_value) const
  {
    char buf[512];
    e2d_config_get_string(c_, section.c_str(), name.c_str(), default_value.c_str(), buf, sizeof buf);
    return buf;
  }
  void
  setString(std::string section, std::string name, std::string value)
  {
    e2d_config_set_string(c_, section.c_str(), name.c_str(), value.c_str());
  }
  const e2d_config *
  handle() const
  {
    return c_;
  }

private:
  e2d_config * c_ = nullptr;
};
#endif

namespace euler2d_b200
{

using real_t = double;

// enumerators of src/HydroParams.h:27-104 under the reference's names
enum ComponentIndex
{
  ID = E2D_ID,
  IP = E2D_IP,
  IE = E2D_IE,
  IU = E2D_IU,
  IV = E2D_IV
};
enum ProblemType
{
  PROBLEM_IMPLODE = E2D_PROBLEM_IMPLODE,
  PROBLEM_BLAST = E2D_PROBLEM_BLAST,
  PROBLEM_FOUR_QUADRANT = E2D_PROBLEM_FOUR_QUADRANT,
  PROBLEM_DISCONTINUITY = E2D_PROBLEM_DISCONTINUITY,
  PROBLEM_SHOCKED_BUBBLE = E2D_PROBLEM_SHOCKED_BUBBLE
};

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
  // HydroParams::setup(ConfigMap &) (src/HydroParams.cpp:43-155); prints the same line as :151
  void
  setup(ConfigMap & configMap)
  {
    e2d_params_setup(this, configMap.handle());
    std::printf("Using Euler implementation version %d\n", implementationVersion);
  }
  // convenience: the ConfigMap is built from the .ini path (a missing file leaves the defaults, like the reference)
  void
  setup(const std::string & ini_path)
  {
    ConfigMap configMap(ini_path);
    setup(configMap);
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

// Stands in for DataArray_t (a Kokkos view) in the method signatures: the array lives in device memory behind the
// handle of the run it belongs to.
struct DataArray
{
  e2d_handle * owner = nullptr;
  int          which = E2D_U;
};

// host wall-clock timer with the interface of the reference's SimpleTimer / CudaTimer (src/SimpleTimer.h)
class Timer
{
public:
  void
  start()
  {
    t0_ = std::chrono::steady_clock::now();
  }
  void
  stop()
  {
    total_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
  }
  double
  elapsed() const
  {
    return total_;
  }

private:
  std::chrono::steady_clock::time_point t0_{};
  double                                total_ = 0.0;
};

// euler2d::HydroRun<device_t> (src/HydroRun.h:44-134).  device_t is accepted and ignored: there is one device
// kind here.  Same constructor arguments as the reference: (params, configMap) (src/HydroRun.h:143, main.cpp:86).
template <class device_t = void>
class HydroRun
{
public:
  struct PhaseTimer
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
  DataArray     U, U2, Q;
  PhaseTimer    boundaries_timer, godunov_timer, compute_primitive_timer, comp_fluxes_timer, update_hydro_timer;

  HydroRun(HydroParams & p, ConfigMap & /*configMap*/, bool timers = true)
    : params(p)
  {
    create(timers);
  }
  explicit HydroRun(HydroParams & p, bool timers = true)
    : params(p)
  {
    create(timers);
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
  void
  create(bool timers)
  {
    check(e2d_create(&params, nullptr, nullptr, nullptr, nullptr, &h_), "HydroRun");
    U = DataArray{ h_, E2D_U };
    U2 = DataArray{ h_, E2D_U2 };
    Q = DataArray{ h_, E2D_Q };
    PhaseTimer * ts[5] = { &boundaries_timer, &godunov_timer, &compute_primitive_timer, &comp_fluxes_timer,
                           &update_hydro_timer };
    for (int k = 0; k < 5; ++k)
    {
      ts[k]->owner = this;
      ts[k]->slot = k;
    }
    e2d_enable_timers(h_, timers ? 1 : 0);
  }
  e2d_handle * h_ = nullptr;
};

// euler2d::ComputeRadialProfileFunctor<device_t> (src/ComputeRadialProfileFunctor.h): apply(params, Udata) bins the
// density of every cell by its distance from the box centre and writes sedov_blast_radial_distances.npy /
// sedov_blast_density_profile.npy into the current directory — same call as src/main.cpp:175-179.
template <class device_t = void>
struct ComputeRadialProfileFunctor
{
  static void
  apply(const HydroParams & /*params*/, DataArray Udata)
  {
    check(e2d_save_radial_profile(Udata.owner, Udata.which, nullptr), "ComputeRadialProfileFunctor::apply");
  }
};

} // namespace euler2d_b200

#endif
