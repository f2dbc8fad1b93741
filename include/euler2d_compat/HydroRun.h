// src/HydroRun.h of the reference, served by the B200 library: euler2d::HydroRun<device_t>.
#ifndef EULER2D_COMPAT_HYDRORUN_H
#define EULER2D_COMPAT_HYDRORUN_H
#include "../../euler2d_kokkos_b200/csrc/HydroRun.hpp"
namespace euler2d = euler2d_b200;
#endif
