// Source-level compatibility for hosts written against the reference (src/kokkos_shared.h): with this directory on
// the include path, /root/reference/src/main.cpp compiles unmodified and drives libeuler2d_b200.so.  Only the names
// that a HOST program of the reference touches are provided — this is not a Kokkos implementation, no kernel code
// can be written against it.
//   Kokkos::initialize / finalize / print_configuration / hwloc::available      (src/main.cpp:43-58,212)
//   Kokkos::Device<ExecSpace, MemSpace>, Kokkos::DefaultExecutionSpace            (src/main.cpp:27-29)
//   Kokkos::Profiling::pushRegion / popRegion -> NVTX ranges (e2d_profile_push)   (src/main.cpp:93,111,123,165)
#ifndef EULER2D_COMPAT_KOKKOS_SHARED_H
#define EULER2D_COMPAT_KOKKOS_SHARED_H

#include <iostream>
#include <sstream>
#include <string>

#include "../euler2d_b200.h"

namespace Kokkos
{
struct B200Space
{
  using memory_space = B200Space;
  using execution_space = B200Space;
};
using DefaultExecutionSpace = B200Space;
template <class ExecSpace, class MemSpace>
struct Device
{
  using execution_space = ExecSpace;
  using memory_space = MemSpace;
};
inline void
initialize(int &, char *[])
{}
inline void
finalize()
{}
inline void
print_configuration(std::ostream & os, bool = false)
{
  os << e2d_version() << ", " << e2d_device_count() << " CUDA device(s)\n";
}
namespace hwloc
{
inline bool
available()
{
  return false;
}
inline unsigned
get_available_numa_count()
{
  return 1;
}
inline unsigned
get_available_cores_per_numa()
{
  return 1;
}
inline unsigned
get_available_threads_per_core()
{
  return 1;
}
} // namespace hwloc
namespace Profiling
{
inline void
pushRegion(const std::string & name)
{
  e2d_profile_push(name.c_str());
}
inline void
popRegion()
{
  e2d_profile_pop();
}
} // namespace Profiling
} // namespace Kokkos

#endif
