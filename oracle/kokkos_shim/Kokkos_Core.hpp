// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Minimal host-only stand-in for the subset of the Kokkos API that the reference's
// hot-path headers use (/root/reference/src/HydroRun.h, HydroRunFunctors.h,
// HydroBaseFunctor.h, kokkos_shared.h, real_type.h).  It lets `oracle/Makefile`
// compile those reference sources *unmodified, where they lie* with plain g++
// (no cmake, no generated config header), so the reference's own arithmetic is what
// runs in oracle/_ref/.  Only the loop runner is substituted:
//
//   * View<T**[N]>       : LayoutRight (i slowest, var fastest) — the layout Kokkos picks
//                          for the OpenMP backend.
//   * parallel_for       : OpenMP static loop over the first index.
//   * parallel_reduce    : per-thread partials joined in thread order (Max is exact and
//                          order independent; Sum is only used by the Sedov initialiser).
//
// Cross-validated bit-for-bit against a real Kokkos 5.1.0/OpenMP build of the same
// sources (see oracle/README.md).
#ifndef E2D_ORACLE_KOKKOS_SHIM_CORE_HPP
#define E2D_ORACLE_KOKKOS_SHIM_CORE_HPP

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <functional>
#include <type_traits>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#ifdef _OPENMP
#  include <omp.h>
#endif

#include "Kokkos_Macros.hpp"

namespace Kokkos
{

struct HostSpace
{};

struct OpenMP
{
  using execution_space = OpenMP;
  using memory_space = HostSpace;
  static const char *
  name()
  {
    return "OpenMP(shim)";
  }
};
using DefaultExecutionSpace = OpenMP;
using DefaultHostExecutionSpace = OpenMP;

template <class Exec, class Mem>
struct Device
{
  using execution_space = Exec;
  using memory_space = Mem;
  using device_type = Device;
};

inline void
initialize(int &, char **)
{}
inline void
initialize()
{}
inline void
finalize()
{}
inline void
fence()
{}
inline void
print_configuration(std::ostream & os, bool = false)
{
  os << "Kokkos API shim (oracle/kokkos_shim), OpenMP loops\n";
}
namespace hwloc
{
inline bool
available()
{
  return false;
}
inline int
get_available_numa_count()
{
  return 1;
}
inline int
get_available_cores_per_numa()
{
  return 1;
}
inline int
get_available_threads_per_core()
{
  return 1;
}
} // namespace hwloc

namespace Profiling
{
inline void
pushRegion(const std::string &)
{}
inline void
popRegion()
{}
} // namespace Profiling

// ---------------------------------------------------------------- Array
template <class T, std::size_t N>
struct Array
{
  T m_internal_implementation_private_member_data[N];
  T &
  operator[](std::size_t k)
  {
    return m_internal_implementation_private_member_data[k];
  }
  const T &
  operator[](std::size_t k) const
  {
    return m_internal_implementation_private_member_data[k];
  }
};

// ---------------------------------------------------------------- View
struct WithoutInitializing_t
{};
constexpr WithoutInitializing_t WithoutInitializing{};

struct ViewAllocProp
{
  std::string label;
  bool        init;
};
inline ViewAllocProp
view_alloc(WithoutInitializing_t, const std::string & label)
{
  return { label, false };
}
inline ViewAllocProp
view_alloc(const std::string & label)
{
  return { label, true };
}

template <class DataType, class... Props>
class View;

// rank-3 view with a compile-time last extent: T**[N], LayoutRight.
template <class T, std::size_t N, class... Props>
class View<T ** [N], Props...>
{
public:
  using host_mirror_type = View;
  using HostMirror = View;
  using value_type = T;

  View() = default;
  View(const std::string & label, std::size_t n0, std::size_t n1)
    : m_n0(n0)
    , m_n1(n1)
    , m_label(label)
    , m_data(new T[n0 * n1 * N](), std::default_delete<T[]>())
  {}
  View(const ViewAllocProp & prop, std::size_t n0, std::size_t n1)
    : m_n0(n0)
    , m_n1(n1)
    , m_label(prop.label)
    , m_data(prop.init ? new T[n0 * n1 * N]() : new T[n0 * n1 * N], std::default_delete<T[]>())
  {}

  T &
  operator()(std::size_t i, std::size_t j, std::size_t v) const
  {
    return m_data.get()[(i * m_n1 + j) * N + v];
  }
  std::size_t
  extent(int r) const
  {
    return r == 0 ? m_n0 : (r == 1 ? m_n1 : N);
  }
  std::size_t
  size() const
  {
    return m_n0 * m_n1 * N;
  }
  T *
  data() const
  {
    return m_data.get();
  }
  const std::string &
  label() const
  {
    return m_label;
  }

private:
  std::size_t        m_n0 = 0, m_n1 = 0;
  std::string        m_label;
  std::shared_ptr<T> m_data;
};

template <class V>
V
create_mirror_view(WithoutInitializing_t, const V & v)
{
  return V(view_alloc(WithoutInitializing, v.label() + "_mirror"), v.extent(0), v.extent(1));
}
template <class V>
V
create_mirror_view(const V & v)
{
  return V(v.label() + "_mirror", v.extent(0), v.extent(1));
}

template <class V>
void
deep_copy(const V & dst, const V & src)
{
  if (dst.data() != src.data())
    std::memcpy(dst.data(), src.data(), sizeof(typename V::value_type) * src.size());
}

// ---------------------------------------------------------------- rank-1 views (ComputeRadialProfileFunctor.h)
// View<T*, ...>: plain or with MemoryTraits<Atomic> (element access returns a proxy whose += is an OpenMP atomic).
// The reference's radial-profile functor indexes bins past the end for the corner ghost cells (bin >= nbins is
// never tested, ComputeRadialProfileFunctor.h:160-165); a real Kokkos allocation happens to absorb that, so this
// stand-in over-allocates by kRank1Slack elements instead of corrupting the heap.
enum MemoryTraitsFlags
{
  Unmanaged = 0x01,
  RandomAccess = 0x02,
  Atomic = 0x04,
  Restrict = 0x08,
  Aligned = 0x10
};
template <unsigned F>
struct MemoryTraits
{
  static constexpr unsigned flags = F;
};

template <class AccessSpace, class MemorySpace>
struct SpaceAccessibility
{
  static constexpr bool accessible = true; // everything lives in host memory here
};

namespace Impl
{
template <class T>
struct AtomicRef
{
  T * p;
  void
  operator+=(const T & v) const
  {
#pragma omp atomic
    *p += v;
  }
  operator T() const { return *p; }
};
template <class... Props>
struct has_atomic_trait : std::false_type
{};
template <class P, class... Rest>
struct has_atomic_trait<P, Rest...> : has_atomic_trait<Rest...>
{};
template <unsigned F, class... Rest>
struct has_atomic_trait<MemoryTraits<F>, Rest...> : std::integral_constant<bool, (F & Atomic) != 0 || has_atomic_trait<Rest...>::value>
{};
constexpr std::size_t kRank1Slack = 64;
} // namespace Impl

template <class T, class... Props>
class View<T *, Props...>
{
public:
  using value_type = T;
  using memory_space = HostSpace;
  static constexpr int  rank = 1;
  static constexpr bool is_atomic = Impl::has_atomic_trait<Props...>::value;

  View() = default;
  View(const std::string & label, std::size_t n0)
    : m_n0(n0)
    , m_label(label)
    , m_data(new T[n0 + Impl::kRank1Slack](), std::default_delete<T[]>())
  {}
  template <class... P2>
  View(const View<T *, P2...> & o) // same allocation seen through other traits
    : m_n0(o.extent(0))
    , m_label(o.label())
    , m_data(o.shared())
  {}

  decltype(auto)
  operator()(std::size_t i) const
  {
    if constexpr (is_atomic)
      return Impl::AtomicRef<T>{ m_data.get() + i };
    else
      return (m_data.get()[i]);
  }
  std::size_t
  extent(int r) const
  {
    return r == 0 ? m_n0 : 1;
  }
  std::size_t
  size() const
  {
    return m_n0;
  }
  T *
  data() const
  {
    return m_data.get();
  }
  const std::string &
  label() const
  {
    return m_label;
  }
  const std::shared_ptr<T> &
  shared() const
  {
    return m_data;
  }

private:
  std::size_t        m_n0 = 0;
  std::string        m_label;
  std::shared_ptr<T> m_data;
};

template <class T, class... Props>
void
deep_copy(const View<T *, Props...> & dst, const T & value)
{
  for (std::size_t i = 0; i < dst.size(); ++i)
    dst.data()[i] = value;
}
template <class T, class... Props>
void
deep_copy(const View<T *, Props...> & dst, int value) requires(!std::is_same<T, int>::value)
{
  for (std::size_t i = 0; i < dst.size(); ++i)
    dst.data()[i] = static_cast<T>(value);
}
template <class Space, class T, class... Props>
View<T *, HostSpace>
create_mirror_view_and_copy(const Space &, const View<T *, Props...> & src)
{
  View<T *, HostSpace> m(src.label() + "_mirror", src.extent(0));
  std::memcpy(m.data(), src.data(), sizeof(T) * src.size());
  return m;
}

// ---------------------------------------------------------------- policies
template <unsigned R>
struct Rank
{};

template <class Exec, class RankT, class Tag = void>
struct MDRangePolicy
{
  using work_tag = Tag;
  long lo[2], hi[2];
  MDRangePolicy(std::initializer_list<long> l, std::initializer_list<long> h)
  {
    std::copy(l.begin(), l.end(), lo);
    std::copy(h.begin(), h.end(), hi);
  }
};

template <class Exec, class Tag = void>
struct RangePolicy
{
  using work_tag = Tag;
  long lo, hi;
  RangePolicy(long l, long h)
    : lo(l)
    , hi(h)
  {}
};

// ---------------------------------------------------------------- reducers
template <class T>
struct Max
{
  using value_type = T;
  T & result;
  explicit Max(T & r)
    : result(r)
  {}
  static T
  identity()
  {
    return -std::numeric_limits<T>::infinity();
  }
  static void
  join(T & dst, const T & src)
  {
    if (src > dst)
      dst = src;
  }
};

template <class T>
struct Sum
{
  using value_type = T;
  T & result;
  explicit Sum(T & r)
    : result(r)
  {}
  static T
  identity()
  {
    return T(0);
  }
  static void
  join(T & dst, const T & src)
  {
    dst += src;
  }
};

namespace Impl
{
inline int
shim_num_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
} // namespace Impl

// ---------------------------------------------------------------- parallel_for
template <class Exec, class RankT, class F>
void
parallel_for(const std::string &, const MDRangePolicy<Exec, RankT, void> & p, const F & f)
{
#pragma omp parallel for schedule(static)
  for (long i = p.lo[0]; i < p.hi[0]; ++i)
    for (long j = p.lo[1]; j < p.hi[1]; ++j)
      f(static_cast<int>(i), static_cast<int>(j));
}

template <class Exec, class RankT, class Tag, class F>
void
parallel_for(const std::string &, const MDRangePolicy<Exec, RankT, Tag> & p, const F & f)
{
#pragma omp parallel for schedule(static)
  for (long i = p.lo[0]; i < p.hi[0]; ++i)
    for (long j = p.lo[1]; j < p.hi[1]; ++j)
      f(Tag{}, static_cast<int>(i), static_cast<int>(j));
}

template <class Exec, class F>
void
parallel_for(const std::string &, const RangePolicy<Exec, void> & p, const F & f)
{
#pragma omp parallel for schedule(static)
  for (long k = p.lo; k < p.hi; ++k)
    f(static_cast<int>(k));
}

// ---------------------------------------------------------------- parallel_reduce
namespace Impl
{
template <class Tag, class F, class T>
inline void
shim_call_reduce(const F & f, int i, int j, T & v)
{
  if constexpr (std::is_void<Tag>::value)
    f(i, j, v);
  else
    f(Tag{}, i, j, v);
}
} // namespace Impl

template <class Exec, class RankT, class Tag, class F, class Reducer>
void
parallel_reduce(const std::string &,
                const MDRangePolicy<Exec, RankT, Tag> & p,
                const F &                               f,
                Reducer                                 reducer)
{
  using T = typename Reducer::value_type;
  const int      nt = Impl::shim_num_threads();
  std::vector<T> partial(static_cast<std::size_t>(nt) * 8, Reducer::identity()); // padded
#pragma omp parallel
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    T local = Reducer::identity();
#pragma omp for schedule(static)
    for (long i = p.lo[0]; i < p.hi[0]; ++i)
      for (long j = p.lo[1]; j < p.hi[1]; ++j)
        Impl::shim_call_reduce<Tag>(f, static_cast<int>(i), static_cast<int>(j), local);
    partial[static_cast<std::size_t>(tid) * 8] = local;
  }
  T total = Reducer::identity();
  for (int t = 0; t < nt; ++t)
    Reducer::join(total, partial[static_cast<std::size_t>(t) * 8]);
  reducer.result = total;
}

// ---------------------------------------------------------------- atomics
template <class T>
inline void
atomic_add(T * dst, const T & v)
{
#pragma omp atomic
  *dst += v;
}
template <class T>
inline void
atomic_sub(T * dst, const T & v)
{
#pragma omp atomic
  *dst -= v;
}

// ---------------------------------------------------------------- math (std:: equivalents)
using std::exp;
using std::fabs;
using std::fmax;
using std::fmin;
using std::fmod;
using std::isnan;
using std::sqrt;

} // namespace Kokkos

#endif // E2D_ORACLE_KOKKOS_SHIM_CORE_HPP
