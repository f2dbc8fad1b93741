// TEST INFRASTRUCTURE ONLY — see Kokkos_Core.hpp in this directory.
#include "Kokkos_Core.hpp"
