"""euler2d_kokkos_b200 — B200-native (sm_100a CUDA) unsplit MUSCL-Hancock Godunov step with the
driver surface of pkestene/euler2d_kokkos (HydroParams / HydroRun), over a C ABI
(include/euler2d_b200.h, libeuler2d_b200.so).  No CPU fallback."""
from ._lib import (BC_DIRICHLET, BC_NEUMANN, BC_PERIODIC, E2D_Q, E2D_U, E2D_U2, FACES_ALL, FACES_X, FACES_YMAX,
                   FACES_YMIN, LAYOUT_KOKKOS_OMP, LAYOUT_SOA, LIB_PATH, E2dError, Params, RunStats, Slab, check, lib)
from .hydro_run import ConfigMap, HydroParams, HydroRun, main_loop

__all__ = ["ConfigMap", "HydroParams", "HydroRun", "main_loop", "Params", "Slab", "RunStats", "E2dError", "lib", "check",
           "LIB_PATH", "E2D_U", "E2D_U2", "E2D_Q", "LAYOUT_SOA", "LAYOUT_KOKKOS_OMP", "FACES_X", "FACES_YMIN",
           "FACES_YMAX", "FACES_ALL", "BC_DIRICHLET", "BC_NEUMANN", "BC_PERIODIC"]
