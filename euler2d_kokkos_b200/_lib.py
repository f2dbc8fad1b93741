"""ctypes binding of libeuler2d_b200.so (include/euler2d_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C euler2d_kokkos_b200/csrc``.
There is no fallback: if the shared library is missing, importing a compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# E2D_LIB_PATH: development aid (tools/: A/B runs of differently built libraries); the default is the in-tree build
LIB_PATH = os.environ.get("E2D_LIB_PATH") or os.path.join(HERE, "libeuler2d_b200.so")

E2D_OK = 0
E2D_U, E2D_U2, E2D_Q = 0, 1, 2
LAYOUT_SOA, LAYOUT_KOKKOS_OMP = 0, 1
FACES_X, FACES_YMIN, FACES_YMAX, FACES_ALL = 3, 4, 8, 15
BC_DIRICHLET, BC_NEUMANN, BC_PERIODIC = 1, 2, 3
ID, IP, IE, IU, IV = 0, 1, 1, 2, 3


class E2dError(RuntimeError):
    pass


class Params(C.Structure):
    """Mirror of ``e2d_params``; field-for-field HydroParams + HydroSettings (src/HydroParams.h:107-265)."""

    _fields_ = (
        [("nStepmax", C.c_int), ("tEnd", C.c_double), ("nOutput", C.c_int), ("enableOutput", C.c_int)]
        + [(n, C.c_int) for n in ("nx", "ny", "ghostWidth", "imin", "imax", "jmin", "jmax", "isize", "jsize")]
        + [(n, C.c_double) for n in ("xmin", "xmax", "ymin", "ymax", "dx", "dy")]
        + [(n, C.c_int) for n in ("boundary_type_xmin", "boundary_type_xmax", "boundary_type_ymin",
                                  "boundary_type_ymax", "ioVTK", "ioHDF5")]
        + [(n, C.c_double) for n in ("gamma0", "gamma6", "cfl", "slope_type", "smallr", "smallc", "smallp",
                                     "smallpp")]
        + [(n, C.c_int) for n in ("niter_riemann", "riemannSolverType", "problemType")]
        + [(n, C.c_double) for n in ("blast_radius", "blast_center_x", "blast_center_y", "blast_density_in",
                                     "blast_density_out", "blast_pressure_in", "blast_pressure_out",
                                     "blast_total_energy_inside")]
        + [("blast_nbins", C.c_int)]
        + [(n, C.c_double) for n in ("bubble_radius", "bubble_center_x", "bubble_center_y", "bubble_density",
                                     "bubble_pressure", "preshock_density", "preshock_pressure",
                                     "postshock_density", "postshock_pressure", "postshock_velocity",
                                     "shock_loc")]
        + [("implementationVersion", C.c_int), ("outputDir", C.c_char * 256), ("outputPrefix", C.c_char * 256),
           ("honourRiemannSolver", C.c_int), ("vtkAppended", C.c_int), ("arithmetic", C.c_int),
           ("unfusedKernels", C.c_int)]
    )

    def as_dict(self):
        d = {}
        for n, _ in self._fields_:
            v = getattr(self, n)
            d[n] = v.decode() if isinstance(v, bytes) else v
        return d

    def copy(self) -> "Params":
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


class Slab(C.Structure):
    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("ny_loc", C.c_int), ("j_off", C.c_int)]


class RunStats(C.Structure):
    _fields_ = [("nStep", C.c_int), ("t", C.c_double), ("dt_last", C.c_double), ("seconds", C.c_double),
                ("launches", C.c_longlong), ("seconds_step_kernel", C.c_double)]


class IpcBlob(C.Structure):
    """Mirror of ``e2d_ipc_blob``: the CUDA IPC handles of one rank's U, U2 and comm block."""

    _fields_ = [("U", C.c_ubyte * 64), ("U2", C.c_ubyte * 64), ("comm", C.c_ubyte * 64), ("rank", C.c_int),
                ("nranks", C.c_int), ("ny_loc", C.c_int), ("device", C.c_int)]


# every symbol include/euler2d_b200.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_pp = C.POINTER(Params)
_vp = C.c_void_p
SIGNATURES = {
    "e2d_version": (C.c_char_p, []),
    "e2d_status_string": (C.c_char_p, [C.c_int]),
    "e2d_last_error": (C.c_char_p, []),
    "e2d_device_count": (C.c_int, []),
    "e2d_kernel_launch_count": (C.c_ulonglong, []),
    "e2d_params_from_ini": (C.c_int, [C.c_char_p, _pp]),
    "e2d_params_from_string": (C.c_int, [C.c_char_p, _pp]),
    "e2d_params_init": (C.c_int, [_pp]),
    "e2d_params_print": (C.c_int, [_pp]),
    "e2d_k_init_problem": (C.c_int, [_pp, _vp, C.c_int, C.c_int, _vp]),
    "e2d_k_make_boundaries": (C.c_int, [_pp, _vp, C.c_int, C.c_int, _vp]),
    "e2d_k_reduce_invdt": (C.c_int, [_pp, _vp, C.c_int, _vp, _vp]),
    "e2d_k_convert_to_primitives": (C.c_int, [_pp, _vp, _vp, C.c_int, _vp]),
    "e2d_k_compute_and_store_fluxes": (C.c_int, [_pp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int, _vp]),
    "e2d_k_update": (C.c_int, [_pp, _vp, _vp, _vp, C.c_int, _vp]),
    "e2d_k_compute_slopes": (C.c_int, [_pp, _vp, _vp, _vp, C.c_int, _vp]),
    "e2d_k_compute_trace_and_fluxes": (C.c_int, [_pp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int,
                                                 C.c_int, _vp]),
    "e2d_k_update_dir": (C.c_int, [_pp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "e2d_k_fused_step": (C.c_int, [_pp, _vp, _vp, C.c_int, C.c_double, _vp, _vp, _vp, _vp]),
    "e2d_k_eval_host": (C.c_int, [_pp, C.c_char_p, _dp, _dp, C.c_long]),
    "e2d_create": (C.c_int, [_pp, C.POINTER(Slab), _vp, _vp, _vp, C.POINTER(_vp)]),
    "e2d_destroy": (C.c_int, [_vp]),
    "e2d_compute_dt": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "e2d_make_boundaries": (C.c_int, [_vp, C.c_int]),
    "e2d_godunov_unsplit": (C.c_int, [_vp, C.c_int, C.c_double]),
    "e2d_godunov_unsplit_nobc": (C.c_int, [_vp, C.c_int, C.c_double]),
    "e2d_run": (C.c_int, [_vp, C.c_long, C.POINTER(RunStats)]),
    "e2d_ipc_export": (C.c_int, [_vp, C.POINTER(IpcBlob)]),
    "e2d_ipc_connect": (C.c_int, [_vp, C.POINTER(IpcBlob), C.c_int]),
    "e2d_peer_connect_local": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "e2d_get_dt_history": (C.c_int, [_vp, _dp, C.c_long, C.POINTER(C.c_long)]),
    "e2d_set_time": (C.c_int, [_vp, C.c_double, C.c_int]),
    "e2d_download": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "e2d_upload": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "e2d_device_ptr": (_vp, [_vp, C.c_int]),
    "e2d_stream": (_vp, [_vp]),
    "e2d_synchronize": (C.c_int, [_vp]),
    "e2d_get_params": (C.c_int, [_vp, _pp]),
    "e2d_step_host": (C.c_int, [_vp, _vp, _vp, _dp]),
    "e2d_step_host_streamed": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_int, _dp, _dp]),
    "e2d_march_host": (C.c_int, [_vp, _vp, _vp, C.c_long, C.c_int, _dp, _dp]),
    "e2d_save_vtk": (C.c_int, [_vp, C.c_int, C.c_int]),
    "e2d_save_vtk_appended": (C.c_int, [_vp, C.c_int, C.c_int]),
    "e2d_save_raw": (C.c_int, [_vp, C.c_int, C.c_char_p]),
    "e2d_compute_radial_profile": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_int)]),
    "e2d_save_radial_profile": (C.c_int, [_vp, C.c_int, C.c_char_p]),
    "e2d_save_npy": (C.c_int, [C.c_char_p, _dp, C.c_long]),
    "e2d_enable_timers": (C.c_int, [_vp, C.c_int]),
    "e2d_get_timers": (C.c_int, [_vp, _dp]),
    "e2d_blast_inside_count": (C.c_int, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]),
    "e2d_blast_renormalise": (C.c_int, [_vp, C.c_ulonglong]),
    "e2d_config_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "e2d_config_from_string": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "e2d_config_close": (None, [_vp]),
    "e2d_config_parse_error": (C.c_int, [_vp]),
    "e2d_config_get_float": (C.c_float, [_vp, C.c_char_p, C.c_char_p, C.c_float]),
    "e2d_config_get_integer": (C.c_long, [_vp, C.c_char_p, C.c_char_p, C.c_long]),
    "e2d_config_get_bool": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_int]),
    "e2d_config_get_string": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    "e2d_config_set_string": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_char_p]),
    "e2d_params_setup": (C.c_int, [_pp, _vp]),
    "e2d_profile_enable": (C.c_int, [C.c_int]),
    "e2d_profile_push": (None, [C.c_char_p]),
    "e2d_profile_pop": (None, []),
    "e2d_profile_stats": (C.c_int, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise E2dError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C euler2d_kokkos_b200/csrc`). euler2d_kokkos_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, what: str = "") -> None:
    if status != E2D_OK:
        L = lib()
        msg = L.e2d_last_error().decode()
        raise E2dError(f"{what or 'euler2d_b200'}: {L.e2d_status_string(status).decode()}" + (f" — {msg}" if msg else ""))
