"""Host-side mirror of the reference's driver surface, over the C ABI.

``HydroParams`` <-> struct HydroParams (src/HydroParams.h:155-265, setup/init in HydroParams.cpp:43-190)
``HydroRun``    <-> class euler2d::HydroRun<device_t> (src/HydroRun.h:44-134): same method names and
                    argument meaning (compute_dt(useU), make_boundaries(Udata), godunov_unsplit(nStep, dt),
                    saveData(Udata, iStep, name)), plus ``run()`` = the loop of src/main.cpp:100-143 kept on
                    the device.
All compute happens in libeuler2d_b200.so on the GPU; numpy is only used to hand arrays to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import E2D_U, E2D_U2, E2D_Q, LAYOUT_SOA, LAYOUT_KOKKOS_OMP, Params, RunStats, Slab, check, lib


class ConfigMap:
    """config/ConfigMap.h:26-46 — the parsed .ini as a key/value map with the reference's typed getters."""

    def __init__(self, filename: str | None = None, text: str | None = None):
        self._c = C.c_void_p()
        if text is not None:
            check(lib().e2d_config_from_string(text.encode(), C.byref(self._c)), "e2d_config_from_string")
        else:
            lib().e2d_config_open(os.fsencode(filename or ""), C.byref(self._c))  # a missing file gives an empty map

    def __del__(self):
        try:
            if self._c:
                lib().e2d_config_close(self._c)
                self._c = C.c_void_p()
        except Exception:
            pass

    def ParseError(self) -> int:
        return lib().e2d_config_parse_error(self._c)

    def getFloat(self, section: str, name: str, default_value: float) -> float:
        return lib().e2d_config_get_float(self._c, section.encode(), name.encode(), default_value)

    def getInteger(self, section: str, name: str, default_value: int) -> int:
        return lib().e2d_config_get_integer(self._c, section.encode(), name.encode(), default_value)

    def getBool(self, section: str, name: str, default_value: bool) -> bool:
        return bool(lib().e2d_config_get_bool(self._c, section.encode(), name.encode(), int(default_value)))

    def getString(self, section: str, name: str, default_value: str) -> str:
        buf = C.create_string_buffer(512)
        lib().e2d_config_get_string(self._c, section.encode(), name.encode(), default_value.encode(), buf, 512)
        return buf.value.decode()

    def setString(self, section: str, name: str, value) -> None:
        check(lib().e2d_config_set_string(self._c, section.encode(), name.encode(), str(value).encode()))

    setFloat = setInteger = setString

    def setBool(self, section: str, name: str, value: bool) -> None:
        self.setString(section, name, "true" if value else "false")


class HydroParams:
    """Parameters read from an .ini with the reference's semantics (floats go through strtof)."""

    def __init__(self, raw: Params | None = None):
        object.__setattr__(self, "raw", raw if raw is not None else Params())

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def from_ini(cls, path: str, strict: bool = True) -> "HydroParams":
        p = Params()
        rc = lib().e2d_params_from_ini(path.encode(), C.byref(p))
        if rc != 0 and strict:
            check(rc, f"e2d_params_from_ini({path})")
        return cls(p)

    @classmethod
    def from_string(cls, text: str) -> "HydroParams":
        p = Params()
        check(lib().e2d_params_from_string(text.encode(), C.byref(p)), "e2d_params_from_string")
        return cls(p)

    def setup(self, configMap) -> None:  # HydroParams::setup(ConfigMap&); a path is accepted too
        p = Params()
        if isinstance(configMap, ConfigMap):
            check(lib().e2d_params_setup(C.byref(p), configMap._c), "e2d_params_setup")
        else:
            check(lib().e2d_params_from_ini(os.fsencode(configMap), C.byref(p)), "e2d_params_from_ini")
        object.__setattr__(self, "raw", p)

    def init(self) -> None:  # HydroParams::init()
        check(lib().e2d_params_init(C.byref(self.raw)), "e2d_params_init")

    def print(self) -> None:  # HydroParams::print()
        lib().e2d_params_print(C.byref(self.raw))

    def copy(self) -> "HydroParams":
        return HydroParams(self.raw.copy())

    # -- attribute passthrough ------------------------------------------------------------------
    def __getattr__(self, name):
        raw = object.__getattribute__(self, "raw")
        if name in dict(Params._fields_):
            v = getattr(raw, name)
            return v.decode() if isinstance(v, bytes) else v
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in dict(Params._fields_):
            setattr(self.raw, name, value.encode() if isinstance(value, str) else value)
        else:
            raise AttributeError(f"HydroParams has no field {name}")


class HydroRun:
    """Drop-in for euler2d::HydroRun<device_t> on one GPU (or one y-slab of a multi-GPU run)."""

    U = E2D_U
    U2 = E2D_U2
    Q = E2D_Q

    def __init__(self, params: HydroParams, configMap: "ConfigMap | None" = None, slab: Slab | None = None,
                 stream: int | None = None, U_ptr: int | None = None, U2_ptr: int | None = None):
        """HydroRun(params, configMap) as in src/HydroRun.h:143 (the map is only read through params here)."""
        if isinstance(configMap, Slab):  # older call order HydroRun(params, slab)
            configMap, slab = None, configMap
        self.params = params
        self._h = C.c_void_p()
        check(lib().e2d_create(C.byref(params.raw), C.byref(slab) if slab is not None else None,
                               U_ptr, U2_ptr, stream, C.byref(self._h)), "e2d_create")
        self.jsize_loc = (slab.ny_loc if slab is not None else params.ny) + 2 * params.ghostWidth
        self.isize = params.isize
        self.shape = (4, self.jsize_loc, self.isize)
        self._steps_taken = 0

    def close(self) -> None:
        if self._h:
            lib().e2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the reference's methods ----------------------------------------------------------------
    def compute_dt(self, useU: int) -> float:
        dt = C.c_double()
        check(lib().e2d_compute_dt(self._h, int(useU), C.byref(dt), None), "e2d_compute_dt")
        return dt.value

    def compute_invdt_local(self, useU: int) -> float:
        dt, inv = C.c_double(), C.c_double()
        check(lib().e2d_compute_dt(self._h, int(useU), C.byref(dt), C.byref(inv)), "e2d_compute_dt")
        return inv.value

    def make_boundaries(self, Udata: int) -> None:
        check(lib().e2d_make_boundaries(self._h, int(Udata)), "e2d_make_boundaries")

    def godunov_unsplit(self, nStep: int, dt: float, fill_boundaries: bool = True) -> None:
        fn = lib().e2d_godunov_unsplit if fill_boundaries else lib().e2d_godunov_unsplit_nobc
        check(fn(self._h, int(nStep), float(dt)), "e2d_godunov_unsplit")

    def saveData(self, Udata: int, iStep: int, name: str = "U") -> None:
        if self.params.ioVTK:
            check(lib().e2d_save_vtk(self._h, int(Udata), int(iStep)), "e2d_save_vtk")

    def save_vtk_appended(self, Udata: int, iStep: int) -> None:
        """Raw-binary appended .vti (full precision), same name/extents/arrays as saveVTK."""
        check(lib().e2d_save_vtk_appended(self._h, int(Udata), int(iStep)), "e2d_save_vtk_appended")

    def save_raw(self, Udata: int, path: str) -> None:
        """Interior cells as raw doubles [var][j][i]."""
        check(lib().e2d_save_raw(self._h, int(Udata), os.fsencode(path)), "e2d_save_raw")

    def radial_profile(self, Udata: int = E2D_U, nbins: int = 0):
        """ComputeRadialProfileFunctor: returns (distances, sums, counts); the saved profile is sums / counts."""
        n = int(nbins) if nbins > 0 else int(self.params.blast_nbins)
        dist, sums, counts = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        check(lib().e2d_compute_radial_profile(self._h, int(Udata), n, dist.ctypes.data_as(dp),
                                               sums.ctypes.data_as(dp), counts.ctypes.data_as(C.POINTER(C.c_int))),
              "e2d_compute_radial_profile")
        return dist, sums, counts

    def save_radial_profile(self, Udata: int = E2D_U, directory: str = "") -> None:
        """ComputeRadialProfileFunctor::apply incl. its two .npy files (main.cpp:175-179 passes U)."""
        check(lib().e2d_save_radial_profile(self._h, int(Udata), os.fsencode(directory)), "e2d_save_radial_profile")

    # -- device-resident loop -------------------------------------------------------------------
    def run(self, max_steps: int = -1) -> RunStats:
        st = RunStats()
        check(lib().e2d_run(self._h, int(max_steps), C.byref(st)), "e2d_run")
        self._steps_taken = max(self._steps_taken, int(st.nStep))
        return st

    def dt_history(self) -> np.ndarray:
        """dt of every step taken by ``run`` so far (the library keeps at most 2^20 entries)."""
        buf = np.zeros(min(max(int(self._steps_taken) + 8, 64), 1 << 20))
        n = C.c_long()
        check(lib().e2d_get_dt_history(self._h, buf.ctypes.data_as(C.POINTER(C.c_double)), buf.size, C.byref(n)),
              "e2d_get_dt_history")
        return buf[: n.value].copy()

    def set_time(self, t: float, nStep: int) -> None:
        check(lib().e2d_set_time(self._h, float(t), int(nStep)), "e2d_set_time")
        self._steps_taken = max(self._steps_taken, int(nStep))

    # -- data movement --------------------------------------------------------------------------
    def download(self, which: int = E2D_U, layout: int = LAYOUT_SOA) -> np.ndarray:
        if layout == LAYOUT_SOA:
            out = np.empty(self.shape)
        else:
            out = np.empty((self.isize, self.jsize_loc, 4))
        check(lib().e2d_download(self._h, int(which), out.ctypes.data, layout), "e2d_download")
        return out

    def upload(self, which: int, host: np.ndarray, layout: int = LAYOUT_SOA) -> None:
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == 4 * self.jsize_loc * self.isize
        check(lib().e2d_upload(self._h, int(which), host.ctypes.data, layout), "e2d_upload")

    def step_host(self, U_in: np.ndarray, U_out: np.ndarray) -> float:
        """One step for a caller whose state lives in host memory (H2D, boundaries, dt, step, D2H)."""
        dt = C.c_double()
        check(lib().e2d_step_host(self._h, U_in.ctypes.data, U_out.ctypes.data, C.byref(dt)), "e2d_step_host")
        return dt.value

    def step_host_ptr(self, in_ptr: int, out_ptr: int) -> float:
        dt = C.c_double()
        check(lib().e2d_step_host(self._h, in_ptr, out_ptr, C.byref(dt)), "e2d_step_host")
        return dt.value

    def step_host_streamed(self, U_in, U_out, dt: float = 0.0, chunk_rows: int = 0):
        """The host-resident step with H2D / fused step / D2H overlapped chunk by chunk (e2d_step_host_streamed).
        `U_in`, `U_out`: numpy arrays or raw host pointers (int).  `dt` > 0: use it (the previous call's dt_next);
        else dt is computed from the input first.  Returns (dt_used, dt_next)."""
        pin = U_in if isinstance(U_in, int) else U_in.ctypes.data
        pout = U_out if isinstance(U_out, int) else U_out.ctypes.data
        used, nxt = C.c_double(), C.c_double()
        check(lib().e2d_step_host_streamed(self._h, pin, pout, float(dt), int(chunk_rows), C.byref(used),
                                           C.byref(nxt)), "e2d_step_host_streamed")
        return used.value, nxt.value

    def march_host(self, buf_a, buf_b, nsteps: int, chunk_rows: int = 0, t0: float = 0.0):
        """nsteps steps of a state that lives in host memory, ping-ponging between two (pinned) host buffers, every step
        streamed through the device and the steps pipelined (e2d_march_host).  buf_a / buf_b: numpy arrays or raw host
        pointers.  Returns (dts, t); the result is in buf_a for an even nsteps, in buf_b for an odd one."""
        pa = buf_a if isinstance(buf_a, int) else buf_a.ctypes.data
        pb = buf_b if isinstance(buf_b, int) else buf_b.ctypes.data
        dts = np.zeros(max(int(nsteps), 1))
        t = C.c_double(t0)
        check(lib().e2d_march_host(self._h, pa, pb, int(nsteps), int(chunk_rows), dts.ctypes.data_as(C.POINTER(C.c_double)),
                                   C.byref(t)), "e2d_march_host")
        return dts[: int(nsteps)], t.value

    def device_ptr(self, which: int) -> int:
        return lib().e2d_device_ptr(self._h, int(which)) or 0

    @property
    def stream(self) -> int:
        return lib().e2d_stream(self._h) or 0

    def synchronize(self) -> None:
        check(lib().e2d_synchronize(self._h), "e2d_synchronize")

    def enable_timers(self, on: bool = True) -> None:
        check(lib().e2d_enable_timers(self._h, int(on)), "e2d_enable_timers")

    def timers(self) -> dict:
        out = (C.c_double * 5)()
        check(lib().e2d_get_timers(self._h, out), "e2d_get_timers")
        return dict(zip(("boundaries", "godunov", "primitive", "fluxes", "update"), list(out)))


def main_loop(ini_path: str, verbose: bool = True, device_loop: bool = False):
    """The program of src/main.cpp:76-204 through the HydroRun mirror. Returns (hydro, nStep, t)."""
    import time

    configMap = ConfigMap(ini_path)          # main.cpp:76-86
    params = HydroParams()
    params.setup(configMap)
    if verbose:
        print(f"Using Euler implementation version {params.implementationVersion}")
        params.print()
    hydro = HydroRun(params, configMap)
    t, nStep = 0.0, 0
    dt = hydro.compute_dt(nStep % 2)          # main.cpp:87
    hydro.make_boundaries(HydroRun.U)         # main.cpp:90-91
    hydro.make_boundaries(HydroRun.U2)
    if verbose:
        print("Start computation....")
    t0 = time.perf_counter()
    if device_loop and not (params.enableOutput and params.nOutput > 0):
        st = hydro.run()
        nStep, t = st.nStep, st.t
    else:
        while t < params.tEnd and nStep < params.nStepmax:  # main.cpp:100
            if verbose and nStep % 10 == 0:
                print("time step=%7d (dt=% 10.8f t=% 10.8f)" % (nStep, dt, t))
            if params.enableOutput and params.nOutput > 0 and nStep % params.nOutput == 0:
                if verbose:
                    print(f"Output results at time t={t:g} step {nStep} dt={dt:g}")
                hydro.saveData(HydroRun.U if nStep % 2 == 0 else HydroRun.U2, nStep, "U")
            dt = hydro.compute_dt(nStep % 2)  # main.cpp:128
            if t + dt > params.tEnd:          # main.cpp:131-134
                dt = params.tEnd - t
            hydro.godunov_unsplit(nStep, dt)  # main.cpp:139
            nStep += 1
            t += dt
        if params.enableOutput and params.nOutput > 0:
            hydro.saveData(HydroRun.U if nStep % 2 == 0 else HydroRun.U2, nStep, "U")
    hydro.synchronize()
    t_tot = time.perf_counter() - t0
    if verbose:
        print("total           time : %5.3f secondes" % t_tot)
        print("Perf                 : %10.2f number of Mcell-updates/s"
              % (1.0 * nStep * params.isize * params.jsize / t_tot * 1e-6))
    return hydro, nStep, t
