"""ctypes wrapper of the CPU oracle (oracle/mc_oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

import microclimate_jl_b200 as mc
from microclimate_jl_b200 import _abi as A
from microclimate_jl_b200 import marshal as M

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmc_oracle.so")


def build(force=False):
    deps = [os.path.join(_HERE, "mc_oracle.cpp"), os.path.join(_HERE, "mc_weather_oracle.cpp"),
            os.path.join(_HERE, "..", "include", "microclimate_b200.h"), os.path.join(_HERE, "..", "include", "mc_math.h")]
    if (not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(d) for d in deps)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B", "libmc_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.mco_create.restype = C.c_void_p
        L.mco_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64]
        L.mco_destroy.argtypes = [C.c_void_p]
        L.mco_set_sites.argtypes = [C.c_void_p, C.POINTER(A.SitesSoA), C.POINTER(A.InitialSoA)]
        L.mco_step_days.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA), C.c_int]
        L.mco_get_counters.argtypes = [C.c_void_p, C.POINTER(A.Counters)]
        L.mco_get_state.argtypes = [C.c_void_p, C.POINTER(A.StateSoA)]
        L.mco_goff_gratch.restype = C.c_double
        L.mco_goff_gratch.argtypes = [C.c_double]
        L.mco_wet_air.argtypes = [C.POINTER(A.FluidConstants)] + [C.c_double] * 3 + [C.POINTER(C.c_double)]
        L.mco_soil_properties.argtypes = [C.POINTER(A.ModelDesc)] + [C.c_double] * 7 + [C.POINTER(C.c_double)]
        L.mco_sky_temperature.restype = C.c_double
        L.mco_sky_temperature.argtypes = [C.POINTER(A.ModelDesc)] + [C.c_double] * 7
        L.mco_surface_profile.argtypes = [C.POINTER(A.ModelDesc)] + [C.c_double] * 6 + [C.POINTER(C.c_double)]
        L.mco_infiltration_steps.restype = C.c_int
        L.mco_infiltration_steps.argtypes = [C.POINTER(A.ModelDesc), C.POINTER(A.SitesSoA)] + [C.c_double] * 4 + [
            C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.mco_fastpow.restype = C.c_double
        L.mco_fastpow.argtypes = [C.c_double, C.c_double]
        L.mco_phase_transition.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 5
        L.mco_populate_weather.argtypes = [C.c_int64, C.c_int32, C.POINTER(A.WeatherDesc), C.POINTER(A.SolarSitesSoA), C.POINTER(A.WeatherInSoA),
                                           C.POINTER(A.WeatherOutSoA)]
        _lib = L
    return _lib


class OracleEngine:
    """Same call shape as microclimate_jl_b200.SoilEngine, backed by the scalar C++ restatement."""

    def __init__(self, desc: A.ModelDesc, n_sites: int):
        self.desc, self.S = desc, n_sites
        self.h = lib().mco_create(C.byref(desc), n_sites)
        if not self.h:
            raise RuntimeError("mco_create failed (unsupported model for the oracle)")

    def close(self):
        if self.h:
            lib().mco_destroy(self.h)
            self.h = None

    __del__ = close

    def set_sites(self, sites: M.SitesBatch):
        cs, ci = sites.c_sites(), sites.c_initial()
        rc = lib().mco_set_sites(self.h, C.byref(cs), C.byref(ci))
        assert rc == 0

    def step_days(self, day0: int, forcing: M.ForcingBatch, out: M.OutputBatch, n_threads: int = 0) -> int:
        cf, co = forcing.c_struct(), out.c_struct()
        rc = lib().mco_step_days(self.h, day0, forcing.ndays, C.byref(cf), C.byref(co), n_threads)
        if rc < 0:
            raise RuntimeError(f"mco_step_days failed rc={rc}")
        return rc

    def counters(self) -> dict:
        c = A.Counters()
        lib().mco_get_counters(self.h, C.byref(c))
        return c.as_dict()

    def get_state(self) -> M.StateBatch:
        st = M.StateBatch(self.desc.n_soil, self.desc.n_snow, self.S)
        cs = st.c_struct()
        lib().mco_get_state(self.h, C.byref(cs))
        return st


def populate_weather(wdesc, win, wout, solar_sites=None):
    """Oracle of populate_weather! + solar geometry (oracle/mc_weather_oracle.cpp); same batches as SoilEngine.populate_weather."""
    ci, co = win.c_struct(), wout.c_struct()
    cs = solar_sites.c_struct() if solar_sites is not None else None
    rc = lib().mco_populate_weather(win.S, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(co))
    if rc:
        raise RuntimeError(f"mco_populate_weather failed (rc={rc}): a required input is missing")
    return wout
