"""ctypes wrapper of tests/hostemu/hostemu.cpp — the product's per-site kernel body (mc_device.cuh) compiled
for the CPU.  TEST HARNESS ONLY: lets the CPU-only tier check the logic the CUDA kernel executes against the
oracle.  The product never loads it (libmicroclimate_b200.so has no CPU path)."""
import ctypes as C
import os
import subprocess

from microclimate_jl_b200 import _abi as A
from microclimate_jl_b200 import marshal as M

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
_SO = os.path.join(_HERE, "libmc_hostemu.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "microclimate.jl_b200", "csrc")
_DEPS = [os.path.join(_HERE, "hostemu.cpp"), os.path.join(_ROOT, "include", "mc_math.h"), os.path.join(_ROOT, "include", "microclimate_b200.h")] + \
        sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh"))


def build(force=False, so=_SO, defs=()):
    if not force and os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(d) for d in _DEPS):
        return so
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-mfma", *defs, "-o", so, os.path.join(_HERE, "hostemu.cpp")])
    return so


_lib = None
_variants = {}


def _bind(L):
    L.he_create.restype = C.c_void_p
    L.he_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64]
    L.he_destroy.argtypes = [C.c_void_p]
    L.he_set_sites.argtypes = [C.c_void_p, C.POINTER(A.SitesSoA), C.POINTER(A.InitialSoA)]
    L.he_step_days.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA)]
    L.he_get_counters.argtypes = [C.c_void_p, C.POINTER(A.Counters)]
    L.he_get_state.argtypes = [C.c_void_p, C.POINTER(A.StateSoA)]
    L.he_set_state.argtypes = [C.c_void_p, C.POINTER(A.StateSoA)]
    return L


def variant(name, defs):
    """The same harness compiled with extra -D switches (compile-time kernel options such as MC_COOP_MOIST=1)."""
    if name not in _variants:
        so = os.path.join(_HERE, f"libmc_hostemu_{name}.so")
        _variants[name] = _bind(C.CDLL(build(so=so, defs=tuple(defs))))
    return _variants[name]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.he_create.restype = C.c_void_p
        L.he_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64]
        L.he_destroy.argtypes = [C.c_void_p]
        L.he_set_sites.argtypes = [C.c_void_p, C.POINTER(A.SitesSoA), C.POINTER(A.InitialSoA)]
        L.he_step_days.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA)]
        L.he_get_counters.argtypes = [C.c_void_p, C.POINTER(A.Counters)]
        L.he_get_state.argtypes = [C.c_void_p, C.POINTER(A.StateSoA)]
        L.he_set_state.argtypes = [C.c_void_p, C.POINTER(A.StateSoA)]
        _lib = L
    return _lib


class HostEmuEngine:
    LIB = None          # subclasses bind a variant() library here

    def __init__(self, desc, n_sites):
        self.desc, self.S = desc, n_sites
        self._L = self.LIB if self.LIB is not None else lib()
        self.h = self._L.he_create(C.byref(desc), n_sites)
        if not self.h:
            raise RuntimeError("he_create failed")

    def close(self):
        if self.h:
            self._L.he_destroy(self.h)
            self.h = None

    __del__ = close

    def set_sites(self, sites):
        cs, ci = sites.c_sites(), sites.c_initial()
        assert self._L.he_set_sites(self.h, C.byref(cs), C.byref(ci)) == 0

    def step_days(self, day0, forcing, out):
        cf, co = forcing.c_struct(), out.c_struct()
        rc = self._L.he_step_days(self.h, day0, forcing.ndays, C.byref(cf), C.byref(co))
        if rc < 0:
            raise RuntimeError(f"he_step_days rc={rc}")
        return rc

    def counters(self):
        c = A.Counters()
        self._L.he_get_counters(self.h, C.byref(c))
        return c.as_dict()

    def get_state(self):
        st = M.StateBatch(self.desc.n_soil, self.desc.n_snow, self.S)
        cs = st.c_struct()
        self._L.he_get_state(self.h, C.byref(cs))
        return st

    def set_state(self, st):
        cs = st.c_struct()
        assert self._L.he_set_state(self.h, C.byref(cs)) == 0


# ---- weather device code (csrc/mc_weather.cuh) on the CPU ----
_WSO = os.path.join(_HERE, "libmc_weather_hostemu.so")
_wlib = None


def weather_lib():
    global _wlib
    if _wlib is None:
        src = os.path.join(_HERE, "weather_hostemu.cpp")
        deps = [src, os.path.join(_CSRC, "mc_weather.cuh"), os.path.join(_ROOT, "include", "mc_math.h"), os.path.join(_ROOT, "include", "microclimate_b200.h")]
        if not os.path.exists(_WSO) or os.path.getmtime(_WSO) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-mfma", "-o", _WSO, src])
        L = C.CDLL(_WSO)
        L.he_populate_weather.argtypes = [C.c_int64, C.c_int32, C.POINTER(A.WeatherDesc), C.POINTER(A.SolarSitesSoA), C.POINTER(A.WeatherInSoA),
                                          C.POINTER(A.WeatherOutSoA)]
        _wlib = L
    return _wlib


def populate_weather(wdesc, win, wout, solar_sites=None):
    """weather_kernel's body (mcw::site_day) run over the (site, day) grid on the CPU; same batches as SoilEngine.populate_weather."""
    ci, co = win.c_struct(), wout.c_struct()
    cs = solar_sites.c_struct() if solar_sites is not None else None
    weather_lib().he_populate_weather(win.S, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(co))
    return wout


# ---- the weather translation unit (csrc/mc_weather.cu: entry points + staging plans + kernel) over a stand-in CUDA runtime ----
_WASO = os.path.join(_HERE, "libmc_weather_api_hostemu.so")
_walib = None


def weather_api_lib():
    global _walib
    if _walib is None:
        srcs = [os.path.join(_HERE, "weather_api_hostemu.cpp"), os.path.join(_HERE, "hostemu.cpp")]
        deps = srcs + _DEPS + [os.path.join(_CSRC, "mc_weather.cu"), os.path.join(_CSRC, "mc_weather.cuh"), os.path.join(_CSRC, "mc_internal.h")]
        if not os.path.exists(_WASO) or os.path.getmtime(_WASO) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-mfma", "-o", _WASO, *srcs])
        L = C.CDLL(_WASO)
        H = C.c_void_p
        L.wa_create.restype = H
        L.wa_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64]
        L.wa_destroy.argtypes = [H]
        L.wa_set_sites.argtypes = [H, C.POINTER(A.SitesSoA), C.POINTER(A.InitialSoA)]
        L.wa_last_error.argtypes = [H]
        L.wa_last_error.restype = C.c_char_p
        L.wa_live_allocations.restype = C.c_long
        wx = [H, C.c_int32, C.POINTER(A.WeatherDesc), C.POINTER(A.SolarSitesSoA), C.POINTER(A.WeatherInSoA), C.POINTER(A.WeatherOutSoA)]
        L.mc_populate_weather.argtypes = wx
        L.mc_populate_weather_device.argtypes = wx
        L.mc_step_days_minmax.argtypes = [H, C.c_int64, C.c_int32, C.POINTER(A.WeatherDesc), C.POINTER(A.SolarSitesSoA), C.POINTER(A.WeatherInSoA),
                                          C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA), C.POINTER(A.WeatherOutSoA)]
        L.mc_multi_step_days_minmax.argtypes = L.mc_step_days_minmax.argtypes
        L.wa_multi_create.restype = H
        L.wa_multi_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64, C.c_int]
        L.wa_multi_destroy.argtypes = [H]
        L.wa_multi_last_error.argtypes = [H]
        L.wa_multi_last_error.restype = C.c_char_p
        L.mc_multi_handle.argtypes = [H, C.c_int]
        L.mc_multi_handle.restype = H
        L.mc_multi_range.argtypes = [H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _walib = L
    return _walib


class WeatherApiEmu:
    """Same weather methods as microclimate_jl_b200.SoilEngine, served by mc_weather.cu compiled for the host (see weather_api_hostemu.cpp)."""

    def __init__(self, desc, n_sites):
        self.desc, self.S = desc, n_sites
        self._L = weather_api_lib()
        self.h = self._L.wa_create(C.byref(desc), n_sites)
        if not self.h:
            raise RuntimeError("wa_create failed")

    def close(self):
        if self.h:
            self._L.wa_destroy(self.h)
            self.h = None

    def _check(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"{what} failed (rc={rc}): {self._L.wa_last_error(self.h).decode()}")
        return rc

    def set_sites(self, sites):
        cs, ci = sites.c_sites(), sites.c_initial()
        self._check(self._L.wa_set_sites(self.h, C.byref(cs), C.byref(ci)), "he_set_sites")

    def populate_weather(self, wdesc, win, wout, solar_sites=None, device_entry=False):
        ci, co = win.c_struct(), wout.c_struct()
        cs = solar_sites.c_struct() if solar_sites is not None else None
        fn = self._L.mc_populate_weather_device if device_entry else self._L.mc_populate_weather
        self._L.wa_allow_foreign_pointers(1 if device_entry else 0)      # the *_device entry is handed caller-owned "device" memory (host arrays here)
        try:
            self._check(fn(self.h, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(co)), "mc_populate_weather")
        finally:
            self._L.wa_allow_foreign_pointers(0)
        return wout

    def step_days_minmax(self, day0, wdesc, win, rest, out, solar_sites=None, weather_out=None, use_rest_zenith=False):
        ci, cf, co = win.c_struct(), rest.c_struct(), out.c_struct()
        if not use_rest_zenith:
            cf.zenith_angle, cf.zenith_slope_angle = None, None
        cs = solar_sites.c_struct() if solar_sites is not None else None
        cw = weather_out.c_struct() if weather_out is not None else None
        return self._check(self._L.mc_step_days_minmax(self.h, day0, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(cf),
                                                       C.byref(co), C.byref(cw) if cw is not None else None), "mc_step_days_minmax")


class WeatherApiMultiEmu:
    """MultiSoilEngine's min/max step on the emulation: `n_devices` emulated devices over contiguous column ranges of one set of host arrays."""

    def __init__(self, desc, n_sites, n_devices):
        self.desc, self.S, self.G = desc, n_sites, n_devices
        self._L = weather_api_lib()
        self.h = self._L.wa_multi_create(C.byref(desc), n_sites, n_devices)

    def close(self):
        if self.h:
            self._L.wa_multi_destroy(self.h)
            self.h = None

    def ranges(self):
        out = []
        for g in range(self.G):
            lo, hi = C.c_int64(), C.c_int64()
            self._L.mc_multi_range(self.h, g, C.byref(lo), C.byref(hi))
            out.append((lo.value, hi.value))
        return out

    def set_sites(self, sites):
        for g, (lo, hi) in enumerate(self.ranges()):
            if hi > lo:
                part = sites.slice(lo, hi)
                cs, ci = part.c_sites(), part.c_initial()
                assert self._L.wa_set_sites(self._L.mc_multi_handle(self.h, g), C.byref(cs), C.byref(ci)) == 0

    def step_days_minmax(self, day0, wdesc, win, rest, out, solar_sites=None, weather_out=None, use_rest_zenith=False):
        ci, cf, co = win.c_struct(), rest.c_struct(), out.c_struct()
        if not use_rest_zenith:
            cf.zenith_angle, cf.zenith_slope_angle = None, None
        cs = solar_sites.c_struct() if solar_sites is not None else None
        cw = weather_out.c_struct() if weather_out is not None else None
        rc = self._L.mc_multi_step_days_minmax(self.h, day0, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(cf),
                                               C.byref(co), C.byref(cw) if cw is not None else None)
        if rc < 0:
            raise RuntimeError(f"mc_multi_step_days_minmax failed (rc={rc}): {self._L.wa_multi_last_error(self.h).decode()}")
        return rc
