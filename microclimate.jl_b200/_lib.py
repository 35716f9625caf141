"""Loader of the CUDA C-ABI library (csrc/libmicroclimate_b200.so, include/microclimate_b200.h).

The library is built IN-TREE by `build()` (nvcc, sm_100a) so it travels with the repo snapshot.  There is
no CPU implementation behind this loader: if the library cannot be built/loaded, or no CUDA device is
present, the engine raises — it never falls back to anything else.
"""
import ctypes as C
import os
import shutil
import subprocess

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO = os.environ.get("MC_LIB") or os.path.join(CSRC, "libmicroclimate_b200.so")   # MC_LIB: development override (tuning variants)
INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "microclimate_b200.h")
SOURCES = [os.path.join(CSRC, "mc_api.cu"), os.path.join(CSRC, "mc_weather.cu")]      # two translation units, no relocatable device code
DEPS = SOURCES + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [INCLUDE, os.path.join(os.path.dirname(INCLUDE), "mc_math.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "--shared", "-Xcompiler", "-fPIC"]

EXPORTS = ("mc_abi_version", "mc_backend_name", "mc_default_fluid", "mc_default_model", "mc_create", "mc_destroy",
           "mc_last_error", "mc_set_sites", "mc_get_state", "mc_set_state", "mc_step_days", "mc_step_days_device",
           "mc_step_days_async", "mc_wait",
           "mc_multi_create", "mc_multi_destroy", "mc_multi_last_error", "mc_multi_device_count", "mc_multi_handle", "mc_multi_range",
           "mc_multi_set_sites", "mc_multi_step_days", "mc_multi_step_days_async", "mc_multi_wait",
           "mc_synchronize", "mc_get_counters", "mc_reset_counters", "mc_get_status", "mc_measure_fp64_peak", "mc_bench_conductivity", "mc_get_async_stats",
           "mc_default_weather", "mc_validate_weather", "mc_populate_weather", "mc_populate_weather_device", "mc_step_days_minmax", "mc_multi_step_days_minmax",
           "mc_abi_sizeof")


def _nvcc():
    return shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)


def is_stale():
    return (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libmicroclimate_b200.so (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return SO
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libmicroclimate_b200.so (there is no CPU fallback)")
    # several processes may get here at once (one rank per GPU under torchrun): serialise on a lock file, compile to a private name, rename atomically
    import fcntl
    with open(SO + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if force or is_stale():
                tmp = f"{SO}.{os.getpid()}.tmp"
                cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
                subprocess.check_call(cmd, cwd=CSRC)
                os.replace(tmp, SO)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if is_stale():
        if _nvcc() is None and os.path.exists(SO):
            pass            # prebuilt library shipped with the snapshot; no compiler on this box
        else:
            build()
    L = C.CDLL(SO)
    H = C.c_void_p
    L.mc_abi_version.restype = C.c_int
    L.mc_backend_name.restype = C.c_char_p
    L.mc_default_fluid.argtypes = [C.POINTER(A.FluidConstants)]
    L.mc_default_model.argtypes = [C.POINTER(A.ModelDesc)]
    L.mc_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64, C.c_int, C.POINTER(H)]
    L.mc_destroy.argtypes = [H]
    L.mc_destroy.restype = None
    L.mc_last_error.argtypes = [H]
    L.mc_last_error.restype = C.c_char_p
    L.mc_set_sites.argtypes = [H, C.POINTER(A.SitesSoA), C.POINTER(A.InitialSoA)]
    L.mc_get_state.argtypes = [H, C.POINTER(A.StateSoA)]
    L.mc_set_state.argtypes = [H, C.POINTER(A.StateSoA)]
    L.mc_step_days.argtypes = [H, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA)]
    L.mc_step_days_device.argtypes = [H, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA)]
    L.mc_synchronize.argtypes = [H]
    L.mc_step_days_async.argtypes = [H, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA)]
    L.mc_wait.argtypes = [H]
    L.mc_multi_create.argtypes = [C.POINTER(A.ModelDesc), C.c_int64, C.POINTER(C.c_int), C.c_int, C.POINTER(H)]
    L.mc_multi_destroy.argtypes = [H]
    L.mc_multi_destroy.restype = None
    L.mc_multi_last_error.argtypes = [H]
    L.mc_multi_last_error.restype = C.c_char_p
    L.mc_multi_device_count.argtypes = [H]
    L.mc_multi_handle.argtypes = [H, C.c_int]
    L.mc_multi_handle.restype = H
    L.mc_multi_range.argtypes = [H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.mc_multi_set_sites.argtypes = [H, C.POINTER(A.SitesSoA), C.POINTER(A.InitialSoA)]
    L.mc_multi_step_days.argtypes = [H, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA)]
    L.mc_multi_step_days_async.argtypes = [H, C.c_int64, C.c_int32, C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA)]
    L.mc_multi_wait.argtypes = [H]
    L.mc_get_counters.argtypes = [H, C.POINTER(A.Counters)]
    L.mc_reset_counters.argtypes = [H]
    L.mc_get_status.argtypes = [H, C.POINTER(C.c_int32)]
    L.mc_measure_fp64_peak.argtypes = [H, C.POINTER(C.c_double)]
    L.mc_bench_conductivity.argtypes = [H, C.c_int64, C.c_int, C.POINTER(C.c_double)]
    L.mc_get_async_stats.argtypes = [H, C.POINTER(C.c_ulonglong), C.c_int]
    L.mc_default_weather.argtypes = [C.POINTER(A.WeatherDesc)]
    L.mc_default_weather.restype = None
    L.mc_validate_weather.argtypes = [C.POINTER(A.WeatherDesc), C.c_char_p, C.c_int]
    _wx = [H, C.c_int32, C.POINTER(A.WeatherDesc), C.POINTER(A.SolarSitesSoA), C.POINTER(A.WeatherInSoA), C.POINTER(A.WeatherOutSoA)]
    L.mc_populate_weather.argtypes = _wx
    L.mc_populate_weather_device.argtypes = _wx
    L.mc_step_days_minmax.argtypes = [H, C.c_int64, C.c_int32, C.POINTER(A.WeatherDesc), C.POINTER(A.SolarSitesSoA), C.POINTER(A.WeatherInSoA),
                                      C.POINTER(A.ForcingSoA), C.POINTER(A.OutputSoA), C.POINTER(A.WeatherOutSoA)]
    L.mc_multi_step_days_minmax.argtypes = L.mc_step_days_minmax.argtypes
    L.mc_abi_sizeof.argtypes = [C.c_int]
    L.mc_abi_sizeof.restype = C.c_int64
    if L.mc_abi_version() != A.MC_ABI_VERSION:
        raise RuntimeError(f"{SO}: ABI version {L.mc_abi_version()}, this package expects {A.MC_ABI_VERSION} - rebuild the library (stale build?)")
    # a stale or foreign library whose struct layout differs from the ctypes mirrors would corrupt memory silently: refuse it
    for k, cls in enumerate(A.ABI_SIZEOF):
        if L.mc_abi_sizeof(k) != C.sizeof(cls):
            raise RuntimeError(f"{SO}: sizeof({cls.__name__}) is {L.mc_abi_sizeof(k)} in the library, {C.sizeof(cls)} in _abi.py - rebuild the library (stale build?)")
    _lib = L
    return L
