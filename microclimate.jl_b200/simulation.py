"""Batched-sites soil engine — the host-side mirror of the reference's `solve_soil!` entry
(src/simulation.jl:355-819) and of its many-sites idiom `reinit!(cache, inputs); solve!(cache)`
(src/simulation.jl:266-277): ONE MicroModel, MANY MicroInputs.

Everything here is marshalling and orchestration; all arithmetic runs in the CUDA library behind the
C ABI (include/microclimate_b200.h).  No CUDA device -> RuntimeError (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi as A
from . import _lib
from . import marshal as M
from . import types as T


class SoilEngine:
    """One handle = one GPU, S sites resident in HBM (mc_create / mc_set_sites / mc_step_days / mc_destroy)."""

    def __init__(self, desc: A.ModelDesc, n_sites: int, device: int = 0):
        self.desc, self.S, self.device = desc, int(n_sites), int(device)
        self._L = _lib.lib()
        self.h = C.c_void_p()
        rc = self._L.mc_create(C.byref(desc), self.S, self.device, C.byref(self.h))
        if rc != 0:
            msg = self._L.mc_last_error(None)
            self.h = None
            raise RuntimeError(f"mc_create failed (rc={rc}): {msg.decode() if msg else ''}")

    # -- lifetime --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self._L.mc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            msg = self._L.mc_last_error(self.h)
            raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
        return rc

    # -- reference-shaped operations ----------------------------------------------------------------
    def set_sites(self, sites: M.SitesBatch):
        """reinit!(cache, inputs): new sites, fresh state (src/simulation.jl:274-277)."""
        if sites.S != self.S or sites.N != self.desc.n_soil:
            raise ValueError(f"SitesBatch is {sites.N} nodes x {sites.S} sites; engine expects {self.desc.n_soil} x {self.S}")
        cs, ci = sites.c_sites(), sites.c_initial()
        self._check(self._L.mc_set_sites(self.h, C.byref(cs), C.byref(ci)), "mc_set_sites")

    def step_days(self, day0: int, forcing: M.ForcingBatch, out: M.OutputBatch) -> int:
        """solve_soil! for days [day0, day0+forcing.ndays): host buffers in/out, H2D + kernel + D2H.
        Returns the number of sites whose status word is non-zero."""
        if forcing.S != self.S or out.S != self.S or out.ndays != forcing.ndays:
            raise ValueError("forcing/output batch shapes do not match the engine")
        cf, co = forcing.c_struct(), out.c_struct()
        return self._check(self._L.mc_step_days(self.h, day0, forcing.ndays, C.byref(cf), C.byref(co)), "mc_step_days")

    def step_days_async(self, day0: int, forcing: M.ForcingBatch, out: M.OutputBatch):
        """mc_step_days_async: enqueue H2D -> kernel -> D2H of this chunk and return; call wait() before reading `out`.  Issuing the next chunk
        before wait() overlaps its H2D (and the previous chunk's D2H) with this chunk's kernel.  The batches must stay alive until wait()."""
        if forcing.S != self.S or out.S != self.S or out.ndays != forcing.ndays:
            raise ValueError("forcing/output batch shapes do not match the engine")
        cf, co = forcing.c_struct(), out.c_struct()
        self._keep = getattr(self, "_keep", [])[-3:] + [(forcing, out, cf, co)]
        self._check(self._L.mc_step_days_async(self.h, day0, forcing.ndays, C.byref(cf), C.byref(co)), "mc_step_days_async")

    def wait(self) -> int:
        """mc_wait: all enqueued chunks complete; returns the flagged-site count of the most recent one."""
        return self._check(self._L.mc_wait(self.h), "mc_wait")

    # -- widening: populate_weather! + solar geometry on the device (weather.py, csrc/mc_weather.cuh) -----------------
    def populate_weather(self, wdesc: A.WeatherDesc, win, wout, solar_sites=None):
        """solve_solar! (geometry) + populate_weather! (src/simulation.jl:20-36,291-352) for this engine's sites: host arrays in, host arrays out."""
        if win.S != self.S or wout.S != self.S or wout.ndays != win.ndays or (solar_sites is not None and solar_sites.S != self.S):
            raise ValueError("weather batch shapes do not match the engine")
        ci, co = win.c_struct(), wout.c_struct()
        cs = solar_sites.c_struct() if solar_sites is not None else None
        self._check(self._L.mc_populate_weather(self.h, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(co)),
                    "mc_populate_weather")
        return wout

    def step_days_minmax(self, day0: int, wdesc: A.WeatherDesc, win, rest: M.ForcingBatch, out: M.OutputBatch, solar_sites=None, weather_out=None,
                         use_rest_zenith: bool = False) -> int:
        """solve!(cache) of a min/max-driven problem (src/simulation.jl:246-260): the hourly weather is expanded on the device straight into the soil kernel's
        forcing arrays, then solve_soil! runs for days [day0, day0 + ndays).  `rest` carries what populate_weather! does not produce (daily members, rainfall,
        longwave, prescribed moisture); with `use_rest_zenith` its zenith / slope-zenith angles (SolarRadiation.jl's own) are used, otherwise the built-in
        geometry computes them from `solar_sites`."""
        if win.S != self.S or out.S != self.S or rest.S != self.S or out.ndays != win.ndays or rest.ndays != win.ndays:
            raise ValueError("weather/forcing/output batch shapes do not match the engine")
        ci, cf, co = win.c_struct(), rest.c_struct(), out.c_struct()
        if not use_rest_zenith:
            cf.zenith_angle, cf.zenith_slope_angle = None, None
        cs = solar_sites.c_struct() if solar_sites is not None else None
        cw = weather_out.c_struct() if weather_out is not None else None
        return self._check(self._L.mc_step_days_minmax(self.h, day0, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(cf),
                                                       C.byref(co), C.byref(cw) if cw is not None else None), "mc_step_days_minmax")

    def step_days_device(self, day0: int, ndays: int, forcing_ptrs: dict, output_ptrs: dict):
        """Same with DEVICE pointers (ints) already resident on this GPU; asynchronous until synchronize()."""
        cf, co = A.ForcingSoA(), A.OutputSoA()
        pd = C.POINTER(C.c_double)
        for k, v in forcing_ptrs.items():
            setattr(cf, k, C.cast(C.c_void_p(int(v)), pd))
        for k, v in output_ptrs.items():
            setattr(co, k, C.cast(C.c_void_p(int(v)), pd))
        self._check(self._L.mc_step_days_device(self.h, day0, ndays, C.byref(cf), C.byref(co)), "mc_step_days_device")

    def synchronize(self):
        self._check(self._L.mc_synchronize(self.h), "mc_synchronize")

    def counters(self) -> dict:
        c = A.Counters()
        self._check(self._L.mc_get_counters(self.h, C.byref(c)), "mc_get_counters")
        return c.as_dict()

    def reset_counters(self):
        self._check(self._L.mc_reset_counters(self.h), "mc_reset_counters")

    def status(self) -> np.ndarray:
        st = np.zeros(self.S, dtype=np.int32)
        self._check(self._L.mc_get_status(self.h, st.ctypes.data_as(C.POINTER(C.c_int32))), "mc_get_status")
        return st

    def get_state(self) -> M.StateBatch:
        st = M.StateBatch(self.desc.n_soil, self.desc.n_snow, self.S)
        cs = st.c_struct()
        self._check(self._L.mc_get_state(self.h, C.byref(cs)), "mc_get_state")
        return st

    def set_state(self, st: M.StateBatch):
        cs = st.c_struct()
        self._check(self._L.mc_set_state(self.h, C.byref(cs)), "mc_set_state")

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self._check(self._L.mc_measure_fp64_peak(self.h, C.byref(v)), "mc_measure_fp64_peak")
        return v.value


class _HandleView(SoilEngine):
    """A per-device handle owned by a MultiSoilEngine (state, counters, status of that device's site range)."""

    def __init__(self, lib, h, desc, n_sites, device):
        self._L, self.h, self.desc, self.S, self.device = lib, C.c_void_p(h), desc, int(n_sites), int(device)

    def close(self):
        self.h = None


class MultiSoilEngine:
    """One host process, several GPUs (mc_multi_*): the sites of ONE set of host arrays are sharded contiguously over `devices`; every device
    runs the asynchronous H2D / kernel / D2H pipeline on its column range and writes into its columns of the caller's output arrays."""

    def __init__(self, desc: A.ModelDesc, n_sites: int, devices: Sequence[int]):
        self.desc, self.S, self.devices = desc, int(n_sites), [int(d) for d in devices]
        self._L = _lib.lib()
        self.h = C.c_void_p()
        ids = (C.c_int * len(self.devices))(*self.devices)
        rc = self._L.mc_multi_create(C.byref(desc), self.S, ids, len(self.devices), C.byref(self.h))
        if rc != 0:
            msg = self._L.mc_multi_last_error(None)
            self.h = None
            raise RuntimeError(f"mc_multi_create failed (rc={rc}): {msg.decode() if msg else ''}")
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self._L.mc_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            msg = self._L.mc_multi_last_error(self.h)
            raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
        return rc

    def ranges(self):
        out = []
        for g in range(len(self.devices)):
            lo, hi = C.c_int64(), C.c_int64()
            self._check(self._L.mc_multi_range(self.h, g, C.byref(lo), C.byref(hi)), "mc_multi_range")
            out.append((lo.value, hi.value))
        return out

    def device_engine(self, g: int) -> SoilEngine:
        lo, hi = self.ranges()[g]
        return _HandleView(self._L, self._L.mc_multi_handle(self.h, g), self.desc, hi - lo, self.devices[g])

    def set_sites(self, sites: M.SitesBatch):
        if sites.S != self.S or sites.N != self.desc.n_soil:
            raise ValueError("SitesBatch shape does not match the engine")
        cs, ci = sites.c_sites(), sites.c_initial()
        self._check(self._L.mc_multi_set_sites(self.h, C.byref(cs), C.byref(ci)), "mc_multi_set_sites")

    def step_days_async(self, day0: int, forcing: M.ForcingBatch, out: M.OutputBatch):
        if forcing.S != self.S or out.S != self.S or out.ndays != forcing.ndays:
            raise ValueError("forcing/output batch shapes do not match the engine")
        cf, co = forcing.c_struct(), out.c_struct()
        self._keep = self._keep[-3:] + [(forcing, out, cf, co)]
        self._check(self._L.mc_multi_step_days_async(self.h, day0, forcing.ndays, C.byref(cf), C.byref(co)), "mc_multi_step_days_async")

    def wait(self) -> int:
        return self._check(self._L.mc_multi_wait(self.h), "mc_multi_wait")

    def step_days(self, day0: int, forcing: M.ForcingBatch, out: M.OutputBatch) -> int:
        self.step_days_async(day0, forcing, out)
        return self.wait()

    def step_days_minmax(self, day0: int, wdesc: A.WeatherDesc, win, rest: M.ForcingBatch, out: M.OutputBatch, solar_sites=None, weather_out=None,
                         use_rest_zenith: bool = False) -> int:
        """SoilEngine.step_days_minmax over all devices of this engine (mc_multi_step_days_minmax): every GPU expands the weather of, and integrates, its column range."""
        if win.S != self.S or out.S != self.S or rest.S != self.S or out.ndays != win.ndays or rest.ndays != win.ndays:
            raise ValueError("weather/forcing/output batch shapes do not match the engine")
        ci, cf, co = win.c_struct(), rest.c_struct(), out.c_struct()
        if not use_rest_zenith:
            cf.zenith_angle, cf.zenith_slope_angle = None, None
        cs = solar_sites.c_struct() if solar_sites is not None else None
        cw = weather_out.c_struct() if weather_out is not None else None
        return self._check(self._L.mc_multi_step_days_minmax(self.h, day0, win.ndays, C.byref(wdesc), C.byref(cs) if cs is not None else None, C.byref(ci), C.byref(cf),
                                                             C.byref(co), C.byref(cw) if cw is not None else None), "mc_multi_step_days_minmax")

    def counters(self) -> dict:
        tot = None
        for g in range(len(self.devices)):
            c = self.device_engine(g).counters()
            tot = c if tot is None else {k: (max(tot[k], v) if k == "gpu_ms" else tot[k] + v) for k, v in c.items()}
        return tot

    def status(self) -> np.ndarray:
        return np.concatenate([self.device_engine(g).status() for g in range(len(self.devices))])


def site_shards(n_sites: int, n_shards: int):
    """Contiguous, near-equal site ranges [lo, hi) — the only partitioning the path needs (no collective)."""
    base, rem = divmod(n_sites, n_shards)
    out, lo = [], 0
    for g in range(n_shards):
        hi = lo + base + (1 if g < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def solve_soil_batched(model: T.MicroModel, inputs: Sequence[T.MicroInputs], *, days: Sequence[int],
                       time_mode=None, fields: Sequence[str] = M.DEFAULT_OUTPUT_FIELDS,
                       chunk_days: Optional[int] = None, device: int = 0, devices: Optional[Sequence[int]] = None) -> M.OutputBatch:
    """Batched `solve_soil!` (+ `solve_air!` when any of the profile fields air_temperature / wind_speed / relative_humidity /
    convective_heat_flux / friction_velocity is selected): one model, many inputs -> one OutputBatch (site-fastest);
    `site_result(s)` un-marshals a site into the reference's (nsteps, N) MicroResult layout.

    `chunk_days` streams time in chunks with the carried state resident on the GPU (results are identical
    to a single chunk: the state carried between chunks is exactly what the reference carries between days — with one exception, the first `soil_humidity`
    row of every chunk after the first, see DESIGN.md "known limitations").  `devices` shards the sites
    over several GPUs from this one process (MultiSoilEngine); chunks are pipelined (H2D / kernel / D2H overlap) either way."""
    time_mode = time_mode if time_mode is not None else T.NonConsecutiveDayMode()
    if len(inputs) == 0:
        raise ValueError("inputs must be non-empty")
    if isinstance(model.snow_model, T.SnowModel):
        # the initial snowpack is model-level in the ABI (mc_snow_desc.initial_snow_*): a batch whose sites differ would silently run on site 0's values
        key = lambda i: (i.initial_snow_depth, i.initial_snow_temperature, i.initial_snow_density)      # noqa: E731
        if any(key(i) != key(inputs[0]) for i in inputs[1:]):
            raise ValueError("initial_snow_depth / _temperature / _density must be identical for all sites of a batch (model-level in the C ABI)")
    for name in ("reference_temperature", "reference_wind_speed", "reference_humidity", "cloud_cover", "global_radiation", "zenith_angle"):
        if any(getattr(i.environment_hourly, name) is None for i in inputs):
            raise ValueError(f"environment_hourly.{name} is missing: solve_soil_batched needs the hourly weather populate_weather! leaves in cache.output "
                             "(for min/max-driven inputs use solve_minmax_batched)")
    desc = M.model_desc(model, time_mode, inputs0=inputs[0])
    sites = M.SitesBatch.from_inputs(model, inputs)
    forcing = M.ForcingBatch.from_inputs(model, inputs, days)
    ndays, S = len(days), len(inputs)
    out = M.OutputBatch(ndays, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
    with (MultiSoilEngine(desc, S, devices) if devices else SoilEngine(desc, S, device)) as eng:
        eng.set_sites(sites)
        if not chunk_days or chunk_days >= ndays:
            eng.step_days(0, forcing, out)
        else:
            d0, parts = 0, []
            while d0 < ndays:                     # enqueue chunk after chunk: copies of neighbouring chunks overlap the kernels
                d1 = min(ndays, d0 + chunk_days)
                part = M.OutputBatch(d1 - d0, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
                fpart = forcing.day_slice(d0, d1)
                eng.step_days_async(d0, fpart, part)
                parts.append((d0, part, fpart))
                if len(parts) > 2:                # chunks older than the pipeline depth are complete
                    r0, p0, _ = parts.pop(0)
                    out.assign_rows(24 * r0, p0)
                d0 = d1
            eng.wait()
            for r0, p0, _ in parts:
                out.assign_rows(24 * r0, p0)
            out.status[:] = eng.status()
    return out


def solve_minmax_batched(model: T.MicroModel, inputs: Sequence[T.MicroInputs], *, days: Sequence[int], clear_sky_direct, clear_sky_diffuse,
                         time_mode=None, shortwave=None, geometry=None, fields: Sequence[str] = M.DEFAULT_OUTPUT_FIELDS, device: int = 0,
                         chunk_days: Optional[int] = None, zenith_angle=None, zenith_slope_angle=None, return_weather: bool = False, engine_factory=None,
                         devices: Optional[Sequence[int]] = None):
    """Batched `solve!(cache)` for min/max-driven problems (src/simulation.jl:246-260): every `inputs[s].environment_minmax` (MonthlyMinMaxEnvironment /
    DailyMinMaxEnvironment, all with the same forcing model) is expanded to hourly weather ON THE DEVICE (populate_weather!), the soil path integrates it
    (solve_soil!, plus solve_air! when profile fields are selected) and only the selected outputs come back (mc_step_days_minmax).

    `clear_sky_direct` / `clear_sky_diffuse` are SolarRadiation.jl's hourly clear-sky irradiances on the horizontal, [24*ndays][S] (its spectral model is not part of
    this library).  Zenith and slope-zenith angles come from the built-in geometry (site latitude / slope / aspect) unless given.  `environment_hourly` supplies
    `pressure` (scalar or hourly), rainfall / longwave when the model uses them, and an optional `global_radiation` override.  `devices` shards the sites over
    several GPUs from this one process.  Returns the OutputBatch, and with `return_weather` also the expanded hourly weather (WeatherOutBatch)."""
    from . import weather as W
    time_mode = time_mode if time_mode is not None else T.NonConsecutiveDayMode()
    if len(inputs) == 0 or any(i.environment_minmax is None for i in inputs):
        raise ValueError("every input needs environment_minmax (use solve_soil_batched for hourly-driven problems)")
    ndays, S = len(days), len(inputs)
    H = 24 * ndays
    env0 = inputs[0].environment_minmax
    curves = {q: W.forcing_of(env0.forcings, q).curve for q in A.WEATHER_QUANTITIES}
    for i in inputs[1:]:
        fi = i.environment_minmax.forcings
        if W.humidity_forcing(fi)[1] != W.humidity_forcing(env0.forcings)[1] or any(
                W.forcing_of(fi, q).curve is not curves[q] and (W.forcing_of(fi, q).curve.shapes, W.forcing_of(fi, q).curve.inputs) != (curves[q].shapes, curves[q].inputs)
                for q in A.WEATHER_QUANTITIES):
            raise ValueError("all sites of a batch must share one forcing model (the DielCurves are model-level in the C ABI)")
    wdesc = W.weather_desc(None, shortwave or getattr(model.radiation, "shortwave_model", None) or W.AngstromMaxwellShortwave(), geometry or W.SolarGeometry(),
                           forcings=env0.forcings)
    desc = M.model_desc(model, time_mode, inputs0=inputs[0])
    sites = M.SitesBatch.from_inputs(model, inputs)
    rest = M.ForcingBatch.from_inputs(model, inputs, days)
    h0 = inputs[0].environment_hourly
    gr = np.ascontiguousarray(rest.global_radiation) if h0.global_radiation is not None else None
    win = W.WeatherInBatch.from_environments(np.asarray(days, dtype=np.int32), [i.environment_minmax for i in inputs],
                                             direct_horizontal=clear_sky_direct, diffuse_horizontal=clear_sky_diffuse,
                                             global_radiation_hourly=gr, pressure_hourly=rest.pressure)
    solar = W.SolarSitesBatch([i.site.latitude for i in inputs], [i.site.slope for i in inputs], [i.site.aspect for i in inputs])
    own_zenith = zenith_angle is not None and zenith_slope_angle is not None
    if own_zenith:
        rest.zenith_angle[...] = np.minimum(zenith_angle, 90.0)
        rest.zenith_slope_angle[...] = np.minimum(zenith_slope_angle, 90.0)
    out = M.OutputBatch(ndays, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
    wx = W.WeatherOutBatch(ndays, S) if return_weather else None
    eng = (engine_factory or (lambda d, n: MultiSoilEngine(d, n, devices) if devices else SoilEngine(d, n, device)))(desc, S)
    try:
        eng.set_sites(sites)
        cd = ndays if not chunk_days else min(int(chunk_days), ndays)
        d0 = 0
        while d0 < ndays:
            d1 = min(ndays, d0 + cd)
            part = out if (d0 == 0 and d1 == ndays) else M.OutputBatch(d1 - d0, desc.n_soil, S, fields, desc.n_snow, desc.n_heights)
            wpart = wx if (d0 == 0 and d1 == ndays) or wx is None else W.WeatherOutBatch(d1 - d0, S)
            eng.step_days_minmax(d0, wdesc, win.day_slice(d0, d1), rest.day_slice(d0, d1), part, solar, wpart, use_rest_zenith=own_zenith)
            if part is not out:
                out.assign_rows(24 * d0, part)
                if wx is not None:
                    for k in A.WEATHER_OUT_HOURLY:
                        getattr(wx, k)[24 * d0:24 * d1] = getattr(wpart, k)
                    for k in A.WEATHER_OUT_DAILY:
                        getattr(wx, k)[d0:d1] = getattr(wpart, k)
            d0 = d1
    finally:
        eng.close()
    return (out, wx) if return_weather else out


def solve_soil(problem: T.MicroProblem, **kw) -> T.MicroResult:
    """Single-site convenience mirroring `solve(problem)` for the soil path."""
    out = solve_soil_batched(problem.model, [problem.inputs], days=problem.days, time_mode=problem.time_mode, **kw)
    return out.site_result(0)


def solve(problem: T.MicroProblem, *, clear_sky_direct=None, clear_sky_diffuse=None, **kw) -> T.MicroResult:
    """`solve(problem)` (src/simulation.jl:262-264) for one site: populate_weather! (when the inputs carry a min/max environment) + solve_soil! (+ solve_air! when profile
    fields are selected).  The spectral clear-sky model of SolarRadiation.jl is not part of this library: a min/max-driven problem needs its hourly clear-sky
    `direct_horizontal` / `diffuse_horizontal` ([24*ndays]) here."""
    if problem.inputs.environment_minmax is None:
        return solve_soil(problem, **kw)
    if clear_sky_direct is None or clear_sky_diffuse is None:
        raise ValueError("a min/max-driven problem needs clear_sky_direct / clear_sky_diffuse (SolarRadiation.jl's hourly clear-sky irradiance on the horizontal)")
    H = 24 * len(problem.days)
    out = solve_minmax_batched(problem.model, [problem.inputs], days=problem.days, time_mode=problem.time_mode,
                               clear_sky_direct=np.asarray(clear_sky_direct, float).reshape(H, 1), clear_sky_diffuse=np.asarray(clear_sky_diffuse, float).reshape(H, 1), **kw)
    return out.site_result(0)
