"""Marshalling: reference-shaped objects  ->  unit-stripped, site-fastest SoA buffers of the C ABI.

This is what the Julia wrapper does with `ustrip` + transposes before the `ccall`
(julia/MicroclimateB200.jl); the Python mirror is what the tests and bench drive.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi as A
from . import types as T


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return C.POINTER(C.c_double)()
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "SoA buffers must be contiguous float64"
    return a.ctypes.data_as(C.POINTER(C.c_double))


def model_desc(model: T.MicroModel, time_mode, *, inputs0: Optional[T.MicroInputs] = None) -> A.ModelDesc:
    """Flatten MicroModel + MicroConfig + time mode into mc_model_desc (include/microclimate_b200.h)."""
    m = A.ModelDesc()
    depths = np.asarray(model.depths, float)
    n = len(depths)
    if not (3 <= n <= A.MC_MAX_SOIL_NODES):
        raise ValueError(f"number of soil nodes must be in 3..{A.MC_MAX_SOIL_NODES}, got {n}")
    m.n_soil = n
    for i, d in enumerate(depths):
        m.depths_m[i] = d
    heights = np.asarray(model.heights, float)
    m.reference_height_m = heights[-1]
    m.local_height_m = heights[0]
    if len(heights) > A.MC_MAX_HEIGHTS:
        raise ValueError(f"at most {A.MC_MAX_HEIGHTS} heights")
    m.n_heights = len(heights)
    for i, hgt in enumerate(heights):
        m.heights_m[i] = hgt
    sem = model.soil_energy_model
    m.reltol, m.abstol, m.maximum_surface_temperature_K = sem.reltol, sem.abstol, sem.maximum_surface_temperature
    sp = model.soil_properties_model
    m.de_vries_shape_factor, m.recirculation_power, m.return_flow_threshold = (
        sp.de_vries_shape_factor, sp.recirculation_power, sp.return_flow_threshold)
    sh = model.soil_hydraulic_model
    m.root_resistance, m.stomatal_closure_potential, m.leaf_resistance = (
        sh.root_resistance, sh.stomatal_closure_potential, sh.leaf_resistance)
    m.stomatal_stability_parameter, m.root_radius = sh.stomatal_stability_parameter, sh.root_radius
    m.infiltration_algorithm = (A.MC_INFIL_MATRIC_FLUX_POTENTIAL if isinstance(sh.infiltration_algorithm, T.MatricFluxPotentialAlgorithm)
                                else A.MC_INFIL_MATRIC_POTENTIAL)
    m.karman_constant, m.dyer_constant = model.boundary_layer_model.karman_constant, model.boundary_layer_model.dyer_constant
    cfg = model.config
    if isinstance(time_mode, T.NonConsecutiveDayMode):
        m.time_mode, m.iterations_per_day, m.spinup_first_day = A.MC_TIME_NONCONSECUTIVE, time_mode.iterations_per_day, 0
    elif isinstance(time_mode, T.ConsecutiveDayMode):
        m.time_mode, m.iterations_per_day, m.spinup_first_day = A.MC_TIME_CONSECUTIVE, 1, int(time_mode.spinup_first_day)
    else:
        raise TypeError(f"unknown time mode {time_mode!r}")
    cv = cfg.convergence
    if isinstance(cv, T.FixedSoilTemperatureIterations):
        m.convergence_kind, m.convergence_iterations, m.convergence_tolerance_K = A.MC_CONV_FIXED, cv.iterations_per_day, 0.0
    elif isinstance(cv, T.SoilTemperatureConvergenceTolerance):
        m.convergence_kind, m.convergence_iterations, m.convergence_tolerance_K = (
            A.MC_CONV_TOLERANCE, cv.max_iterations_per_day, cv.tolerance)
    else:
        raise TypeError(f"unknown convergence strategy {cv!r}")
    m.rainfall_schedule = A.MC_RAIN_HOURLY if isinstance(cfg.rainfall_schedule, T.HourlyRainfall) else A.MC_RAIN_DAILY
    sm = cfg.soil_moisture_strategy
    if isinstance(sm, T.DynamicSoilMoisture):
        m.moisture_strategy = A.MC_MOIST_DYNAMIC
        m.moisture_tolerance, m.moisture_max_iterations, m.moisture_timestep_s = (
            sm.moisture_tolerance, sm.moisture_max_iterations, sm.moisture_timestep)
    else:
        m.moisture_strategy = A.MC_MOIST_PRESCRIBED
        m.moisture_tolerance, m.moisture_max_iterations, m.moisture_timestep_s = 1e-6, 500, 360.0
    m.max_surface_pool = cfg.max_surface_pool
    arm = model.radiation.longwave_model.atmospheric_radiation_model
    m.atmospheric_radiation = (A.MC_ATMRAD_SWINBANK if isinstance(arm, T.SwinbankAtmosphericRadiation)
                               else A.MC_ATMRAD_CAMPBELL_NORMAN)
    m.fluid = A.default_fluid()
    sn = model.snow_model
    if isinstance(sn, T.SnowModel):
        m.n_snow = A.MC_SNOW_NODES
        s = m.snow
        s.snow_temperature_threshold_C, s.snow_density, s.snow_melt_factor = sn.snow_temperature_threshold, sn.snow_density, sn.snow_melt_factor
        s.undercatch, s.rain_multiplier, s.rain_melt_factor = sn.undercatch, sn.rain_multiplier, sn.rain_melt_factor
        for i in range(4):
            s.density_function[i] = sn.density_function[i]
        s.snow_conductivity, s.canopy_interception, s.min_snow_depth_cm = sn.snow_conductivity, sn.canopy_interception, sn.min_snow_depth
        if len(sn.snow_node_thresholds) != A.MC_SNOW_NODES:
            raise ValueError("SnowModel must have 8 node thresholds")
        for i in range(A.MC_SNOW_NODES):
            s.node_thresholds_cm[i] = sn.snow_node_thresholds[i]
        s.melt_threshold_C = sn.melt_threshold
        ahc = sn.apparent_heat_capacity
        s.ahc_T_low_C, s.ahc_T_high_C, s.ahc_window_K, s.ahc_smoothing_K, s.ahc_T_melt_C, s.ahc_width_K = -0.45, 0.4, 1.0, 0.1, 0.0, 0.2
        if isinstance(ahc, T.BonacinaStep):
            s.ahc_kind, s.ahc_T_low_C, s.ahc_T_high_C, s.ahc_window_K = A.MC_AHC_BONACINA, ahc.T_low, ahc.T_high, ahc.window
        elif isinstance(ahc, T.TanhSmoothed):
            s.ahc_kind, s.ahc_T_low_C, s.ahc_T_high_C, s.ahc_window_K, s.ahc_smoothing_K = A.MC_AHC_TANH, ahc.T_low, ahc.T_high, ahc.window, ahc.smoothing
        elif isinstance(ahc, T.Gaussian):
            s.ahc_kind, s.ahc_T_melt_C, s.ahc_width_K = A.MC_AHC_GAUSSIAN, ahc.T_melt, ahc.width
        elif isinstance(ahc, T.WestermannSigmoid):
            s.ahc_kind, s.ahc_T_melt_C, s.ahc_width_K = A.MC_AHC_WESTERMANN, ahc.T_melt, ahc.width
        else:
            raise TypeError(f"unknown apparent heat capacity {ahc!r}")
        if inputs0 is not None:
            s.initial_snow_depth_cm = inputs0.initial_snow_depth
            s.initial_snow_temperature_K = inputs0.initial_snow_temperature
            s.initial_snow_density = -1.0 if inputs0.initial_snow_density is None else inputs0.initial_snow_density
        else:
            s.initial_snow_depth_cm, s.initial_snow_temperature_K, s.initial_snow_density = 0.0, 273.15, -1.0
    else:
        m.n_snow = 0
    return m


class SitesBatch:
    """SoA buffers for S sites: mc_sites_soa + mc_initial_soa. Keeps the numpy arrays alive."""

    NODE_FIELDS = ("bulk_density", "mineral_density", "mineral_conductivity", "mineral_heat_capacity",
                   "air_entry_water_potential", "saturated_hydraulic_conductivity", "campbell_b_parameter", "root_density")
    SITE_FIELDS = ("roughness_height", "slope", "sky_view_fraction", "albedo")

    def __init__(self, n_soil: int, n_sites: int, has_initial_temperature: bool):
        self.N, self.S = n_soil, n_sites
        for f in self.SITE_FIELDS:
            setattr(self, f, np.zeros(n_sites))
        for f in self.NODE_FIELDS:
            setattr(self, f, np.zeros((n_soil, n_sites)))
        self.initial_soil_moisture = np.zeros((n_soil, n_sites))
        self.initial_soil_temperature = np.zeros((n_soil, n_sites)) if has_initial_temperature else None

    @classmethod
    def from_inputs(cls, model: T.MicroModel, inputs: Sequence[T.MicroInputs]) -> "SitesBatch":
        n = len(model.depths)
        has_T = inputs[0].initial_soil_temperature is not None
        b = cls(n, len(inputs), has_T)
        for s, inp in enumerate(inputs):
            if (inp.initial_soil_temperature is not None) != has_T:
                raise ValueError("initial_soil_temperature must be given for all sites or for none")
            if has_T and len(inp.initial_soil_temperature) != n:
                raise ValueError("Initial soil temperature must match length of 'depths'")   # src/simulation.jl:399-401
            b.roughness_height[s] = inp.site.roughness_height
            b.slope[s] = inp.site.slope
            b.sky_view_fraction[s] = inp.site.sky_view_fraction
            b.albedo[s] = inp.site.albedo
            sp, hy = inp.soil_profile, inp.soil_profile.hydraulics
            b.bulk_density[:, s] = sp.bulk_density
            b.mineral_density[:, s] = sp.mineral_density
            b.mineral_conductivity[:, s] = sp.mineral_conductivity
            b.mineral_heat_capacity[:, s] = sp.mineral_heat_capacity
            b.air_entry_water_potential[:, s] = hy.air_entry_water_potential
            b.saturated_hydraulic_conductivity[:, s] = hy.saturated_hydraulic_conductivity
            b.campbell_b_parameter[:, s] = hy.campbell_b_parameter
            b.root_density[:, s] = hy.root_density
            b.initial_soil_moisture[:, s] = inp.initial_soil_moisture
            if has_T:
                b.initial_soil_temperature[:, s] = inp.initial_soil_temperature
        return b

    def c_sites(self) -> A.SitesSoA:
        c = A.SitesSoA()
        for f in self.SITE_FIELDS + self.NODE_FIELDS:
            setattr(c, f, _ptr(getattr(self, f)))
        return c

    def c_initial(self) -> A.InitialSoA:
        c = A.InitialSoA()
        c.initial_soil_temperature = _ptr(self.initial_soil_temperature)
        c.initial_soil_moisture = _ptr(self.initial_soil_moisture)
        return c

    def slice(self, lo: int, hi: int) -> "SitesBatch":
        """Contiguous site range [lo, hi) — the unit of multi-GPU sharding."""
        b = SitesBatch.__new__(SitesBatch)
        b.N, b.S = self.N, hi - lo
        for f in self.SITE_FIELDS:
            setattr(b, f, np.ascontiguousarray(getattr(self, f)[lo:hi]))
        for f in self.NODE_FIELDS + ("initial_soil_moisture",):
            setattr(b, f, np.ascontiguousarray(getattr(self, f)[:, lo:hi]))
        b.initial_soil_temperature = (None if self.initial_soil_temperature is None
                                      else np.ascontiguousarray(self.initial_soil_temperature[:, lo:hi]))
        return b


class ForcingBatch:
    """mc_forcing_soa for a chunk of ndays days: hourly [24*ndays][S], daily [ndays][S]."""

    def __init__(self, ndays: int, n_sites: int, *, hourly_rain=False, longwave=False, albedo_daily=False,
                 prescribed_moisture_nodes: int = 0, day_of_year: Optional[Sequence[int]] = None):
        self.ndays, self.S = ndays, n_sites
        H = 24 * ndays
        for f in A.FORCING_HOURLY:
            setattr(self, f, np.zeros((H, n_sites)))
        for f in A.FORCING_DAILY:
            setattr(self, f, np.zeros((ndays, n_sites)))
        if not hourly_rain:
            self.rainfall_hourly = None
        if not longwave:
            self.longwave_radiation = None
        if not albedo_daily:
            self.albedo_daily = None
        self.prescribed_soil_moisture = (np.zeros((ndays, prescribed_moisture_nodes, n_sites))
                                         if prescribed_moisture_nodes else None)
        self.day_of_year = np.asarray(day_of_year if day_of_year is not None else np.arange(1, ndays + 1), dtype=np.int32)

    @classmethod
    def from_inputs(cls, model: T.MicroModel, inputs: Sequence[T.MicroInputs], days: Sequence[int]) -> "ForcingBatch":
        ndays, S, N = len(days), len(inputs), len(model.depths)
        cfg = model.config
        h0, d0 = inputs[0].environment_hourly, inputs[0].environment_daily
        pre = cfg.soil_moisture_strategy
        has_pre = isinstance(pre, T.PrescribedSoilMoisture) and pre.precomputed_soil_moisture is not None
        b = cls(ndays, S, hourly_rain=isinstance(cfg.rainfall_schedule, T.HourlyRainfall),
                longwave=h0.longwave_radiation is not None, albedo_daily=d0.albedo is not None,
                prescribed_moisture_nodes=N if has_pre else 0, day_of_year=days)
        H = 24 * ndays
        for s, inp in enumerate(inputs):
            h, d = inp.environment_hourly, inp.environment_daily

            def hv(x):
                if x is None:               # a min/max-driven problem leaves these `nothing` (test/micro_testrun_monthly.jl:66-76): populate_weather! fills them
                    return np.zeros(H)
                x = np.asarray(x, float)
                return np.broadcast_to(x, (H,)) if x.ndim == 0 or x.shape[0] == 1 else x[:H]
            b.pressure[:, s] = hv(h.pressure)
            b.reference_temperature[:, s] = hv(h.reference_temperature)
            b.reference_wind_speed[:, s] = hv(h.reference_wind_speed)
            b.reference_humidity[:, s] = hv(h.reference_humidity)
            b.cloud_cover[:, s] = hv(h.cloud_cover)
            b.global_radiation[:, s] = hv(h.global_radiation)
            b.zenith_angle[:, s] = np.minimum(hv(h.zenith_angle), 90.0)        # src/simulation.jl:301
            b.diffuse_fraction[:, s] = 0.0 if h.diffuse_fraction is None else hv(h.diffuse_fraction)
            b.zenith_slope_angle[:, s] = (b.zenith_angle[:, s] if h.zenith_slope_angle is None
                                          else np.minimum(hv(h.zenith_slope_angle), 90.0))
            if b.rainfall_hourly is not None:
                b.rainfall_hourly[:, s] = hv(h.rainfall)
            if b.longwave_radiation is not None:
                b.longwave_radiation[:, s] = hv(h.longwave_radiation)
            b.leaf_area_index[:, s] = d.leaf_area_index[:ndays]
            b.shade[:, s] = d.shade[:ndays]
            b.surface_emissivity[:, s] = d.surface_emissivity[:ndays]
            b.cloud_emissivity[:, s] = d.cloud_emissivity[:ndays]
            b.soil_wetness[:, s] = d.soil_wetness[:ndays]
            b.deep_soil_temperature[:, s] = d.deep_soil_temperature[:ndays]
            b.rainfall_daily[:, s] = d.rainfall[:ndays]
            if b.albedo_daily is not None:
                b.albedo_daily[:, s] = d.albedo[:ndays]
            if has_pre:
                b.prescribed_soil_moisture[:, :, s] = np.asarray(pre.precomputed_soil_moisture, float)[:, :ndays].T
        return b

    def c_struct(self) -> A.ForcingSoA:
        c = A.ForcingSoA()
        for f in A.FORCING_HOURLY + A.FORCING_DAILY:
            setattr(c, f, _ptr(getattr(self, f)))
        c.prescribed_soil_moisture = _ptr(self.prescribed_soil_moisture)
        c.day_of_year = self.day_of_year.ctypes.data_as(C.POINTER(C.c_int32))
        return c

    def day_slice(self, d0: int, d1: int) -> "ForcingBatch":
        b = ForcingBatch.__new__(ForcingBatch)
        b.ndays, b.S = d1 - d0, self.S
        for f in A.FORCING_HOURLY:
            a = getattr(self, f)
            setattr(b, f, None if a is None else np.ascontiguousarray(a[24 * d0:24 * d1]))
        for f in A.FORCING_DAILY:
            a = getattr(self, f)
            setattr(b, f, None if a is None else np.ascontiguousarray(a[d0:d1]))
        b.prescribed_soil_moisture = (None if self.prescribed_soil_moisture is None
                                      else np.ascontiguousarray(self.prescribed_soil_moisture[d0:d1]))
        b.day_of_year = np.ascontiguousarray(self.day_of_year[d0:d1])
        return b

    def site_slice(self, lo: int, hi: int) -> "ForcingBatch":
        b = ForcingBatch.__new__(ForcingBatch)
        b.ndays, b.S = self.ndays, hi - lo
        for f in A.FORCING_HOURLY + A.FORCING_DAILY:
            a = getattr(self, f)
            setattr(b, f, None if a is None else np.ascontiguousarray(a[:, lo:hi]))
        b.prescribed_soil_moisture = (None if self.prescribed_soil_moisture is None
                                      else np.ascontiguousarray(self.prescribed_soil_moisture[:, :, lo:hi]))
        b.day_of_year = self.day_of_year
        return b


DEFAULT_OUTPUT_FIELDS = A.OUTPUT_NODE_FIELDS + ("surface_water", "sky_temperature")


class OutputBatch:
    """mc_output_soa for a chunk: node fields [24*ndays][N][S], scalars [24*ndays][S].  Unselected = None."""

    def __init__(self, ndays: int, n_soil: int, n_sites: int, fields: Sequence[str] = DEFAULT_OUTPUT_FIELDS, n_snow: int = 0,
                 n_heights: int = 0):
        self.ndays, self.N, self.S = ndays, n_soil, n_sites
        H = 24 * ndays
        for f in A.OUTPUT_PROFILE_FIELDS:
            setattr(self, f, np.zeros((H, n_heights, n_sites)) if (n_heights and f in fields) else None)
        for f in A.OUTPUT_PROFILE_SCALARS:
            setattr(self, f, np.zeros((H, n_sites)) if (n_heights and f in fields) else None)
        for f in A.OUTPUT_NODE_FIELDS:
            setattr(self, f, np.zeros((H, n_soil, n_sites)) if f in fields else None)
        for f in A.OUTPUT_SCALAR_FIELDS:
            setattr(self, f, np.zeros((H, n_sites)) if f in fields else None)
        self.snow_temperature = np.zeros((H, n_snow, n_sites)) if (n_snow and "snow_temperature" in fields) else None
        self.status = np.zeros(n_sites, dtype=np.int32)

    def c_struct(self) -> A.OutputSoA:
        c = A.OutputSoA()
        for f in A.OUTPUT_NODE_FIELDS + A.OUTPUT_SCALAR_FIELDS:
            setattr(c, f, _ptr(getattr(self, f)))
        c.snow_temperature = _ptr(self.snow_temperature)
        for f in A.OUTPUT_PROFILE_FIELDS + A.OUTPUT_PROFILE_SCALARS:
            setattr(c, f, _ptr(getattr(self, f, None)))
        c.status = self.status.ctypes.data_as(C.POINTER(C.c_int32))
        return c

    def assign_rows(self, row0: int, part: "OutputBatch"):
        """Copy a day-chunk's rows into this batch at row offset `row0` (time-chunked runs)."""
        H = 24 * part.ndays
        for f in A.OUTPUT_NODE_FIELDS + A.OUTPUT_SCALAR_FIELDS + ("snow_temperature",) + A.OUTPUT_PROFILE_FIELDS + A.OUTPUT_PROFILE_SCALARS:
            dst, src = getattr(self, f, None), getattr(part, f, None)
            if dst is not None and src is not None:
                dst[row0:row0 + H] = src
        self.status |= part.status

    def site_result(self, s: int) -> T.MicroResult:
        """Un-marshal one site into the reference's per-site (nsteps, N) layout."""
        r = T.MicroResult()
        for f in A.OUTPUT_NODE_FIELDS:
            a = getattr(self, f)
            setattr(r, f, None if a is None else np.ascontiguousarray(a[:, :, s]))
        for f in A.OUTPUT_SCALAR_FIELDS:
            a = getattr(self, f)
            setattr(r, f, None if a is None else np.ascontiguousarray(a[:, s]))
        r.snow_temperature = None if self.snow_temperature is None else np.ascontiguousarray(self.snow_temperature[:, :, s])
        for f in A.OUTPUT_PROFILE_FIELDS:
            a = getattr(self, f, None)
            setattr(r, f, None if a is None else np.ascontiguousarray(a[:, :, s]))
        for f in A.OUTPUT_PROFILE_SCALARS:
            a = getattr(self, f, None)
            setattr(r, f, None if a is None else np.ascontiguousarray(a[:, s]))
        r.status = int(self.status[s])
        return r


class StateBatch:
    """mc_state_soa host buffers (checkpoint / time-chunking)."""

    def __init__(self, n_soil: int, n_snow: int, n_sites: int):
        M = n_soil + n_snow
        self.soil_temperature = np.zeros((n_soil, n_sites))
        self.soil_moisture = np.zeros((n_soil, n_sites))
        self.sum_phase = np.zeros((n_soil, n_sites))
        self.pool = np.zeros(n_sites)
        self.soil_wetness = np.zeros(n_sites)
        self.obukhov_ode = np.zeros(n_sites)
        self.obukhov_profile = np.zeros(n_sites)
        self.deep_temperature_prev = np.zeros(n_sites)
        self.fsal = np.zeros((M, n_sites))
        self.snow_temperature = np.zeros((n_snow, n_sites)) if n_snow else None
        self.snow_state = np.zeros((A.MC_SNOW_STATE_FIELDS, n_sites)) if n_snow else None
        self.snow_node_depths = np.zeros((n_snow, n_sites)) if n_snow else None
        self.q_freeze = np.zeros(n_sites) if n_snow else None
        self.obukhov_air = np.zeros(n_sites)
        self.days_done = np.zeros(1, dtype=np.int64)
        self.status = np.zeros(n_sites, dtype=np.int32)          # MC_STATUS_* bits accumulated so far

    def c_struct(self) -> A.StateSoA:
        c = A.StateSoA()
        for f, _ in A.StateSoA._fields_:
            if f == "days_done":
                c.days_done = self.days_done.ctypes.data_as(C.POINTER(C.c_int64))
            elif f == "status":
                c.status = self.status.ctypes.data_as(C.POINTER(C.c_int32))
            else:
                setattr(c, f, _ptr(getattr(self, f)))
        return c
