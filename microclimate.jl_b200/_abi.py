"""ctypes mirror of include/microclimate_b200.h (the C-ABI structs and enums).

Pure data layout — no compute.  Shared by the product loader (`_lib.py`) and by the test-only
oracle wrapper (`oracle/oracle.py`), so both sides are fed byte-identical buffers.
"""
import ctypes as C

MC_ABI_VERSION = 2
MC_MAX_SOIL_NODES = 24
MC_SNOW_NODES = 8
MC_MAX_ODE_NODES = MC_MAX_SOIL_NODES + MC_SNOW_NODES
MC_SNOW_STATE_FIELDS = 10
MC_MAX_HEIGHTS = 16

MC_TIME_NONCONSECUTIVE, MC_TIME_CONSECUTIVE = 0, 1
MC_CONV_FIXED, MC_CONV_TOLERANCE = 0, 1
MC_RAIN_DAILY, MC_RAIN_HOURLY = 0, 1
MC_MOIST_PRESCRIBED, MC_MOIST_DYNAMIC = 0, 1
MC_INFIL_MATRIC_POTENTIAL, MC_INFIL_MATRIC_FLUX_POTENTIAL = 0, 1
MC_ATMRAD_CAMPBELL_NORMAN, MC_ATMRAD_SWINBANK = 0, 1
MC_AHC_BONACINA, MC_AHC_TANH, MC_AHC_GAUSSIAN, MC_AHC_WESTERMANN = 0, 1, 2, 3

MC_STATUS_NAN, MC_STATUS_DT_UNDERFLOW, MC_STATUS_NEWTON_MAXITER, MC_STATUS_STEP_LIMIT = 1, 2, 4, 8

_d = C.c_double
_pd = C.POINTER(C.c_double)
_i32 = C.c_int32


class FluidConstants(C.Structure):
    _fields_ = [(n, _d) for n in (
        "molar_mass_water", "molar_mass_air", "ideal_gas_departure", "cp_dry_air", "cp_water_vapour",
        "gas_constant", "stefan_boltzmann", "standard_gravity", "standard_pressure",
        "latent_heat_a", "latent_heat_b", "molar_latent_heat_a", "molar_latent_heat_b")]


class SnowDesc(C.Structure):
    _fields_ = [
        ("snow_temperature_threshold_C", _d), ("snow_density", _d), ("snow_melt_factor", _d),
        ("undercatch", _d), ("rain_multiplier", _d), ("rain_melt_factor", _d),
        ("density_function", _d * 4), ("snow_conductivity", _d), ("canopy_interception", _d),
        ("min_snow_depth_cm", _d), ("node_thresholds_cm", _d * MC_SNOW_NODES), ("melt_threshold_C", _d),
        ("ahc_kind", _i32),
        ("ahc_T_low_C", _d), ("ahc_T_high_C", _d), ("ahc_window_K", _d), ("ahc_smoothing_K", _d),
        ("ahc_T_melt_C", _d), ("ahc_width_K", _d),
        ("initial_snow_depth_cm", _d), ("initial_snow_temperature_K", _d), ("initial_snow_density", _d),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_soil", _i32), ("n_snow", _i32),
        ("depths_m", _d * MC_MAX_SOIL_NODES),
        ("reference_height_m", _d), ("local_height_m", _d),
        ("n_heights", _i32), ("_pad1", _i32), ("heights_m", _d * MC_MAX_HEIGHTS),
        ("reltol", _d), ("abstol", _d), ("maximum_surface_temperature_K", _d),
        ("de_vries_shape_factor", _d), ("recirculation_power", _d), ("return_flow_threshold", _d),
        ("root_resistance", _d), ("stomatal_closure_potential", _d), ("leaf_resistance", _d),
        ("stomatal_stability_parameter", _d), ("root_radius", _d),
        ("infiltration_algorithm", _i32),
        ("karman_constant", _d), ("dyer_constant", _d),
        ("time_mode", _i32), ("iterations_per_day", _i32), ("spinup_first_day", _i32),
        ("convergence_kind", _i32), ("convergence_iterations", _i32),
        ("convergence_tolerance_K", _d),
        ("rainfall_schedule", _i32), ("moisture_strategy", _i32),
        ("moisture_tolerance", _d), ("moisture_max_iterations", _i32), ("moisture_timestep_s", _d),
        ("max_surface_pool", _d),
        ("atmospheric_radiation", _i32), ("_pad0", _i32),
        ("fluid", FluidConstants), ("snow", SnowDesc),
    ]


class SitesSoA(C.Structure):
    _fields_ = [(n, _pd) for n in (
        "roughness_height", "slope", "sky_view_fraction", "albedo",
        "bulk_density", "mineral_density", "mineral_conductivity", "mineral_heat_capacity",
        "air_entry_water_potential", "saturated_hydraulic_conductivity", "campbell_b_parameter", "root_density")]


class InitialSoA(C.Structure):
    _fields_ = [("initial_soil_temperature", _pd), ("initial_soil_moisture", _pd)]


class StateSoA(C.Structure):
    _fields_ = [(n, _pd) for n in (
        "soil_temperature", "soil_moisture", "sum_phase", "pool", "soil_wetness", "obukhov_ode",
        "obukhov_profile", "deep_temperature_prev", "fsal", "snow_temperature", "snow_state",
        "snow_node_depths", "q_freeze", "obukhov_air")] + [("days_done", C.POINTER(C.c_int64)), ("status", C.POINTER(_i32))]


FORCING_HOURLY = ("pressure", "reference_temperature", "reference_wind_speed", "reference_humidity",
                  "cloud_cover", "global_radiation", "diffuse_fraction", "zenith_angle", "zenith_slope_angle",
                  "rainfall_hourly", "longwave_radiation")
FORCING_DAILY = ("leaf_area_index", "shade", "surface_emissivity", "cloud_emissivity", "soil_wetness",
                 "deep_soil_temperature", "rainfall_daily", "albedo_daily")


class ForcingSoA(C.Structure):
    _fields_ = ([(n, _pd) for n in FORCING_HOURLY] + [(n, _pd) for n in FORCING_DAILY]
                + [("prescribed_soil_moisture", _pd), ("day_of_year", C.POINTER(_i32))])


OUTPUT_NODE_FIELDS = ("soil_temperature", "soil_moisture", "soil_water_potential", "soil_humidity",
                      "soil_thermal_conductivity", "soil_heat_capacity", "soil_bulk_density")
OUTPUT_SCALAR_FIELDS = ("surface_water", "sky_temperature", "snow_fall", "snow_depth", "snow_density")
OUTPUT_PROFILE_FIELDS = ("air_temperature", "wind_speed", "relative_humidity")          # [H][n_heights][S]
OUTPUT_PROFILE_SCALARS = ("convective_heat_flux", "friction_velocity")                  # [H][S]


class OutputSoA(C.Structure):
    _fields_ = ([(n, _pd) for n in OUTPUT_NODE_FIELDS] + [(n, _pd) for n in OUTPUT_SCALAR_FIELDS]
                + [("snow_temperature", _pd)] + [(n, _pd) for n in OUTPUT_PROFILE_FIELDS + OUTPUT_PROFILE_SCALARS]
                + [("status", C.POINTER(_i32))])


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_rhs", "n_accept", "n_reject", "n_obukhov_iter", "n_newton_iter", "n_infil_steps",
        "n_site_days", "n_day_passes")] + [("gpu_ms", _d), ("n_kernel_launches", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def default_fluid():
    """mc_default_fluid(): NicheMapR WETAIR/VAPPRS lineage constants (FluidProperties.jl is un-vendored)."""
    f = FluidConstants()
    f.molar_mass_water = 0.018016
    f.molar_mass_air = 0.028966
    f.ideal_gas_departure = 1.0053
    f.cp_dry_air = 1004.84
    f.cp_water_vapour = 1846.40
    f.gas_constant = 8.314462618
    f.stefan_boltzmann = 5.670374419e-8
    f.standard_gravity = 9.80665
    f.standard_pressure = 101325.0
    f.latent_heat_a = 2.5012e6
    f.latent_heat_b = 2.3787e3
    f.molar_latent_heat_a = 45144.0
    f.molar_latent_heat_b = 48.0
    return f


# ---- widening rows f2 / f3: populate_weather! + solar geometry (include/microclimate_b200.h, "Widening") ----
MC_TOD_CLOCK, MC_TOD_SUNRISE, MC_TOD_SUNSET, MC_TOD_MIDDAY, MC_TOD_MIDNIGHT = 0, 1, 2, 3, 4
MC_SHAPE_SINE, MC_SHAPE_DECAY, MC_SHAPE_LINEAR = 0, 1, 2
MC_DIEL_MAX_SHAPES = 6
MC_DIEL_MAX_INPUTS = 6
MC_DIFFUSE_ERBS = 0
MC_HUMIDITY_CURVE, MC_HUMIDITY_FROM_VAPOUR_PRESSURE = 0, 1


class TimeOfDay(C.Structure):
    _fields_ = [("kind", _i32), ("_pad", _i32), ("value", _d)]


class DielShape(C.Structure):
    _fields_ = [("kind", _i32), ("_pad", _i32), ("a", TimeOfDay), ("b", TimeOfDay)]


class DielCurve(C.Structure):
    _fields_ = [("n_shapes", _i32), ("n_inputs", _i32), ("shapes", DielShape * MC_DIEL_MAX_SHAPES), ("inputs", TimeOfDay * MC_DIEL_MAX_INPUTS)]


WEATHER_QUANTITIES = ("reference_temperature", "reference_wind_speed", "reference_humidity", "cloud_cover")


class WeatherDesc(C.Structure):
    _fields_ = ([(n, DielCurve) for n in WEATHER_QUANTITIES]
                + [("angstrom_a", _d), ("angstrom_b", _d), ("angstrom_gamma", _d), ("solar_constant", _d),
                   ("diffuse_model", _i32), ("refraction", _i32), ("humidity_source", _i32), ("_pad0", _i32),
                   ("orbital_eccentricity", _d), ("sin_obliquity", _d), ("days_per_year", _d), ("equinox_day", _d)])


SOLAR_SITE_FIELDS = ("latitude", "slope", "aspect", "solar_noon_offset_h")


class SolarSitesSoA(C.Structure):
    _fields_ = [(n, _pd) for n in SOLAR_SITE_FIELDS]


WEATHER_IN_DAILY = ("hour_solar_noon", "hour_angle_sunrise")
WEATHER_IN_HOURLY = ("direct_horizontal", "diffuse_horizontal", "global_radiation_hourly", "pressure_hourly")


class WeatherInSoA(C.Structure):
    _fields_ = ([(n + "_values", _pd * MC_DIEL_MAX_INPUTS) for n in WEATHER_QUANTITIES]
                + [(n, _pd) for n in WEATHER_IN_DAILY] + [(n, _pd) for n in WEATHER_IN_HOURLY]
                + [("pressure_site", _pd), ("day_of_year", C.POINTER(_i32))])


WEATHER_OUT_HOURLY = ("pressure", "reference_temperature", "reference_wind_speed", "reference_humidity", "cloud_cover",
                      "global_radiation", "diffuse_fraction", "zenith_angle", "zenith_slope_angle")
WEATHER_OUT_DAILY = ("hour_solar_noon", "hour_angle_sunrise")


class WeatherOutSoA(C.Structure):
    _fields_ = [(n, _pd) for n in WEATHER_OUT_HOURLY + WEATHER_OUT_DAILY]


# mc_abi_sizeof(which): the order of include/microclimate_b200.h
ABI_SIZEOF = (ModelDesc, SitesSoA, InitialSoA, StateSoA, ForcingSoA, OutputSoA, Counters, WeatherDesc, SolarSitesSoA, WeatherInSoA, WeatherOutSoA)
