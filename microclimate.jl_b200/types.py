"""Host-side mirror of the reference's site / forcing / model / output types for the soil path.

Same names and field meanings as Microclimate.jl (file:line cited per class); values are plain
unit-stripped SI floats / numpy arrays instead of Unitful quantities:
K, Pa, m, s, W/m2, kg/m2, J/kg, kg s/m3, fractions 0..1, angles in degrees.
Only what `solve_soil!` (src/simulation.jl:355-819) reads or writes is mirrored.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

# src/constants.jl:7-13
DEFAULT_HEIGHTS = np.array([0.01, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.75, 1.0, 1.2])
DEFAULT_DEPTHS = np.array([0.0, 1.25, 2.5, 3.75, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0, 25.0, 30.0, 40.0, 50.0,
                           75.0, 100.0, 150.0, 200.0]) / 100.0  # metres
DEFAULT_DAYS = [15, 46, 74, 105, 135, 166, 196, 227, 258, 288, 319, 349]
DEFAULT_SNOW_NODE_THRESHOLDS = (2.0, 5.0, 10.0, 20.0, 50.0, 100.0, 200.0, 300.0)  # cm


def atmospheric_pressure(elevation_m: float) -> float:
    """FluidProperties.atmospheric_pressure(elevation): standard-atmosphere pressure (Pa)."""
    return 101325.0 * (1.0 - 0.0065 * elevation_m / 288.0) ** (1.0 / 0.190284)


# ---- time modes: src/time_mode/{non_consecutive,consecutive}.jl ---------------------------------
@dataclass(frozen=True)
class NonConsecutiveDayMode:
    iterations_per_day: int = 3


@dataclass(frozen=True)
class ConsecutiveDayMode:
    spinup_first_day: bool = False


# ---- convergence: src/soil_temperature_convergence/{fixed,tolerance}.jl ----------------------
@dataclass(frozen=True)
class FixedSoilTemperatureIterations:
    iterations_per_day: int = 3


@dataclass(frozen=True)
class SoilTemperatureConvergenceTolerance:
    tolerance: float = 0.1            # K
    max_iterations_per_day: int = 10


# ---- rainfall: src/rainfall/*.jl ---------------------------------------------------------------
@dataclass(frozen=True)
class DailyRainfall:
    pass


@dataclass(frozen=True)
class HourlyRainfall:
    pass


# ---- soil moisture strategy: src/soil_moisture/{prescribed,dynamic}.jl -------------------------
@dataclass
class PrescribedSoilMoisture:
    precomputed_soil_moisture: Optional[np.ndarray] = None   # (ndepths, ndays)


@dataclass
class DynamicSoilMoisture:
    moisture_tolerance: float = 1e-6          # kg/m2/s
    moisture_max_iterations: int = 500
    moisture_timestep: float = 360.0          # s


# ---- infiltration algorithms: src/soil_hydraulics/infiltration_algorithm/*.jl -----------------
@dataclass(frozen=True)
class MatricPotentialAlgorithm:
    pass


@dataclass(frozen=True)
class MatricFluxPotentialAlgorithm:
    pass


# ---- atmospheric radiation: src/radiation/atmospheric/*.jl --------------------------------------
@dataclass(frozen=True)
class CampbellNormanAtmosphericRadiation:
    pass


@dataclass(frozen=True)
class SwinbankAtmosphericRadiation:
    pass


@dataclass(frozen=True)
class ViewFactorLongwave:                       # src/radiation/longwave/view_factor.jl:12-14
    atmospheric_radiation_model: object = CampbellNormanAtmosphericRadiation()


@dataclass(frozen=True)
class RadiationModel:                           # src/types.jl:52-56 (the solar_radiation_model stays SolarRadiation.jl's)
    longwave_model: ViewFactorLongwave = ViewFactorLongwave()
    shortwave_model: object = None              # weather.AngstromMaxwellShortwave; None = its defaults (used by solve_minmax_batched / populate_weather)


# ---- apparent heat capacity: src/apparent_heat_capacity/*.jl ------------------------------------
@dataclass(frozen=True)
class BonacinaStep:
    T_low: float = -0.45     # degC
    T_high: float = 0.4      # degC
    window: float = 1.0      # K


@dataclass(frozen=True)
class TanhSmoothed:
    T_low: float = -0.45
    T_high: float = 0.4
    window: float = 1.0
    smoothing: float = 0.1


@dataclass(frozen=True)
class Gaussian:
    T_melt: float = 0.0
    width: float = 0.2


@dataclass(frozen=True)
class WestermannSigmoid:
    T_melt: float = 0.0
    width: float = 0.2


# ---- snow: src/snow/{no_snow,snow_model}.jl ---------------------------------------------------
@dataclass(frozen=True)
class NoSnow:
    pass


@dataclass(frozen=True)
class SnowModel:                                 # src/snow/snow_model.jl:3-42
    snow_temperature_threshold: float = 1.5      # degC
    snow_density: float = 0.375                  # g/cm3
    snow_melt_factor: float = 1.0
    undercatch: float = 1.0
    rain_multiplier: float = 1.0
    rain_melt_factor: float = 0.0125
    density_function: tuple = (0.0, 0.0, 0.0, 0.0)
    snow_conductivity: float = 0.0               # W/m/K, 0 = Aggarwal
    canopy_interception: float = 0.0
    min_snow_depth: float = 2.0                  # cm
    snow_node_thresholds: tuple = DEFAULT_SNOW_NODE_THRESHOLDS
    melt_threshold: float = 0.4                  # degC
    apparent_heat_capacity: object = BonacinaStep()


# ---- physical-process models ---------------------------------------------------------------------
@dataclass(frozen=True)
class CampbelldeVriesSoilProperties:             # src/soil_properties/campbell_devries.jl:5-9,11-20
    de_vries_shape_factor: float = 0.1
    recirculation_power: float = 4.0
    return_flow_threshold: float = 0.162


@dataclass(frozen=True)
class CampbellSoilHydraulics:                    # src/soil_hydraulics/campbell.jl:47-54,69-81
    root_resistance: float = 2.5e10              # m3/kg/s
    stomatal_closure_potential: float = -1500.0  # J/kg
    leaf_resistance: float = 2.0e6               # m4/kg/s
    stomatal_stability_parameter: float = 10.0
    root_radius: float = 0.001                   # m
    infiltration_algorithm: object = MatricPotentialAlgorithm()


@dataclass(frozen=True)
class MoninObukhov:                              # src/boundary_layer/monin_obukhov.jl:17-20
    karman_constant: float = 0.4
    dyer_constant: float = 16.0


@dataclass(frozen=True)
class PhaseTransitionLatentHeat:                 # src/soil_balance/freezing/latent_heat.jl:9
    pass


@dataclass(frozen=True)
class SoilHeatTransport1D:                       # src/soil_balance/energy/heat_transport_1d.jl:24-29
    reltol: float = 1e-6                         # ode_kwargs.reltol (Tsit5)
    abstol: float = 1e-3                         # ode_kwargs.abstol, K
    maximum_surface_temperature: float = 85.0 + 273.15   # K
    freezing_model: object = PhaseTransitionLatentHeat()


@dataclass(frozen=True)
class BulkTransferEvaporation:                   # src/evaporation/bulk_transfer.jl:9
    pass


@dataclass
class MicroConfig:                               # src/types.jl:71-76
    convergence: object = FixedSoilTemperatureIterations(3)
    rainfall_schedule: object = DailyRainfall()
    soil_moisture_strategy: object = field(default_factory=PrescribedSoilMoisture)
    max_surface_pool: float = 1.0e4              # kg/m2


@dataclass
class MicroModel:                                # src/types.jl:107-120
    soil_properties_model: CampbelldeVriesSoilProperties = CampbelldeVriesSoilProperties()
    soil_hydraulic_model: CampbellSoilHydraulics = CampbellSoilHydraulics()
    depths: np.ndarray = field(default_factory=lambda: DEFAULT_DEPTHS.copy())      # metres
    heights: np.ndarray = field(default_factory=lambda: np.array([0.01, 2.0]))     # metres, last = reference height
    hours: np.ndarray = field(default_factory=lambda: np.arange(24.0))
    radiation: RadiationModel = RadiationModel()
    snow_model: object = NoSnow()
    boundary_layer_model: MoninObukhov = MoninObukhov()
    evaporation_model: BulkTransferEvaporation = BulkTransferEvaporation()
    soil_energy_model: SoilHeatTransport1D = SoilHeatTransport1D()
    config: MicroConfig = field(default_factory=MicroConfig)


# ---- per-site inputs ------------------------------------------------------------------------------
@dataclass
class Site:                                      # src/site/site.jl:11-22 (hot path reads the last 4 + slope)
    latitude: float = 43.07305
    longitude: float = -89.40123
    elevation: float = 226.0
    slope: float = 0.0
    aspect: float = 0.0
    sky_view_fraction: float = 1.0
    albedo: float = 0.15
    roughness_height: float = 0.004
    atmospheric_pressure: Optional[float] = None


@dataclass
class CampbellHydraulicProfile:                  # src/soil_hydraulics/campbell.jl:14-19
    air_entry_water_potential: np.ndarray        # J/kg
    saturated_hydraulic_conductivity: np.ndarray # kg s/m3
    campbell_b_parameter: np.ndarray
    root_density: np.ndarray                     # m/m3


@dataclass
class SoilProfile:                               # src/types.jl:14-20
    bulk_density: np.ndarray                     # kg/m3
    mineral_density: np.ndarray                  # kg/m3
    mineral_conductivity: np.ndarray             # W/m/K
    mineral_heat_capacity: np.ndarray            # J/kg/K
    hydraulics: CampbellHydraulicProfile


@dataclass
class DailyTimeseries:                           # src/inputs/daily_timeseries.jl:1-14
    shade: np.ndarray
    soil_wetness: np.ndarray
    surface_emissivity: np.ndarray
    cloud_emissivity: np.ndarray
    rainfall: np.ndarray                         # kg/m2
    deep_soil_temperature: np.ndarray            # K
    leaf_area_index: np.ndarray
    albedo: Optional[np.ndarray] = None


@dataclass
class HourlyWeather:
    """What `populate_weather!` + `solve_solar!` leave in `cache.output` / `solar_out` for solve_soil!
    to read (src/simulation.jl:1006-1021,1056-1062): hourly vectors of length 24*ndays.
    `rainfall`/`longwave_radiation` are HourlyTimeseries fields (src/inputs/hourly_timeseries.jl:1-11).
    For a min/max-driven problem (MicroInputs.environment_minmax set) only `pressure` (+ rainfall / longwave / an optional global_radiation override)
    is given, as in the reference's HourlyTimeseries with `nothing` fields; solve_minmax_batched fills the rest on the device."""
    pressure: np.ndarray
    reference_temperature: Optional[np.ndarray] = None   # K
    reference_wind_speed: Optional[np.ndarray] = None
    reference_humidity: Optional[np.ndarray] = None      # 0..1
    cloud_cover: Optional[np.ndarray] = None             # 0..1
    global_radiation: Optional[np.ndarray] = None        # W/m2
    zenith_angle: Optional[np.ndarray] = None            # deg (capped at 90)
    diffuse_fraction: Optional[np.ndarray] = None
    zenith_slope_angle: Optional[np.ndarray] = None
    rainfall: Optional[np.ndarray] = None
    longwave_radiation: Optional[np.ndarray] = None


@dataclass
class MicroInputs:                               # src/types.jl:146-157
    site: Site
    soil_profile: SoilProfile
    environment_daily: DailyTimeseries
    environment_hourly: HourlyWeather
    initial_soil_moisture: np.ndarray
    initial_soil_temperature: Optional[np.ndarray] = None   # K; None -> day-mean air temperature
    initial_snow_depth: float = 0.0              # cm
    initial_snow_temperature: float = 273.15
    initial_snow_density: Optional[float] = None
    environment_minmax: Optional[object] = None  # weather.MonthlyMinMaxEnvironment / DailyMinMaxEnvironment (src/types.jl:149); None = hourly-driven


@dataclass
class MicroProblem:                              # src/types.jl:174-183
    model: MicroModel
    inputs: MicroInputs
    days: Sequence[int] = tuple(DEFAULT_DAYS)
    time_mode: object = NonConsecutiveDayMode()


@dataclass
class MicroResult:
    """The solve_soil! slice of MicroResult (src/outputs.jl:18-41): (nsteps, N) matrices / (nsteps,) vectors."""
    soil_temperature: Optional[np.ndarray] = None
    soil_moisture: Optional[np.ndarray] = None
    soil_water_potential: Optional[np.ndarray] = None
    soil_humidity: Optional[np.ndarray] = None
    soil_thermal_conductivity: Optional[np.ndarray] = None
    soil_heat_capacity: Optional[np.ndarray] = None
    soil_bulk_density: Optional[np.ndarray] = None
    surface_water: Optional[np.ndarray] = None
    sky_temperature: Optional[np.ndarray] = None
    snow_fall: Optional[np.ndarray] = None
    snow_depth: Optional[np.ndarray] = None
    snow_density: Optional[np.ndarray] = None
    snow_temperature: Optional[np.ndarray] = None
    # AtmosphericProfile (src/outputs.jl:1-16), filled by solve_air!
    air_temperature: Optional[np.ndarray] = None
    wind_speed: Optional[np.ndarray] = None
    relative_humidity: Optional[np.ndarray] = None
    convective_heat_flux: Optional[np.ndarray] = None
    friction_velocity: Optional[np.ndarray] = None
    status: int = 0


# ---- example_* constructors (src/types.jl:22-50,185-217; src/site/site.jl:29-46; ...) -------------------
def example_site(**kw) -> Site:
    s = Site(**kw)
    if s.atmospheric_pressure is None:
        s.atmospheric_pressure = atmospheric_pressure(s.elevation)
    return s


def example_campbell_hydraulic_profile(depths=DEFAULT_DEPTHS, *, air_entry_water_potential=None,
                                       saturated_hydraulic_conductivity=None, campbell_b_parameter=None,
                                       root_density=None) -> CampbellHydraulicProfile:
    n = len(depths)
    if root_density is None:
        rd = np.array([0, 0, 8.2, 8.0, 7.8, 7.4, 7.1, 6.4, 5.8, 4.8, 4.0, 1.8, 0.9, 0.6, 0.8, 0.4, 0.4, 0, 0]) * 1e4
        root_density = rd if n == 19 else np.zeros(n)
    return CampbellHydraulicProfile(
        air_entry_water_potential=np.full(n, 0.7) if air_entry_water_potential is None else np.asarray(air_entry_water_potential, float),
        saturated_hydraulic_conductivity=np.full(n, 0.0058) if saturated_hydraulic_conductivity is None else np.asarray(saturated_hydraulic_conductivity, float),
        campbell_b_parameter=np.full(n, 1.7) if campbell_b_parameter is None else np.asarray(campbell_b_parameter, float),
        root_density=np.asarray(root_density, float))


def example_soil_profile(depths=DEFAULT_DEPTHS, *, bulk_density=1300.0, mineral_density=2560.0,
                         mineral_conductivity=1.25, mineral_heat_capacity=870.0, hydraulics=None) -> SoilProfile:
    n = len(depths)

    def vec(x):
        return np.asarray(x, float) if np.ndim(x) else np.full(n, float(x))
    return SoilProfile(vec(bulk_density), vec(mineral_density), vec(mineral_conductivity), vec(mineral_heat_capacity),
                       hydraulics if hydraulics is not None else example_campbell_hydraulic_profile(depths))


def example_daily_environment(days=DEFAULT_DAYS, **kw) -> DailyTimeseries:
    n = len(days)
    d = dict(shade=np.zeros(n), soil_wetness=np.zeros(n), surface_emissivity=np.full(n, 0.96),
             cloud_emissivity=np.full(n, 0.96),
             rainfall=np.resize(np.array([28, 28.2, 54.6, 79.7, 81.3, 100.1, 101.3, 102.5, 89.7, 62.4, 54.9, 41.2]), n),
             deep_soil_temperature=np.full(n, 7.741666 + 273.15), leaf_area_index=np.full(n, 0.1), albedo=None)
    d.update(kw)
    return DailyTimeseries(**d)


def example_hourly_environment(days=DEFAULT_DAYS, *, elevation: float = 226.0, **kw) -> HourlyWeather:      # src/inputs/hourly_timeseries.jl:13-29
    """Only the pressure is given; every weather field stays None (`nothing` upstream) and is filled by populate_weather! from the min/max environment."""
    d = dict(pressure=np.full(24 * len(days), atmospheric_pressure(elevation)))
    d.update(kw)
    return HourlyWeather(**d)


def example_soil_properties_model(**kw) -> CampbelldeVriesSoilProperties:                                   # src/soil_properties/campbell_devries.jl:11-19
    return CampbelldeVriesSoilProperties(**kw)


def example_soil_hydraulic_model(**kw) -> CampbellSoilHydraulics:                                           # src/soil_hydraulics/campbell.jl:69-81
    return CampbellSoilHydraulics(**kw)


def example_microclimate_problem(*, days=DEFAULT_DAYS, depths=DEFAULT_DEPTHS, heights=(0.01, 2.0), site=None, soil_profile=None, soil_properties_model=None,
                                 soil_hydraulic_model=None, snow_model=None, environment_minmax=None, environment_daily=None, environment_hourly=None, config=None,
                                 initial_soil_temperature=None, initial_soil_moisture=None) -> MicroProblem:
    """example_microclimate_problem() (src/types.jl:185-210): Madison, Wisconsin, twelve mid-month days driven by monthly min/max weather — solve it with
    `simulation.solve` / `solve_minmax_batched` (the clear-sky irradiance of SolarRadiation.jl is an input there)."""
    from . import weather as W
    site = site if site is not None else example_site()
    depths = np.asarray(depths, float)
    model = MicroModel(depths=depths, heights=np.asarray(heights, float),
                       soil_properties_model=soil_properties_model or example_soil_properties_model(),
                       soil_hydraulic_model=soil_hydraulic_model or example_soil_hydraulic_model(),
                       snow_model=snow_model if snow_model is not None else NoSnow(), config=config or MicroConfig())
    inputs = MicroInputs(site=site, soil_profile=soil_profile or example_soil_profile(depths),
                         environment_minmax=environment_minmax or W.example_monthly_weather(),
                         environment_daily=environment_daily or example_daily_environment(days),
                         environment_hourly=environment_hourly or example_hourly_environment(days, elevation=site.elevation),
                         initial_soil_temperature=np.full(len(depths), 7.741667 + 273.15) if initial_soil_temperature is None else np.asarray(initial_soil_temperature, float),
                         initial_soil_moisture=np.full(len(depths), 0.42 * 0.25) if initial_soil_moisture is None else np.asarray(initial_soil_moisture, float))
    return MicroProblem(model, inputs, days=tuple(days))
