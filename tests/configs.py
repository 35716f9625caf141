"""The reference's test configurations (BASELINE.json configs 1-3) rebuilt from the committed golden
fixtures, in the reference's own types (test/micro_testrun_monthly.jl:31-153, test/micro_testrun_daily.jl:33-180)."""
import os

import numpy as np

import microclimate_jl_b200 as mc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def config1_monthly(snow=False):
    """Config 1: S=1, N=10, 12 non-consecutive days x 3 passes, prescribed moisture (config 3 with snow=True)."""
    g = load("config1_monthly_snow.npz" if snow else "config1_monthly.npz")
    n = len(g["depths"])
    site = mc.Site(elevation=float(g["elevation"]), slope=float(g["slope"]), sky_view_fraction=1.0,
                   albedo=float(g["albedo"]), roughness_height=float(g["roughness_height"]))
    soil_profile = mc.SoilProfile(
        bulk_density=np.full(n, float(g["bulk_density"])), mineral_density=np.full(n, float(g["mineral_density"])),
        mineral_conductivity=np.full(n, float(g["mineral_conductivity"])),
        mineral_heat_capacity=np.full(n, float(g["mineral_heat_capacity"])),
        hydraulics=mc.example_campbell_hydraulic_profile(g["depths"], root_density=np.zeros(n)))
    nd = len(g["days"])
    daily = mc.DailyTimeseries(shade=g["shade"], soil_wetness=g["soil_wetness"], surface_emissivity=g["surface_emissivity"],
                               cloud_emissivity=g["surface_emissivity"], rainfall=g["rainfall"],
                               deep_soil_temperature=g["deep_soil_temperature"], leaf_area_index=np.full(nd, 0.1))
    hourly = mc.HourlyWeather(pressure=np.full(24 * nd, float(g["pressure"])), reference_temperature=g["TAREF"],
                              reference_wind_speed=g["VREF"], reference_humidity=g["RH"], cloud_cover=g["cloud"],
                              global_radiation=g["SOLR"], zenith_angle=g["ZEN"])
    snow_model = mc.NoSnow()
    if snow:
        snow_model = mc.SnowModel(
            snow_temperature_threshold=float(g["snowtemp"]), snow_density=float(g["snowdens"]), snow_melt_factor=float(g["snowmelt"]),
            undercatch=float(g["undercatch"]), rain_multiplier=float(g["rainmult"]), rain_melt_factor=float(g["rainmelt"]),
            density_function=tuple(float(g[f"densfun{i}"]) for i in (1, 2, 3, 4)), snow_conductivity=float(g["snowcond"]),
            canopy_interception=float(g["intercept"]))
    config = mc.MicroConfig(convergence=mc.FixedSoilTemperatureIterations(10), rainfall_schedule=mc.DailyRainfall(),
                            soil_moisture_strategy=mc.PrescribedSoilMoisture(precomputed_soil_moisture=g["moists"]))
    model = mc.MicroModel(depths=g["depths"], heights=g["heights"], snow_model=snow_model, config=config)
    inputs = mc.MicroInputs(site=site, soil_profile=soil_profile, environment_daily=daily, environment_hourly=hourly,
                            initial_soil_temperature=None, initial_soil_moisture=g["moists"][:, 0])
    problem = mc.MicroProblem(model, inputs, days=[int(d) for d in g["days"]],
                              time_mode=mc.NonConsecutiveDayMode(iterations_per_day=int(g["iterations_per_day"])))
    return problem, g


def config2_daily(ndays=30):
    """Config 2: S=1, N=19, consecutive days, dynamic soil moisture, hourly rainfall."""
    g = load("config2_daily.npz")
    H = 24 * ndays
    site = mc.Site(elevation=float(g["elevation"]), slope=float(g["slope"]), sky_view_fraction=1.0,
                   albedo=float(g["albedo"]), roughness_height=float(g["roughness_height"]))
    n = len(g["depths"])
    soil_profile = mc.SoilProfile(
        bulk_density=g["bulk_density"], mineral_density=g["mineral_density"],
        mineral_conductivity=np.full(n, float(g["mineral_conductivity"])),
        mineral_heat_capacity=np.full(n, float(g["mineral_heat_capacity"])),
        hydraulics=mc.CampbellHydraulicProfile(g["air_entry_water_potential"], g["saturated_hydraulic_conductivity"],
                                               g["campbell_b_parameter"], g["root_density"]))
    daily = mc.DailyTimeseries(shade=g["shade"][:ndays], soil_wetness=g["soil_wetness"][:ndays],
                               surface_emissivity=g["surface_emissivity"][:ndays], cloud_emissivity=g["surface_emissivity"][:ndays],
                               rainfall=g["rainfall_daily"][:ndays], deep_soil_temperature=g["deep_soil_temperature"][:ndays],
                               leaf_area_index=g["leaf_area_index"][:ndays])
    hourly = mc.HourlyWeather(pressure=np.full(H, float(g["pressure"])), reference_temperature=g["TAIRhr"][:H],
                              reference_wind_speed=g["WNhr"][:H], reference_humidity=g["RHhr"][:H], cloud_cover=g["CLDhr"][:H],
                              global_radiation=g["SOLRhr"][:H], zenith_angle=g["ZEN"][:H], rainfall=g["RAINhr"][:H])
    hyd = mc.CampbellSoilHydraulics(root_resistance=float(g["root_resistance"]),
                                    stomatal_closure_potential=float(g["stomatal_closure_potential"]),
                                    leaf_resistance=float(g["leaf_resistance"]),
                                    stomatal_stability_parameter=float(g["stomatal_stability_parameter"]),
                                    root_radius=float(g["root_radius"]))
    config = mc.MicroConfig(convergence=mc.FixedSoilTemperatureIterations(int(g["ndmax"])),
                            rainfall_schedule=mc.HourlyRainfall() if int(g["rainhourly"]) else mc.DailyRainfall(),
                            soil_moisture_strategy=mc.DynamicSoilMoisture(float(g["moisture_tolerance"]),
                                                                          int(g["moisture_max_iterations"]),
                                                                          float(g["moisture_timestep"])),
                            max_surface_pool=float(g["max_surface_pool"]))
    model = mc.MicroModel(depths=g["depths"], heights=g["heights"], soil_hydraulic_model=hyd, config=config)
    inputs = mc.MicroInputs(site=site, soil_profile=soil_profile, environment_daily=daily, environment_hourly=hourly,
                            initial_soil_temperature=g["initial_soil_temperature"], initial_soil_moisture=g["initial_soil_moisture"])
    problem = mc.MicroProblem(model, inputs, days=list(range(1, ndays + 1)),
                              time_mode=mc.ConsecutiveDayMode(spinup_first_day=bool(int(g["spinup"]))))
    return problem, g


def config4_scan329(ndays=365):
    """Config 4 (SCAN / SNOTEL 329, Utah, 2013; test/micro_testrun_snow_scan329.jl:1-238): S=1, N=19 + 8 snow nodes, consecutive days,
    dynamic soil moisture, DailyRainfall, SnowModel with the density function - the only configuration with snow + moisture + a full year."""
    g = load("config4_scan329.npz")
    H, n = 24 * ndays, 19
    site = mc.Site(elevation=float(g["elevation"]), slope=0.0, sky_view_fraction=1.0, albedo=float(g["albedo"]), roughness_height=float(g["roughness_height"]))
    soil_profile = mc.SoilProfile(
        bulk_density=np.full(n, float(g["bulk_density"])), mineral_density=np.full(n, float(g["mineral_density"])),
        mineral_conductivity=g["mineral_conductivity"], mineral_heat_capacity=g["mineral_heat_capacity"],
        hydraulics=mc.CampbellHydraulicProfile(np.full(n, float(g["air_entry_water_potential"])), np.full(n, float(g["saturated_hydraulic_conductivity"])),
                                               np.full(n, float(g["campbell_b_parameter"])), g["root_density"]))
    e = float(g["emissivity"])
    daily = mc.DailyTimeseries(shade=np.zeros(ndays), soil_wetness=np.zeros(ndays), surface_emissivity=np.full(ndays, e), cloud_emissivity=np.full(ndays, e),
                               rainfall=g["rainfall"][:ndays], deep_soil_temperature=g["deep_soil_temperature"][:ndays].astype(float),
                               leaf_area_index=np.full(ndays, 0.1))
    hourly = mc.HourlyWeather(pressure=np.full(H, float(g["pressure"])), reference_temperature=g["TAREF"][:H], reference_wind_speed=g["VREF"][:H],
                              reference_humidity=g["RH"][:H], cloud_cover=g["cloud"][:H], global_radiation=g["SOLR"][:H], zenith_angle=g["ZEN"][:H])
    snow_model = mc.SnowModel(snow_temperature_threshold=float(g["snowtemp"]), snow_density=float(g["snowdens"]), snow_melt_factor=float(g["snowmelt"]),
                              undercatch=float(g["undercatch"]), rain_multiplier=1.0, rain_melt_factor=float(g["rainmelt"]),
                              density_function=tuple(float(x) for x in g["densfun"]), snow_conductivity=float(g["snowcond"]),
                              canopy_interception=float(g["intercept"]))
    config = mc.MicroConfig(convergence=mc.FixedSoilTemperatureIterations(3), rainfall_schedule=mc.DailyRainfall(),
                            soil_moisture_strategy=mc.DynamicSoilMoisture())
    model = mc.MicroModel(depths=g["depths"], heights=g["heights"], snow_model=snow_model, config=config,
                          soil_properties_model=mc.CampbelldeVriesSoilProperties(de_vries_shape_factor=float(g["de_vries_shape_factor"]),
                                                                                 recirculation_power=4.0, return_flow_threshold=0.162))
    inputs = mc.MicroInputs(site=site, soil_profile=soil_profile, environment_daily=daily, environment_hourly=hourly,
                            initial_soil_temperature=g["initial_soil_temperature"], initial_soil_moisture=g["initial_soil_moisture"])
    problem = mc.MicroProblem(model, inputs, days=[int(d) for d in g["days"][:ndays]], time_mode=mc.ConsecutiveDayMode(spinup_first_day=False))
    return problem, g


def regrid(problem, depths):
    """The same problem on another depth grid (per-node soil parameters / initial state taken from the nearest old node)."""
    import copy
    p = copy.deepcopy(problem)
    old, new = np.asarray(p.model.depths, float), np.asarray(depths, float)
    idx = np.abs(old[None, :] - new[:, None]).argmin(axis=1)
    take = lambda a: None if a is None else np.asarray(a)[idx].copy()
    p.model.depths = new
    sp, hy = p.inputs.soil_profile, p.inputs.soil_profile.hydraulics
    for o, names in ((sp, ("bulk_density", "mineral_density", "mineral_conductivity", "mineral_heat_capacity")),
                     (hy, ("air_entry_water_potential", "saturated_hydraulic_conductivity", "campbell_b_parameter", "root_density"))):
        for nme in names:
            object.__setattr__(o, nme, take(getattr(o, nme)))
    p.inputs.initial_soil_temperature = take(p.inputs.initial_soil_temperature)
    p.inputs.initial_soil_moisture = take(p.inputs.initial_soil_moisture)
    return p


def run_engine(engine_cls, problem, inputs_list=None, fields=None, **kw):
    """Run a whole problem in one chunk on `engine_cls` (SoilEngine or OracleEngine). Returns OutputBatch."""
    from microclimate_jl_b200 import marshal as M
    inputs_list = inputs_list or [problem.inputs]
    desc = M.model_desc(problem.model, problem.time_mode, inputs0=inputs_list[0])
    sites = M.SitesBatch.from_inputs(problem.model, inputs_list)
    forcing = M.ForcingBatch.from_inputs(problem.model, inputs_list, problem.days)
    out = M.OutputBatch(len(problem.days), desc.n_soil, len(inputs_list), fields or M.DEFAULT_OUTPUT_FIELDS, desc.n_snow, desc.n_heights)
    eng = engine_cls(desc, len(inputs_list), **kw)
    eng.set_sites(sites)
    eng.step_days(0, forcing, out)
    return out, eng
