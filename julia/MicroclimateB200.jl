# MicroclimateB200.jl — Julia side of the drop-in: `solve_soil!` for MANY sites on B200s through the C ABI
# (include/microclimate_b200.h).  The host stays Julia, keeps Microclimate.jl's own types, strips units, lays the data out
# site-fastest SoA and `ccall`s libmicroclimate_b200.so.
#
# NOT RUN in the build container (no Julia there); it is the binding INTEGRATION.md describes.  The Python mirror
# (microclimate.jl_b200/marshal.py + simulation.py) performs the same marshalling and is what the tests drive.  The struct
# mirrors below are checked field by field against the C header by tests/test_julia_binding.py, so they cannot rot unnoticed.
module MicroclimateB200

using Microclimate
using Unitful
import Microclimate: MicroModel, MicroInputs, MicroProblem, MicroCache, NoSnow, SnowModel,
    NonConsecutiveDayMode, ConsecutiveDayMode, FixedSoilTemperatureIterations, SoilTemperatureConvergenceTolerance,
    DailyRainfall, HourlyRainfall, PrescribedSoilMoisture, DynamicSoilMoisture, MatricFluxPotentialAlgorithm,
    SwinbankAtmosphericRadiation, BonacinaStep, TanhSmoothed, Gaussian, WestermannSigmoid

const LIB = get(ENV, "MICROCLIMATE_B200_LIB", joinpath(@__DIR__, "..", "microclimate.jl_b200", "csrc", "libmicroclimate_b200.so"))
const MC_MAX_SOIL_NODES = 24
const MC_SNOW_NODES = 8
const MC_MAX_HEIGHTS = 16

# ---- POD mirrors of include/microclimate_b200.h (field order and types must match the header) -----------------
struct McFluid
    molar_mass_water::Cdouble; molar_mass_air::Cdouble; ideal_gas_departure::Cdouble; cp_dry_air::Cdouble
    cp_water_vapour::Cdouble; gas_constant::Cdouble; stefan_boltzmann::Cdouble; standard_gravity::Cdouble
    standard_pressure::Cdouble; latent_heat_a::Cdouble; latent_heat_b::Cdouble; molar_latent_heat_a::Cdouble
    molar_latent_heat_b::Cdouble
end
struct McSnow
    snow_temperature_threshold_C::Cdouble; snow_density::Cdouble; snow_melt_factor::Cdouble; undercatch::Cdouble
    rain_multiplier::Cdouble; rain_melt_factor::Cdouble; density_function::NTuple{4,Cdouble}; snow_conductivity::Cdouble
    canopy_interception::Cdouble; min_snow_depth_cm::Cdouble; node_thresholds_cm::NTuple{MC_SNOW_NODES,Cdouble}
    melt_threshold_C::Cdouble; ahc_kind::Cint
    ahc_T_low_C::Cdouble; ahc_T_high_C::Cdouble; ahc_window_K::Cdouble; ahc_smoothing_K::Cdouble
    ahc_T_melt_C::Cdouble; ahc_width_K::Cdouble
    initial_snow_depth_cm::Cdouble; initial_snow_temperature_K::Cdouble; initial_snow_density::Cdouble
end
struct McModel
    n_soil::Cint; n_snow::Cint
    depths_m::NTuple{MC_MAX_SOIL_NODES,Cdouble}
    reference_height_m::Cdouble; local_height_m::Cdouble
    n_heights::Cint; _pad1::Cint
    heights_m::NTuple{MC_MAX_HEIGHTS,Cdouble}
    reltol::Cdouble; abstol::Cdouble; maximum_surface_temperature_K::Cdouble
    de_vries_shape_factor::Cdouble; recirculation_power::Cdouble; return_flow_threshold::Cdouble
    root_resistance::Cdouble; stomatal_closure_potential::Cdouble; leaf_resistance::Cdouble
    stomatal_stability_parameter::Cdouble; root_radius::Cdouble
    infiltration_algorithm::Cint
    karman_constant::Cdouble; dyer_constant::Cdouble
    time_mode::Cint; iterations_per_day::Cint; spinup_first_day::Cint
    convergence_kind::Cint; convergence_iterations::Cint
    convergence_tolerance_K::Cdouble
    rainfall_schedule::Cint; moisture_strategy::Cint
    moisture_tolerance::Cdouble; moisture_max_iterations::Cint; moisture_timestep_s::Cdouble
    max_surface_pool::Cdouble
    atmospheric_radiation::Cint; _pad0::Cint
    fluid::McFluid; snow::McSnow
end
struct McSites;   p::NTuple{12,Ptr{Cdouble}}; end           # roughness … root_density (header order)
struct McInitial; T::Ptr{Cdouble}; theta::Ptr{Cdouble}; end
struct McForcing; hourly::NTuple{11,Ptr{Cdouble}}; daily::NTuple{8,Ptr{Cdouble}}; prescribed::Ptr{Cdouble}; doy::Ptr{Cint}; end
struct McOutput;  node::NTuple{7,Ptr{Cdouble}}; scalar::NTuple{5,Ptr{Cdouble}}; snow_T::Ptr{Cdouble}
                  profile::NTuple{3,Ptr{Cdouble}}; profile_scalar::NTuple{2,Ptr{Cdouble}}; status::Ptr{Cint}; end   # profile = solve_air! outputs (NULL = skip)

# widening rows f2 / f3: populate_weather! + solar geometry (header section "Widening")
const MC_DIEL_MAX_SHAPES = 6
const MC_DIEL_MAX_INPUTS = 6
struct McTimeOfDay; kind::Cint; _pad::Cint; value::Cdouble; end
struct McDielShape; kind::Cint; _pad::Cint; a::McTimeOfDay; b::McTimeOfDay; end
struct McDielCurve; n_shapes::Cint; n_inputs::Cint; shapes::NTuple{MC_DIEL_MAX_SHAPES,McDielShape}; inputs::NTuple{MC_DIEL_MAX_INPUTS,McTimeOfDay}; end
struct McWeather
    reference_temperature::McDielCurve; reference_wind_speed::McDielCurve; reference_humidity::McDielCurve; cloud_cover::McDielCurve
    angstrom_a::Cdouble; angstrom_b::Cdouble; angstrom_gamma::Cdouble; solar_constant::Cdouble
    diffuse_model::Cint; refraction::Cint; humidity_source::Cint; _pad0::Cint
    orbital_eccentricity::Cdouble; sin_obliquity::Cdouble; days_per_year::Cdouble; equinox_day::Cdouble
end
struct McSolarSites; p::NTuple{4,Ptr{Cdouble}}; end         # latitude, slope, aspect, solar_noon_offset_h
struct McWeatherIn                                          # values[q][i]: quantity q (T, wind, RH, cloud), curve input i
    values::NTuple{24,Ptr{Cdouble}}; daily::NTuple{2,Ptr{Cdouble}}; hourly::NTuple{4,Ptr{Cdouble}}; pressure_site::Ptr{Cdouble}; doy::Ptr{Cint}
end
struct McWeatherOut; hourly::NTuple{9,Ptr{Cdouble}}; daily::NTuple{2,Ptr{Cdouble}}; end

const NULLD = Ptr{Cdouble}(C_NULL)
last_error(h) = unsafe_string(ccall((:mc_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
multi_error(h) = unsafe_string(ccall((:mc_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(h, rc, what; err=last_error) = rc < 0 ? error("$what failed (rc=$rc): " * err(h)) : rc

si(x, u) = Float64(ustrip(u, x))
degC(x) = Float64(ustrip(u"°C", x))

"SnowModel + the inputs' initial snowpack -> mc_snow_desc (src/snow/snow_model.jl:3-54, src/types.jl:146-157)."
function snow_desc(sn::SnowModel, inputs0::MicroInputs)
    length(sn.snow_node_thresholds) == MC_SNOW_NODES || error("SnowModel must have $MC_SNOW_NODES snow nodes")
    a = sn.apparent_heat_capacity
    # defaults of the closures that are not selected (the kernel reads only the fields of ahc_kind)
    kind, Tlo, Thi, win, smo, Tm, wid = 0, -0.45, 0.4, 1.0, 0.1, 0.0, 0.2
    if a isa BonacinaStep
        kind, Tlo, Thi, win = 0, degC(a.T_low), degC(a.T_high), si(a.window, u"K")
    elseif a isa TanhSmoothed
        kind, Tlo, Thi, win, smo = 1, degC(a.T_low), degC(a.T_high), si(a.window, u"K"), si(a.smoothing, u"K")
    elseif a isa Gaussian
        kind, Tm, wid = 2, degC(a.T_melt), si(a.width, u"K")
    elseif a isa WestermannSigmoid
        kind, Tm, wid = 3, degC(a.T_melt), si(a.width, u"K")
    else
        error("unsupported apparent heat capacity closure $(typeof(a))")
    end
    McSnow(degC(sn.snow_temperature_threshold), si(sn.snow_density, u"g/cm^3"), sn.snow_melt_factor, sn.undercatch,
        sn.rain_multiplier, sn.rain_melt_factor, NTuple{4,Cdouble}(sn.density_function), si(sn.snow_conductivity, u"W/m/K"),
        sn.canopy_interception, si(sn.min_snow_depth, u"cm"), ntuple(i -> si(sn.snow_node_thresholds[i], u"cm"), MC_SNOW_NODES),
        degC(sn.melt_threshold), kind, Tlo, Thi, win, smo, Tm, wid,
        si(inputs0.initial_snow_depth, u"cm"), si(u"K"(inputs0.initial_snow_temperature), u"K"),
        isnothing(inputs0.initial_snow_density) ? -1.0 : si(inputs0.initial_snow_density, u"g/cm^3"))
end

"Flatten MicroModel + MicroConfig + time mode into mc_model_desc (reference src/types.jl:71-120)."
function model_desc(model::MicroModel, time_mode, inputs0::MicroInputs)
    d = Ref{McModel}()
    ccall((:mc_default_model, LIB), Cvoid, (Ref{McModel},), d)
    m = d[]
    N = length(model.depths)
    N <= MC_MAX_SOIL_NODES || error("at most $MC_MAX_SOIL_NODES soil nodes")
    depths = ntuple(i -> i <= N ? si(model.depths[i], u"m") : 0.0, MC_MAX_SOIL_NODES)
    sem, sp, sh, cfg = model.soil_energy_model, model.soil_properties_model, model.soil_hydraulic_model, model.config
    tm, ipd, spin = time_mode isa NonConsecutiveDayMode ? (0, time_mode.iterations_per_day, 0) : (1, 1, Int(time_mode.spinup_first_day))
    cv = cfg.convergence
    ck, ci, ct = cv isa FixedSoilTemperatureIterations ? (0, cv.iterations_per_day, 0.0) : (1, cv.max_iterations_per_day, si(cv.tolerance, u"K"))
    sm = cfg.soil_moisture_strategy
    dyn = sm isa DynamicSoilMoisture
    has_snow = model.snow_model isa SnowModel
    swinbank = model.radiation.longwave_model.atmospheric_radiation_model isa SwinbankAtmosphericRadiation
    McModel(N, has_snow ? MC_SNOW_NODES : 0, depths,
        si(last(model.heights), u"m"), si(first(model.heights), u"m"),
        length(model.heights), 0, ntuple(i -> i <= length(model.heights) ? si(model.heights[i], u"m") : 0.0, MC_MAX_HEIGHTS),
        ustrip(sem.ode_kwargs.reltol), si(sem.ode_kwargs.abstol, u"K"), si(u"K"(sem.maximum_surface_temperature), u"K"),
        sp.de_vries_shape_factor, sp.recirculation_power, sp.return_flow_threshold,
        si(sh.root_resistance, u"m^3/kg/s"), si(sh.stomatal_closure_potential, u"J/kg"), si(sh.leaf_resistance, u"m^4/kg/s"),
        sh.stomatal_stability_parameter, si(sh.root_radius, u"m"),
        sh.infiltration_algorithm isa MatricFluxPotentialAlgorithm ? 1 : 0,
        model.boundary_layer_model.karman_constant, model.boundary_layer_model.dyer_constant,
        tm, ipd, spin, ck, ci, ct,
        cfg.rainfall_schedule isa HourlyRainfall ? 1 : 0, dyn ? 1 : 0,
        dyn ? si(sm.moisture_tolerance, u"kg/m^2/s") : 1e-6, dyn ? sm.moisture_max_iterations : 500, dyn ? si(sm.moisture_timestep, u"s") : 360.0,
        si(cfg.max_surface_pool, u"kg/m^2"), swinbank ? 1 : 0, 0, m.fluid,
        has_snow ? snow_desc(model.snow_model, inputs0) : m.snow)
end

"""
    solve_soil_batched!(caches::Vector{<:MicroCache}; device=0, devices=nothing, chunk_days=nothing)

Batched replacement of `Microclimate.solve_soil!(cache)` (src/simulation.jl:355-819): all caches share one `MicroModel`
(the reference's own many-sites idiom, src/types.jl:143-144) and have had `solve_solar!` + `populate_weather!` run, so
`cache.output` holds the hourly weather the path reads (src/simulation.jl:1006-1021,1051-1070).  Fills
`cache.output.{soil_temperature, soil_moisture, soil_water_potential, soil_humidity, soil_thermal_conductivity,
soil_heat_capacity, soil_bulk_density, surface_water, sky_temperature}` (and the snow fields with a `SnowModel`) of every cache,
so `solve_air!` runs unchanged.

`devices = [0, 1, ...]` shards the sites over several GPUs from this one process (`mc_multi_*`); `chunk_days` streams time in
chunks through the asynchronous entry (`mc_step_days_async` / `mc_wait`): the H2D of chunk q+1 and the D2H of chunk q-1 overlap
the kernel of chunk q, and the carried state stays on the device between chunks (identical results to one chunk, except the first
`soil_humidity` row of every chunk after the first: `initialise_soil_humidity!` addresses row 1 of the whole run, which a later chunk does not hold).
The initial snowpack (`initial_snow_*`) is model-level in the ABI: it must be the same for every cache of a batch.
"""
function solve_soil_batched!(caches::AbstractVector{<:MicroCache}; device::Integer=0, devices=nothing, chunk_days=nothing)
    mp = first(caches).problem
    model, S, N, D = mp.model, length(caches), length(mp.model.depths), length(mp.days)
    H = 24D
    inp(c) = c.problem.inputs
    i0 = inp(first(caches))
    has_snow = model.snow_model isa SnowModel
    if has_snow
        all(c -> inp(c).initial_snow_depth == i0.initial_snow_depth && inp(c).initial_snow_temperature == i0.initial_snow_temperature &&
                 inp(c).initial_snow_density == i0.initial_snow_density, caches) || error("initial_snow_* must be identical for all sites of a batch")
    end
    desc = Ref(model_desc(model, mp.time_mode, i0))
    multi = !isnothing(devices)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if multi
        devs = Cint.(collect(devices))
        check(C_NULL, ccall((:mc_multi_create, LIB), Cint, (Ref{McModel}, Int64, Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), desc, S, devs, length(devs), h), "mc_multi_create"; err=multi_error)
    else
        check(C_NULL, ccall((:mc_create, LIB), Cint, (Ref{McModel}, Int64, Cint, Ref{Ptr{Cvoid}}), desc, S, device, h), "mc_create")
    end
    err = multi ? multi_error : last_error
    f_set = multi ? :mc_multi_set_sites : :mc_set_sites
    f_async = multi ? :mc_multi_step_days_async : :mc_step_days_async
    f_wait = multi ? :mc_multi_wait : :mc_wait
    try
        # ---- sites: [S] and [N][S] (site fastest) ------------------------------------------------------------
        sitev(f) = Float64[f(inp(c)) for c in caches]
        nodev(f) = Float64[f(inp(c), i) for c in caches, i in 1:N]            # S×N column-major == [N][S] row-major
        site_arrays = (sitev(x -> si(x.site.roughness_height, u"m")), sitev(x -> si(x.site.slope, u"°")),
                       sitev(x -> x.site.sky_view_fraction), sitev(x -> x.site.albedo),
                       nodev((x, i) -> si(x.soil_profile.bulk_density[i], u"kg/m^3")), nodev((x, i) -> si(x.soil_profile.mineral_density[i], u"kg/m^3")),
                       nodev((x, i) -> si(x.soil_profile.mineral_conductivity[i], u"W/m/K")), nodev((x, i) -> si(x.soil_profile.mineral_heat_capacity[i], u"J/kg/K")),
                       nodev((x, i) -> si(x.soil_profile.hydraulics.air_entry_water_potential[i], u"J/kg")),
                       nodev((x, i) -> si(x.soil_profile.hydraulics.saturated_hydraulic_conductivity[i], u"kg*s/m^3")),
                       nodev((x, i) -> x.soil_profile.hydraulics.campbell_b_parameter[i]), nodev((x, i) -> si(x.soil_profile.hydraulics.root_density[i], u"m/m^3")))
        has_T = !isnothing(i0.initial_soil_temperature)
        T_init = has_T ? nodev((x, i) -> si(u"K"(x.initial_soil_temperature[i]), u"K")) : Float64[]
        th_init = nodev((x, i) -> x.initial_soil_moisture[i])
        # ---- forcing: hourly [H][S], daily [D][S], prescribed moisture [D][N][S] ------------------------------------
        hourly(f) = Float64[f(c, t) for c in caches, t in 1:H]
        daily(f) = Float64[f(inp(c).environment_daily, d) for c in caches, d in 1:D]
        o(c) = c.output
        so(c) = c.buffers.solar_out
        hrain = model.config.rainfall_schedule isa HourlyRainfall
        has_lw = !isnothing(i0.environment_hourly.longwave_radiation)
        has_alb = !isnothing(i0.environment_daily.albedo)
        sms = model.config.soil_moisture_strategy
        has_presc = sms isa PrescribedSoilMoisture && !isnothing(sms.precomputed_soil_moisture)
        hourly_arrays = (hourly((c, t) -> si(o(c).pressure[t], u"Pa")), hourly((c, t) -> si(u"K"(o(c).reference_temperature[t]), u"K")),
                         hourly((c, t) -> si(o(c).reference_wind_speed[t], u"m/s")), hourly((c, t) -> o(c).reference_humidity[t]),
                         hourly((c, t) -> o(c).cloud_cover[t]), hourly((c, t) -> si(o(c).global_radiation[t], u"W/m^2")),
                         hourly((c, t) -> o(c).diffuse_fraction[t]), hourly((c, t) -> min(si(so(c).zenith_angle[t], u"°"), 90.0)),
                         hourly((c, t) -> min(si(so(c).zenith_slope_angle[t], u"°"), 90.0)),
                         hrain ? hourly((c, t) -> si(inp(c).environment_hourly.rainfall[t], u"kg/m^2")) : Float64[],
                         has_lw ? hourly((c, t) -> si(inp(c).environment_hourly.longwave_radiation[t], u"W/m^2")) : Float64[])
        daily_arrays = (daily((e, d) -> ustrip(e.leaf_area_index[d])), daily((e, d) -> e.shade[d]), daily((e, d) -> e.surface_emissivity[d]),
                        daily((e, d) -> e.cloud_emissivity[d]), daily((e, d) -> e.soil_wetness[d]),
                        daily((e, d) -> si(u"K"(e.deep_soil_temperature[d]), u"K")), daily((e, d) -> si(e.rainfall[d], u"kg/m^2")),
                        has_alb ? daily((e, d) -> e.albedo[d]) : Float64[])
        # PrescribedSoilMoisture.precomputed_soil_moisture is one N x D matrix of the MODEL (src/soil_moisture/prescribed.jl:10-14): replicated per site
        presc = has_presc ? Float64[sms.precomputed_soil_moisture[i, d] for c in caches, i in 1:N, d in 1:D] : Float64[]      # S×N×D == [D][N][S]
        doy = Cint.(mp.days)
        # ---- outputs: [H][N][S], [H][S], snow temperature [H][8][S] ------------------------------------------------------
        node_out = ntuple(_ -> zeros(S, N, H), 7)
        scal_out = ntuple(k -> (k <= 2 || has_snow) ? zeros(S, H) : Float64[], 5)      # surface water, sky T, snow fall / depth / density
        snowT_out = has_snow ? zeros(S, MC_SNOW_NODES, H) : Float64[]
        status = zeros(Cint, S)
        ptr(a) = isempty(a) ? NULLD : pointer(a)
        at(a, off) = isempty(a) ? NULLD : pointer(a, off + 1)                          # row offset into a site-fastest array
        cd = isnothing(chunk_days) ? D : min(Int(chunk_days), D)
        GC.@preserve site_arrays T_init th_init hourly_arrays daily_arrays presc doy node_out scal_out snowT_out status begin
            sites = Ref(McSites(map(ptr, site_arrays)))
            init = Ref(McInitial(ptr(T_init), ptr(th_init)))
            check(h[], ccall((f_set, LIB), Cint, (Ptr{Cvoid}, Ref{McSites}, Ref{McInitial}), h[], sites, init), String(f_set); err)
            d0 = 0
            while d0 < D                       # enqueue chunk after chunk; the library keeps two in flight and overlaps their copies with the kernels
                nd = min(cd, D - d0)
                forcing = Ref(McForcing(map(a -> at(a, 24 * d0 * S), hourly_arrays), map(a -> at(a, d0 * S), daily_arrays), at(presc, d0 * N * S), pointer(doy, d0 + 1)))
                out = Ref(McOutput(map(a -> at(a, 24 * d0 * N * S), node_out), map(a -> at(a, 24 * d0 * S), scal_out), at(snowT_out, 24 * d0 * MC_SNOW_NODES * S),
                                   ntuple(_ -> NULLD, 3), ntuple(_ -> NULLD, 2), pointer(status)))   # pass profile buffers here to run solve_air! on the GPU too
                check(h[], ccall((f_async, LIB), Cint, (Ptr{Cvoid}, Int64, Cint, Ref{McForcing}, Ref{McOutput}), h[], d0, nd, forcing, out), String(f_async); err)
                d0 += nd
            end
            check(h[], ccall((f_wait, LIB), Cint, (Ptr{Cvoid},), h[]), String(f_wait); err)
        end
        # ---- un-marshal into each cache's MicroResult (nsteps × N, unit-tagged) so solve_air! runs unchanged ----------
        fields = (:soil_temperature => u"K", :soil_moisture => 1, :soil_water_potential => u"J/kg", :soil_humidity => 1,
                  :soil_thermal_conductivity => u"W/m/K", :soil_heat_capacity => u"J/kg/K", :soil_bulk_density => u"kg/m^3")
        for (s, c) in enumerate(caches), (k, (f, unit)) in enumerate(fields), i in 1:N, t in 1:H
            getproperty(c.output, f)[t, i] = node_out[k][s, i, t] * unit
        end
        for (s, c) in enumerate(caches), t in 1:H
            c.output.surface_water[t] = scal_out[1][s, t] * u"kg/m^2"
            c.output.sky_temperature[t] = scal_out[2][s, t] * u"K"
            if has_snow
                c.output.snow_fall[t] = scal_out[3][s, t] * u"cm/hr"
                c.output.snow_depth[t] = scal_out[4][s, t] * u"cm"
                c.output.snow_density[t] = scal_out[5][s, t] * u"g/cm^3"
                for k in 1:MC_SNOW_NODES
                    c.output.snow_temperature[t, k] = snowT_out[s, k, t] * u"K"
                end
            end
        end
        return status
    finally
        multi ? ccall((:mc_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[]) : ccall((:mc_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

# ---- populate_weather! on the device (widening rows f2 / f3) ---------------------------------------------------------------------
import Microclimate: Sunrise, Sunset, Midday, Midnight, ClockTime, Sine, Decay, Linear, DielCurve, DielForcing, Derived, RelativeHumidityFromVapourPressureAndTemperature,
    AngstromMaxwellShortwave, Angstrom, ErbsDiffuseFraction
import Microclimate: derived_inputs

tod(t::ClockTime) = McTimeOfDay(0, 0, t.hour)
tod(t::Sunrise) = McTimeOfDay(1, 0, t.offset)
tod(t::Sunset) = McTimeOfDay(2, 0, t.offset)
tod(t::Midday) = McTimeOfDay(3, 0, t.offset)
tod(t::Midnight) = McTimeOfDay(4, 0, t.offset)
shape(s::Sine) = McDielShape(0, 0, tod(s.trough), tod(s.peak))
shape(s::Decay) = McDielShape(1, 0, tod(s.from), tod(s.to))
shape(s::Linear) = McDielShape(2, 0, tod(s.from), tod(s.to))
const TOD0 = McTimeOfDay(0, 0, 0.0)
function curve_desc(c::DielCurve)
    length(c.shapes) <= MC_DIEL_MAX_SHAPES && length(c.inputs) <= MC_DIEL_MAX_INPUTS || error("DielCurve does not fit the C ABI (at most 6 shapes and 6 inputs)")
    McDielCurve(length(c.shapes), length(c.inputs),
        ntuple(i -> i <= length(c.shapes) ? shape(c.shapes[i]) : McDielShape(0, 0, TOD0, TOD0), MC_DIEL_MAX_SHAPES),
        ntuple(i -> i <= length(c.inputs) ? tod(c.inputs[i]) : TOD0, MC_DIEL_MAX_INPUTS))
end

"mc_weather_desc from an environment's forcings (DielForcing per quantity) and the shortwave model; geometry constants = the library defaults."
function weather_desc(forcings::NamedTuple, sw::AngstromMaxwellShortwave)
    sw.diffuse_fraction_model isa ErbsDiffuseFraction || error("only ErbsDiffuseFraction exists on the device")
    all(f -> f isa DielForcing, (forcings.reference_temperature, forcings.reference_wind_speed, forcings.cloud_cover)) ||
        error("Derived is supported on the device for reference_humidity only: expand the others with Microclimate.populate_weather!")
    hum, hum_source = humidity_forcing(forcings)
    d = Ref{McWeather}()
    ccall((:mc_default_weather, LIB), Cvoid, (Ref{McWeather},), d)
    w, a = d[], sw.sunshine_fraction_model
    McWeather(curve_desc(forcings.reference_temperature.curve), curve_desc(forcings.reference_wind_speed.curve), curve_desc(hum.curve),
              curve_desc(forcings.cloud_cover.curve), a.a, a.b, a.gamma, w.solar_constant, w.diffuse_model, w.refraction, hum_source, 0,
              w.orbital_eccentricity, w.sin_obliquity, w.days_per_year, w.equinox_day)
end

"The DielForcing that fills the ABI's humidity slot and MC_HUMIDITY_*: the humidity's own curve (0), or the vapour-pressure forcing a Derived humidity reads (1)."
function humidity_forcing(forcings::NamedTuple)
    f = forcings.reference_humidity
    f isa DielForcing && return f, Cint(0)
    if f isa Derived && f.functor isa RelativeHumidityFromVapourPressureAndTemperature
        names = derived_inputs(f)
        length(names) == 2 && names[2] == :reference_temperature && getproperty(forcings, names[1]) isa DielForcing && return getproperty(forcings, names[1]), Cint(1)
    end
    error("reference_humidity must be a DielForcing or Derived(RelativeHumidityFromVapourPressureAndTemperature(GoffGratch()), (<vapour pressure>, :reference_temperature))")
end

"""
    populate_weather_batched!(caches::Vector{<:MicroCache}; device=0, builtin_geometry=false)

Batched replacement of `populate_weather!` (src/simulation.jl:20-36) for caches that share one model and day list: the diel expansion of every
site's min/max series, the Angstrom cloud adjustment and the Erbs diffuse fraction run in one `weather_kernel` launch (`mc_populate_weather`), and the
hourly results are written into each `cache.output` exactly where `populate_weather!` leaves them.  `solve_solar!` must have run (its clear-sky
`direct_horizontal` / `diffuse_horizontal` are inputs; with `builtin_geometry=false` its `hour_solar_noon` / `hour_angle_sunrise` are used too,
otherwise the library's own geometry).  For the fused form that never brings the hourly weather back to the host see `mc_step_days_minmax`.
"""
function populate_weather_batched!(caches::AbstractVector{<:MicroCache}; device::Integer=0, builtin_geometry::Bool=false)
    mp = first(caches).problem
    S, D = length(caches), length(mp.days)
    H = 24D
    envs = [c.problem.inputs.environment_minmax for c in caches]
    w = Ref(weather_desc(first(envs).forcings, mp.model.radiation.shortwave_model))
    so(c) = c.buffers.solar_out
    hum_source = humidity_forcing(first(envs).forcings)[2]
    slot(e, q) = q == :reference_humidity ? humidity_forcing(e.forcings)[1] : getproperty(e.forcings, q)      # the DielForcing behind each ABI slot
    quantities = (:reference_temperature, :reference_wind_speed, :reference_humidity, :cloud_cover)
    units = (u"K", u"m/s", hum_source == 1 ? u"Pa" : 1, 1)
    # values[q][i][d, s] -> [D][S] site-fastest; unit conversion as evaluate! does it (diel.jl:337); the means of minmax_forcings are lazy Broadcasted objects
    num(x, unit) = unit == 1 ? Float64(x) : Float64(ustrip(unit, uconvert(unit, x)))
    series = [[(v = Base.materialize.([slot(e, q).values[i] for e in envs]); Float64[num(v[s][d], unit) for s in 1:S, d in 1:D])
               for i in 1:length(slot(first(envs), q).values)] for (q, unit) in zip(quantities, units)]
    daily(f) = Float64[f(c, d) for c in caches, d in 1:D]
    hourly(f) = Float64[f(c, t) for c in caches, t in 1:H]
    noon = builtin_geometry ? Float64[] : daily((c, d) -> Float64(ustrip(so(c).hour_solar_noon[d])))
    ha = builtin_geometry ? Float64[] : daily((c, d) -> Float64(ustrip(so(c).hour_angle_sunrise[d])))
    direct = hourly((c, t) -> si(so(c).direct_horizontal[t], u"W/m^2"))
    diffuse = hourly((c, t) -> si(so(c).diffuse_horizontal[t], u"W/m^2"))
    eh(c) = c.problem.inputs.environment_hourly
    has_gr = !isnothing(eh(first(caches)).global_radiation)
    gr = has_gr ? hourly((c, t) -> si(eh(c).global_radiation[t], u"W/m^2")) : Float64[]
    pres = hourly((c, t) -> si(eh(c).pressure[t], u"Pa"))
    lat = Float64[si(c.problem.inputs.site.latitude, u"°") for c in caches]
    doy = Cint.(mp.days)
    out = ntuple(_ -> zeros(S, H), 7)                       # pressure, T, wind, RH, cloud, global radiation, diffuse fraction
    ptr(a) = isempty(a) ? NULLD : pointer(a)
    desc = Ref(model_desc(mp.model, mp.time_mode, first(caches).problem.inputs))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:mc_create, LIB), Cint, (Ref{McModel}, Int64, Cint, Ref{Ptr{Cvoid}}), desc, S, device, h), "mc_create")
    try
        GC.@preserve series noon ha direct diffuse gr pres lat doy out begin
            vals = ntuple(k -> (q = (k - 1) ÷ MC_DIEL_MAX_INPUTS + 1; i = (k - 1) % MC_DIEL_MAX_INPUTS + 1; i <= length(series[q]) ? pointer(series[q][i]) : NULLD), 24)
            win = Ref(McWeatherIn(vals, (ptr(noon), ptr(ha)), (pointer(direct), pointer(diffuse), ptr(gr), pointer(pres)), NULLD, pointer(doy)))
            sites = Ref(McSolarSites((pointer(lat), NULLD, NULLD, NULLD)))
            wout = Ref(McWeatherOut((map(pointer, out)..., NULLD, NULLD), (NULLD, NULLD)))
            check(h[], ccall((:mc_populate_weather, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{McWeather}, Ref{McSolarSites}, Ref{McWeatherIn}, Ref{McWeatherOut}),
                             h[], D, w, sites, win, wout), "mc_populate_weather")
        end
        for (s, c) in enumerate(caches), t in 1:H
            c.output.pressure[t] = out[1][s, t] * u"Pa"
            c.output.reference_temperature[t] = out[2][s, t] * u"K"
            c.output.reference_wind_speed[t] = out[3][s, t] * u"m/s"
            c.output.reference_humidity[t] = out[4][s, t]
            c.output.cloud_cover[t] = out[5][s, t]
            c.output.global_radiation[t] = out[6][s, t] * u"W/m^2"
            c.output.diffuse_fraction[t] = out[7][s, t]
        end
    finally
        ccall((:mc_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    return caches
end

end # module
