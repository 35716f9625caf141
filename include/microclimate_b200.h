/*
 * microclimate_b200.h — C ABI of the B200-native batched-sites soil engine.
 *
 * This is the drop-in boundary for ONE path of BiophysicalEcology/Microclimate.jl:
 *     solve_soil!(cache::MicroCache)            (reference src/simulation.jl:355-819)
 * batched over S independent sites.  The reference has no FFI for this path (it is a
 * plain Julia method); a Julia wrapper (julia/MicroclimateB200.jl) strips units from the
 * reference's own types and `ccall`s the entry points below.  Every entry point cites the
 * reference interface it replaces.
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions across the boundary, no torch types;
 *   - everything is FP64 in unit-stripped SI: K, Pa, m, s, W/m2, kg/m2, J/kg, kg s/m3,
 *     fractions 0..1, angles in degrees;
 *   - structure-of-arrays with the SITE index fastest:  x[node][site], f[hour][site],
 *     out[hour][node][site]  (reference: one Matrix(nsteps, N) per site, time fastest,
 *     src/outputs.jl:52-58);
 *   - the caller owns every host array for the duration of the call; the library owns
 *     device memory; only the handle needs freeing;
 *   - return value: 0 ok, <0 error (see mc_last_error), >0 number of sites whose status
 *     word is non-zero (numeric flag).
 *   - a handle is confined to one host thread at a time and drives ONE GPU; multi-GPU
 *     runs create one handle per GPU over disjoint, contiguous site ranges (no collective),
 *     either one process per GPU or one process with a multi handle (mc_multi_*).
 *
 * Widened beyond solve_soil!, each selected / called separately: solve_air! (the
 * profile members of mc_output_soa), populate_weather! + solar geometry (section "Widening":
 * mc_populate_weather, mc_step_days_minmax, mc_multi_step_days_minmax).
 */
#ifndef MICROCLIMATE_B200_H
#define MICROCLIMATE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bumped on every layout change of a struct below (2: mc_state_soa.status, the weather structs); bindings compare it and mc_abi_sizeof() with their mirrors */
#define MC_ABI_VERSION 2
#define MC_MAX_SOIL_NODES 24
#define MC_SNOW_NODES 8
#define MC_MAX_ODE_NODES (MC_MAX_SOIL_NODES + MC_SNOW_NODES)
#define MC_HOURS_PER_DAY 24
#define MC_MAX_HEIGHTS 16

/* enums mirror the reference's strategy singletons */
enum { MC_TIME_NONCONSECUTIVE = 0, MC_TIME_CONSECUTIVE = 1 };   /* src/time_mode/{non_consecutive,consecutive}.jl */
enum { MC_CONV_FIXED = 0, MC_CONV_TOLERANCE = 1 };              /* src/soil_temperature_convergence/{fixed,tolerance}.jl */
enum { MC_RAIN_DAILY = 0, MC_RAIN_HOURLY = 1 };                 /* src/rainfall/{daily,hourly}_rainfall.jl */
enum { MC_MOIST_PRESCRIBED = 0, MC_MOIST_DYNAMIC = 1 };         /* src/soil_moisture/{prescribed,dynamic}.jl */
enum { MC_INFIL_MATRIC_POTENTIAL = 0, MC_INFIL_MATRIC_FLUX_POTENTIAL = 1 }; /* src/soil_hydraulics/infiltration_algorithm */
enum { MC_ATMRAD_CAMPBELL_NORMAN = 0, MC_ATMRAD_SWINBANK = 1 }; /* src/radiation/atmospheric */
enum { MC_AHC_BONACINA = 0, MC_AHC_TANH = 1, MC_AHC_GAUSSIAN = 2, MC_AHC_WESTERMANN = 3 }; /* src/apparent_heat_capacity */

/* per-site status bits (mc_output_soa.status / mc_get_status) */
enum {
    MC_STATUS_NAN            = 1,   /* a NaN/Inf reached the soil temperature or moisture state */
    MC_STATUS_DT_UNDERFLOW   = 2,   /* adaptive step fell below 1e-12 min */
    MC_STATUS_NEWTON_MAXITER = 4,   /* infiltration Newton loop hit moisture_max_iterations */
    MC_STATUS_STEP_LIMIT     = 8    /* more than MC_MAX_STEPS_PER_HOUR RK attempts in one hour */
};

/*
 * Psychrometric constants.  The arithmetic lives in the un-vendored, un-pinned
 * FluidProperties.jl (reference Project.toml:29,43) — kept as data so a maintainer with
 * Julia can set them to whatever that package uses.  Defaults: mc_default_fluid().
 */
typedef struct mc_fluid_constants {
    double molar_mass_water;        /* kg/mol, 0.018016 */
    double molar_mass_air;          /* kg/mol, 0.028966 */
    double ideal_gas_departure;     /* 1.0053 */
    double cp_dry_air;              /* J/kg/K, 1004.84 */
    double cp_water_vapour;         /* J/kg/K, 1846.40 */
    double gas_constant;            /* J/mol/K, 8.314462618 */
    double stefan_boltzmann;        /* W/m2/K4, 5.670374419e-8 */
    double standard_gravity;        /* m/s2, 9.80665 (Unitful.gn) */
    double standard_pressure;       /* Pa, 101325 (Unitful.atm) */
    double latent_heat_a;           /* enthalpy_of_vaporisation(T) = a - b*(T-273.15)  J/kg; a=2.5012e6 */
    double latent_heat_b;           /* b=2.3787e3 */
    double molar_latent_heat_a;     /* molar_enthalpy_of_vaporisation(T) = a - b*(T-273.15) J/mol; a=45144 */
    double molar_latent_heat_b;     /* b=48 */
} mc_fluid_constants;

/* SnowModel (src/snow/snow_model.jl:3-42) */
typedef struct mc_snow_desc {
    double snow_temperature_threshold_C;   /* degC */
    double snow_density;                   /* g/cm3 */
    double snow_melt_factor;
    double undercatch;
    double rain_multiplier;
    double rain_melt_factor;
    double density_function[4];
    double snow_conductivity;              /* W/m/K; 0 = Aggarwal cubic */
    double canopy_interception;
    double min_snow_depth_cm;
    double node_thresholds_cm[MC_SNOW_NODES];
    double melt_threshold_C;
    int32_t ahc_kind;                      /* MC_AHC_* */
    double ahc_T_low_C, ahc_T_high_C, ahc_window_K, ahc_smoothing_K; /* Bonacina / Tanh */
    double ahc_T_melt_C, ahc_width_K;                                 /* Gaussian / Westermann */
    double initial_snow_depth_cm;          /* MicroInputs.initial_snow_depth (model-level here) */
    double initial_snow_temperature_K;
    double initial_snow_density;           /* <=0: use default */
} mc_snow_desc;

/*
 * mc_model_desc — flattened MicroModel + MicroConfig + time mode
 * (reference src/types.jl:71-120,174-183; strategy objects in src/time_mode, src/soil_moisture,
 *  src/soil_temperature_convergence, src/rainfall; solver knobs src/soil_balance/energy/heat_transport_1d.jl:24-29).
 * Constant across sites, exactly as MicroModel is constant across `reinit!`ed inputs (src/types.jl:143-144).
 */
typedef struct mc_model_desc {
    int32_t n_soil;                         /* N = length(depths) */
    int32_t n_snow;                         /* 0 (NoSnow) or 8 (SnowModel) */
    double  depths_m[MC_MAX_SOIL_NODES];    /* MicroModel.depths, metres */
    double  reference_height_m;             /* last(heights) */
    double  local_height_m;                 /* first(heights) */
    int32_t n_heights;                      /* length(MicroModel.heights): air nodes of solve_air! (src/simulation.jl:822-851) */
    int32_t _pad1;
    double  heights_m[MC_MAX_HEIGHTS];      /* MicroModel.heights, metres; the last one is the reference height */
    /* SoilHeatTransport1D */
    double  reltol, abstol;                 /* 1e-6, 1e-3 (K) */
    double  maximum_surface_temperature_K;  /* 85 degC = 358.15 */
    /* CampbelldeVriesSoilProperties (scalar form) */
    double  de_vries_shape_factor, recirculation_power, return_flow_threshold;
    /* CampbellSoilHydraulics plant scalars */
    double  root_resistance, stomatal_closure_potential, leaf_resistance,
            stomatal_stability_parameter, root_radius;
    int32_t infiltration_algorithm;         /* MC_INFIL_* */
    /* MoninObukhov */
    double  karman_constant, dyer_constant;
    /* time mode */
    int32_t time_mode;                      /* MC_TIME_* */
    int32_t iterations_per_day;             /* NonConsecutiveDayMode.iterations_per_day */
    int32_t spinup_first_day;               /* ConsecutiveDayMode.spinup_first_day */
    /* MicroConfig */
    int32_t convergence_kind;               /* MC_CONV_* */
    int32_t convergence_iterations;         /* iterations_per_day / max_iterations_per_day */
    double  convergence_tolerance_K;
    int32_t rainfall_schedule;              /* MC_RAIN_* */
    int32_t moisture_strategy;              /* MC_MOIST_* */
    double  moisture_tolerance;             /* kg/m2/s */
    int32_t moisture_max_iterations;
    double  moisture_timestep_s;
    double  max_surface_pool;               /* kg/m2 */
    int32_t atmospheric_radiation;          /* MC_ATMRAD_* */
    int32_t _pad0;
    mc_fluid_constants fluid;
    mc_snow_desc snow;
} mc_model_desc;

/*
 * mc_sites_soa — Site + SoilProfile + CampbellHydraulicProfile for S sites
 * (reference src/site/site.jl:11-22 — only the four fields the hot path reads,
 *  src/types.jl:14-20, src/soil_hydraulics/campbell.jl:14-19).
 */
typedef struct mc_sites_soa {
    const double* roughness_height;    /* [S] m */
    const double* slope;               /* [S] degrees */
    const double* sky_view_fraction;   /* [S] */
    const double* albedo;              /* [S] Site.albedo */
    const double* bulk_density;        /* [N][S] kg/m3 */
    const double* mineral_density;     /* [N][S] kg/m3 */
    const double* mineral_conductivity;/* [N][S] W/m/K */
    const double* mineral_heat_capacity;/*[N][S] J/kg/K */
    const double* air_entry_water_potential;        /* [N][S] J/kg */
    const double* saturated_hydraulic_conductivity; /* [N][S] kg s/m3 */
    const double* campbell_b_parameter;             /* [N][S] */
    const double* root_density;                     /* [N][S] m/m3 */
} mc_sites_soa;

/*
 * mc_initial_soa — MicroInputs initial conditions (reference src/types.jl:146-157).
 * initial_soil_temperature == NULL reproduces `initial_soil_temperature=nothing`
 * (TINS mean of the first day's air temperature, src/simulation.jl:395-397).
 */
typedef struct mc_initial_soa {
    const double* initial_soil_temperature; /* [N][S] K, or NULL */
    const double* initial_soil_moisture;    /* [N][S] m3/m3 */
} mc_initial_soa;

/*
 * mc_state_soa — everything solve_soil! carries from one hour/day to the next
 * (locals at src/simulation.jl:369-371,451-452 + the mutable strategy/buffer slots the
 * reference hides: DynamicSoilMoisture.soil_wetness, the two obukhov_length_prev Refs,
 * the integrator's stale FSAL slot).  Exposed so runs can be time-chunked / checkpointed.
 * All pointers are caller-owned host arrays.
 */
typedef struct mc_state_soa {
    double* soil_temperature;   /* [N][S] K */
    double* soil_moisture;      /* [N][S] */
    double* sum_phase;          /* [N][S] J/m2 (accumulated latent heat) */
    double* pool;               /* [S] kg/m2 */
    double* soil_wetness;       /* [S] */
    double* obukhov_ode;        /* [S] m   buffers.soil_energy_balance.obukhov_length_prev */
    double* obukhov_profile;    /* [S] m   buffers.soil_water_profile.obukhov_length_prev */
    double* deep_temperature_prev; /* [S] K  environment_instant.deep_soil_temperature left over from the previous day */
    double* fsal;               /* [M][S] K/min  integrator.fsalfirst left over from the previous day-pass */
    double* snow_temperature;   /* [8][S] K (n_snow>0) or NULL */
    double* snow_state;         /* [MC_SNOW_STATE_FIELDS][S] (n_snow>0) or NULL */
    double* snow_node_depths;   /* [8][S] cm (n_snow>0) or NULL */
    double* q_freeze;           /* [S] W/m2 (n_snow>0) or NULL */
    double* obukhov_air;        /* [S] m   buffers.air_profile.obukhov_length_prev (solve_air!) */
    int64_t* days_done;         /* [1] days integrated so far (global day index of the next day) */
    int32_t* status;            /* [S] MC_STATUS_* bits accumulated so far (a resumed run keeps its flagged sites), or NULL */
} mc_state_soa;
#define MC_SNOW_STATE_FIELDS 10 /* depth, age, days_since_snow, density, prev_density, cum_melt, extra_snow, active_nodes, sum_phase, prev_hour_depth */

/*
 * mc_forcing_soa — what solve_soil! reads from cache.output / solar_out / environment_daily
 * for a chunk of `ndays` days (reference src/simulation.jl:1006-1021,1025-1070):
 * hourly arrays are [24*ndays][S], daily arrays [ndays][S].
 */
typedef struct mc_forcing_soa {
    const double* pressure;              /* Pa */
    const double* reference_temperature; /* K */
    const double* reference_wind_speed;  /* m/s */
    const double* reference_humidity;    /* 0..1 */
    const double* cloud_cover;           /* 0..1 */
    const double* global_radiation;      /* W/m2 (already cloud adjusted) */
    const double* diffuse_fraction;      /* 0..1 */
    const double* zenith_angle;          /* deg, capped at 90 */
    const double* zenith_slope_angle;    /* deg, capped at 90 */
    const double* rainfall_hourly;       /* kg/m2 per hour; NULL unless MC_RAIN_HOURLY */
    const double* longwave_radiation;    /* W/m2 user supplied downward longwave, or NULL */
    /* daily */
    const double* leaf_area_index;
    const double* shade;
    const double* surface_emissivity;
    const double* cloud_emissivity;
    const double* soil_wetness;
    const double* deep_soil_temperature; /* K */
    const double* rainfall_daily;        /* kg/m2 per day */
    const double* albedo_daily;          /* or NULL -> Site.albedo */
    const double* prescribed_soil_moisture; /* [ndays][N][S] or NULL (PrescribedSoilMoisture.precomputed_soil_moisture) */
    const int32_t* day_of_year;          /* [ndays] MicroProblem.days */
} mc_forcing_soa;

/*
 * mc_output_soa — the MicroResult fields solve_soil! writes (reference src/outputs.jl:18-41,
 * writes at src/simulation.jl:423-444,562-587,788-797).  A NULL pointer deselects a field.
 * Node fields are [24*ndays][N][S]; scalars [24*ndays][S].  Row convention is the
 * reference's: hour i of a day is written to row i+1, row 1 of each day is the state at
 * the start of the output pass, the last hour of the last day is not written.
 */
typedef struct mc_output_soa {
    double* soil_temperature;          /* K */
    double* soil_moisture;
    double* soil_water_potential;      /* J/kg */
    double* soil_humidity;
    double* soil_thermal_conductivity; /* W/m/K */
    double* soil_heat_capacity;        /* J/kg/K */
    double* soil_bulk_density;         /* kg/m3 */
    double* surface_water;             /* [H][S] kg/m2 */
    double* sky_temperature;           /* [H][S] K */
    double* snow_fall;                 /* [H][S] cm/hr */
    double* snow_depth;                /* [H][S] cm */
    double* snow_density;              /* [H][S] g/cm3 */
    double* snow_temperature;          /* [H][8][S] K */
    /* solve_air! (src/simulation.jl:822-851): MOST profiles from the final surface temperature; needs soil_temperature (and, with
     * snow, snow_depth + snow_temperature) selected as well.  NULL = do not run solve_air! */
    double* air_temperature;           /* [H][n_heights][S] K */
    double* wind_speed;                /* [H][n_heights][S] m/s */
    double* relative_humidity;         /* [H][n_heights][S] */
    double* convective_heat_flux;      /* [H][S] W/m2 */
    double* friction_velocity;         /* [H][S] m/s */
    int32_t* status;                   /* [S] MC_STATUS_* bits, or NULL */
} mc_output_soa;

/* device-side work counters, summed over sites (SURVEY §8d definition of algorithmic work) */
typedef struct mc_counters {
    uint64_t n_rhs;            /* soil_energy_balance evaluations */
    uint64_t n_accept;         /* accepted Tsit5 steps */
    uint64_t n_reject;         /* rejected Tsit5 steps */
    uint64_t n_obukhov_iter;   /* calc_Obukhov_length iterations (ODE + hourly warm start + moisture profile) */
    uint64_t n_newton_iter;    /* infiltration Newton iterations */
    uint64_t n_infil_steps;    /* infiltration_step! calls */
    uint64_t n_site_days;      /* site-days integrated (all passes of a day count once) */
    uint64_t n_day_passes;     /* site-day-passes integrated */
    double   gpu_ms;           /* device time of the last mc_step_days* call (CUDA events) */
    uint64_t n_kernel_launches;/* kernels launched by the last mc_step_days* call */
} mc_counters;

typedef struct mc_handle mc_handle;

/* library identity */
int         mc_abi_version(void);
const char* mc_backend_name(void);              /* "cuda-sm_100a" */
void        mc_default_fluid(mc_fluid_constants* out);
void        mc_default_model(mc_model_desc* out); /* MicroModel()/MicroConfig() defaults, src/types.jl:71-120 */

/* replaces CommonSolve.init(::MicroProblem) for the soil path (src/simulation.jl:141-243):
 * allocates device state for n_sites sites on CUDA device `device`. */
int  mc_create(const mc_model_desc* model, int64_t n_sites, int device, mc_handle** out);
void mc_destroy(mc_handle* h);
const char* mc_last_error(const mc_handle* h);  /* h may be NULL: last mc_create error */

/* replaces reinit!(cache, inputs) (src/simulation.jl:274-277): new sites, fresh state. */
int  mc_set_sites(mc_handle* h, const mc_sites_soa* sites, const mc_initial_soa* init);

/* checkpoint / resume of the carried state. */
int  mc_get_state(mc_handle* h, mc_state_soa* out);
int  mc_set_state(mc_handle* h, const mc_state_soa* in);

/* replaces solve_soil!(cache) (src/simulation.jl:355-819) for days [day0, day0+ndays) of the
 * run; host buffers in/out, H2D and D2H included.  day0 must equal the number of days
 * already integrated (0 after mc_set_sites). */
int  mc_step_days(mc_handle* h, int64_t day0, int32_t ndays,
                  const mc_forcing_soa* forcing, mc_output_soa* out);

/* Asynchronous form of mc_step_days: enqueues H2D(forcing) -> kernel -> D2H(outputs, status) on the handle's copy-in / compute / copy-out
 * streams and returns.  Two staging sets alternate, so with the calls for chunk q+1 issued before mc_wait the H2D of chunk q+1 and the D2H of
 * chunk q-1 overlap the kernel of chunk q (a third call waits for the oldest chunk).  Host buffers (pinned for real overlap) must stay valid
 * until mc_wait, which returns the flagged-site count of the most recent chunk.  This is what a host loop over day chunks - and a single host
 * thread driving several devices - uses; the reference's equivalent is the serial day loop of solve_soil! (src/simulation.jl:454-819). */
int  mc_step_days_async(mc_handle* h, int64_t day0, int32_t ndays,
                        const mc_forcing_soa* forcing, mc_output_soa* out);
int  mc_wait(mc_handle* h);

/* same, but every pointer inside `forcing` and `out` is a DEVICE pointer on the handle's
 * GPU (day_of_year stays a host pointer); no copies.  Asynchronous on the handle's stream
 * until mc_synchronize. */
int  mc_step_days_device(mc_handle* h, int64_t day0, int32_t ndays,
                         const mc_forcing_soa* forcing_dev, mc_output_soa* out_dev);
int  mc_synchronize(mc_handle* h);

int  mc_get_counters(mc_handle* h, mc_counters* out);
int  mc_reset_counters(mc_handle* h);
int  mc_get_status(mc_handle* h, int32_t* status_host /* [S] */);

/* One host process, several GPUs (the reference idiom is one process sweeping sites with reinit!(cache, inputs); solve!(cache),
 * src/simulation.jl:266-277).  A multi handle shards the n_sites columns of ONE set of host arrays contiguously over the listed devices (a device
 * may be listed more than once); mc_multi_step_days* runs the asynchronous pipeline of every device on its column range, so all GPUs copy and
 * integrate concurrently and write into disjoint column ranges of the caller's output arrays.  Same call order and error behaviour as the
 * single-device entry points; mc_multi_wait returns the flagged-site count summed over devices. */
typedef struct mc_multi mc_multi;
int  mc_multi_create(const mc_model_desc* model, int64_t n_sites, const int* device_ids, int n_devices, mc_multi** out);
void mc_multi_destroy(mc_multi* m);
const char* mc_multi_last_error(const mc_multi* m);
int  mc_multi_device_count(const mc_multi* m);
mc_handle* mc_multi_handle(mc_multi* m, int g);                                  /* per-device handle (state checkpoints, counters, status) */
int  mc_multi_range(const mc_multi* m, int g, int64_t* lo, int64_t* hi);         /* its site range [lo, hi) */
int  mc_multi_set_sites(mc_multi* m, const mc_sites_soa* sites, const mc_initial_soa* init);
int  mc_multi_step_days(mc_multi* m, int64_t day0, int32_t ndays, const mc_forcing_soa* forcing, mc_output_soa* out);
int  mc_multi_step_days_async(mc_multi* m, int64_t day0, int32_t ndays, const mc_forcing_soa* forcing, mc_output_soa* out);
int  mc_multi_wait(mc_multi* m);

/* ------------------------------------------------------------------------------------------------------------------------------
 * Widening: populate_weather! and the solar geometry it needs, on the device.
 *
 * Replaces, for S sites at once, what solve! does between init and solve_soil! (reference src/simulation.jl:20-36, 291-352):
 *     solve_solar!  -> zenith / slope-zenith angle per hour, solar noon and sunrise hour angle per day   (geometry only, see below)
 *     interpolate_minmax! -> the declarative diel engine of src/inputs/diel.jl (DielCurve of Sine / Decay / Linear shapes over
 *                            Sunrise / Sunset / Midday / Midnight / ClockTime anchors, seam heights handed off between shapes)
 *     adjust_for_cloud_cover! -> AngstromMaxwellShortwave + ErbsDiffuseFraction (src/radiation/shortwave/angstrom_maxwell.jl:21-46,
 *                            src/radiation/diffuse/erbs.jl:14-23, src/radiation/sunshine_fraction/angstrom.jl:8-14)
 * The hourly arrays it writes are exactly the hourly members of mc_forcing_soa, in the same [24*ndays][S] layout, so a min/max-driven
 * run ships ~60 doubles per site-day to the GPU instead of 223 (mc_step_days_minmax).
 *
 * Solar geometry lives in the un-vendored SolarRadiation.jl (reference Project.toml; call at src/simulation.jl:291-304).  The geometry
 * restated here is the published McCullough & Porter (1971) form NicheMapR's SOLRAD uses, anchored on the reference's golden ZEN columns
 * (tests/test_weather.py: three data sets, <= 3e-6 degrees).  The spectral clear-sky model (111 wavelengths with ozone / aerosol /
 * water-vapour tables) cannot be restated without its source: clear-sky direct and diffuse horizontal irradiance stay CALLER-SUPPLIED
 * hourly arrays (SolarRadiation.solar_radiation!'s `direct_horizontal`, `diffuse_horizontal`).
 * ---------------------------------------------------------------------------------------------------------------------------- */

/* TimeOfDay (src/inputs/diel.jl:19-85): value = offset in hours for the solar events and Midnight, the clock hour for ClockTime */
enum { MC_TOD_CLOCK = 0, MC_TOD_SUNRISE = 1, MC_TOD_SUNSET = 2, MC_TOD_MIDDAY = 3, MC_TOD_MIDNIGHT = 4 };
typedef struct mc_time_of_day { int32_t kind; int32_t _pad; double value; } mc_time_of_day;
/* Shape (src/inputs/diel.jl:98-150): Sine(trough, peak) fills whatever the bounded shapes leave; Decay / Linear(from, to) cover [from, to) */
enum { MC_SHAPE_SINE = 0, MC_SHAPE_DECAY = 1, MC_SHAPE_LINEAR = 2 };
typedef struct mc_diel_shape { int32_t kind; int32_t _pad; mc_time_of_day a, b; } mc_diel_shape;
#define MC_DIEL_MAX_SHAPES 6
#define MC_DIEL_MAX_INPUTS 6
/* DielCurve (src/inputs/diel.jl:152-180): shapes + the times whose heights are fed from outside, in value-series order */
typedef struct mc_diel_curve {
    int32_t n_shapes, n_inputs;
    mc_diel_shape shapes[MC_DIEL_MAX_SHAPES];
    mc_time_of_day inputs[MC_DIEL_MAX_INPUTS];
} mc_diel_curve;

enum { MC_DIFFUSE_ERBS = 0 };                                    /* src/radiation/diffuse/erbs.jl */
/* where the hourly relative humidity comes from: its own diel curve, or Derived(RelativeHumidityFromVapourPressureAndTemperature(GoffGratch()),
 * (:actual_vapour_pressure, :reference_temperature)) (src/inputs/diel.jl:264-301, fill_derived! :438-444): the `reference_humidity` curve and value series then
 * describe the ACTUAL VAPOUR PRESSURE in Pa and humidity = clamp(e / e_sat(T), 0, 1) hour by hour */
enum { MC_HUMIDITY_CURVE = 0, MC_HUMIDITY_FROM_VAPOUR_PRESSURE = 1 };
/* forcing model + shortwave model + geometry constants; mc_default_weather() = minmax_forcing_model() with its default offsets
 * (src/inputs/diel.jl:456-495), AngstromMaxwellShortwave(), Angstrom(0.36, 0.64, 1.0) */
typedef struct mc_weather_desc {
    mc_diel_curve reference_temperature, reference_wind_speed, reference_humidity, cloud_cover;
    double  angstrom_a, angstrom_b, angstrom_gamma;
    double  solar_constant;                 /* W/m2, 1367 (angstrom_maxwell.jl:27) */
    int32_t diffuse_model;                  /* MC_DIFFUSE_* */
    int32_t refraction;                     /* 1: apparent zenith within 2 degrees of the horizon (NicheMapR SOLRAD; goldens need it) */
    int32_t humidity_source;                /* MC_HUMIDITY_* */
    int32_t _pad0;
    double  orbital_eccentricity;           /* 0.0167238 */
    double  sin_obliquity;                  /* 0.39784993 */
    double  days_per_year;                  /* 365 */
    double  equinox_day;                    /* 80 */
} mc_weather_desc;

/* the Site fields SolarTerrain(site) passes to the geometry (src/site/site.jl:48-60) */
typedef struct mc_solar_sites_soa {
    const double* latitude;                 /* [S] degrees, north positive */
    const double* slope;                    /* [S] degrees, or NULL = 0 */
    const double* aspect;                   /* [S] degrees clockwise from north of the downslope direction, or NULL = 0 */
    const double* solar_noon_offset_h;      /* [S] hours added to 12:00 for local solar noon (longitude / time-zone correction), or NULL = 0 */
} mc_solar_sites_soa;

/* per-day value series of every curve input ([ndays][S]; values[i] belongs to curve.inputs[i]) + what populate_weather! takes from solve_solar! */
typedef struct mc_weather_in_soa {
    const double* reference_temperature_values[MC_DIEL_MAX_INPUTS];   /* K */
    const double* reference_wind_speed_values[MC_DIEL_MAX_INPUTS];    /* m/s */
    const double* reference_humidity_values[MC_DIEL_MAX_INPUTS];      /* 0..1 */
    const double* cloud_cover_values[MC_DIEL_MAX_INPUTS];             /* 0..1 */
    const double* hour_solar_noon;          /* [ndays][S] h, or NULL: built-in geometry (needs mc_solar_sites_soa) */
    const double* hour_angle_sunrise;       /* [ndays][S] h from noon back to sunrise, or NULL: built-in geometry */
    const double* direct_horizontal;        /* [24*ndays][S] W/m2 clear-sky direct on the horizontal (SolarRadiation.jl) */
    const double* diffuse_horizontal;       /* [24*ndays][S] W/m2 clear-sky diffuse */
    const double* global_radiation_hourly;  /* [24*ndays][S] W/m2 environment_hourly.global_radiation override, or NULL */
    const double* pressure_hourly;          /* [24*ndays][S] Pa environment_hourly.pressure, or NULL: pressure_site repeated */
    const double* pressure_site;            /* [S] Pa */
    const int32_t* day_of_year;             /* [ndays] MicroProblem.days — always a HOST pointer */
} mc_weather_in_soa;

/* what populate_weather! / solve_solar! leave in cache.output and solar_out; NULL = not wanted.  Hourly [24*ndays][S], daily [ndays][S]. */
typedef struct mc_weather_out_soa {
    double* pressure;
    double* reference_temperature;
    double* reference_wind_speed;
    double* reference_humidity;
    double* cloud_cover;
    double* global_radiation;
    double* diffuse_fraction;
    double* zenith_angle;                   /* capped at 90 (src/simulation.jl:301) */
    double* zenith_slope_angle;             /* capped at 90 (src/simulation.jl:302) */
    double* hour_solar_noon;                /* [ndays][S] */
    double* hour_angle_sunrise;             /* [ndays][S] */
} mc_weather_out_soa;

void mc_default_weather(mc_weather_desc* out);
/* DielCurve constructor checks (src/inputs/diel.jl:182-222: overlap, gaps without a space-filling shape, inputs that are no endpoint)
 * plus the ABI's own limits; 0 ok, <0 with a message in `err` (may be NULL) */
int  mc_validate_weather(const mc_weather_desc* w, char* err, int err_len);

/* replaces solve_solar! (geometry part) + populate_weather! (src/simulation.jl:20-36,291-352) for S sites: host arrays in, host arrays out */
int  mc_populate_weather(mc_handle* h, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                         const mc_weather_in_soa* in, mc_weather_out_soa* out);
/* same with DEVICE pointers everywhere (day_of_year stays a host pointer); asynchronous on the handle's stream until mc_synchronize */
int  mc_populate_weather_device(mc_handle* h, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites_dev,
                                const mc_weather_in_soa* in_dev, mc_weather_out_soa* out_dev);
/* solve!(cache) for a min/max-driven problem (src/simulation.jl:246-260 minus solve_air!'s own outputs, which `out` may select as usual):
 * weather expanded on the device straight into the soil kernel's forcing arrays, then solve_soil! for days [day0, day0+ndays).
 * `rest` supplies what populate_weather! does not produce — the daily members, rainfall_hourly, longwave_radiation, prescribed_soil_moisture,
 * day_of_year — and MAY supply zenith_angle + zenith_slope_angle (SolarRadiation.jl's own; NULL = built-in geometry); its other hourly members
 * are ignored.  `weather_host` (may be NULL) receives the expanded hourly weather. */
int  mc_step_days_minmax(mc_handle* h, int64_t day0, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                         const mc_weather_in_soa* in, const mc_forcing_soa* rest, mc_output_soa* out, mc_weather_out_soa* weather_host);

/* mc_step_days_minmax over the devices of a multi handle (declared above with mc_multi_*): every device expands and integrates its contiguous column range of
 * ONE set of host arrays [rows][n_sites]; returns the flagged-site count summed over devices */
int  mc_multi_step_days_minmax(mc_multi* m, int64_t day0, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                               const mc_weather_in_soa* in, const mc_forcing_soa* rest, mc_output_soa* out, mc_weather_out_soa* weather_host);

/* struct sizes of the loaded library, for bindings to check their mirrors against (index: 0 mc_model_desc, 1 mc_sites_soa, 2 mc_initial_soa,
 * 3 mc_state_soa, 4 mc_forcing_soa, 5 mc_output_soa, 6 mc_counters, 7 mc_weather_desc, 8 mc_solar_sites_soa, 9 mc_weather_in_soa,
 * 10 mc_weather_out_soa); -1 for an unknown index */
int64_t mc_abi_sizeof(int which);

/* measured FP64 DFMA peak of the handle's device in TFLOP/s (micro-benchmark; SURVEY §8d:
 * MEASURED_PEAKS.json has no FP64 figure). */
int  mc_measure_fp64_peak(mc_handle* h, double* tflops);

/* development probe (not part of the drop-in): work census of the asynchronous kernel (claims / items / cycles per task type, idle polls), collected
 * when the environment variable MC_ASYNC_STATS is set; returns the number of values written (32) and resets the census */
int  mc_get_async_stats(mc_handle* h, unsigned long long* out, int n);

/* development probe (not part of the drop-in): rate of the hottest device function over independent (site, node) items —
 * the headroom of a node-parallel formulation, see DESIGN.md */
int  mc_bench_conductivity(mc_handle* h, int64_t n_items, int reps, double* evals_per_s);

#ifdef __cplusplus
}
#endif
#endif /* MICROCLIMATE_B200_H */
