/*
 * mc_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar FP64 restatement, one site at a time, of Microclimate.jl's soil path
 *     solve_soil!(cache)   /root/reference/src/simulation.jl:355-819
 * and everything it calls.  Each function cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (microclimate.jl_b200/csrc) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" at the north-star tolerance.  Julia is not installed in the
 * build container and two dependencies that carry arithmetic are not vendored in the
 * reference tree: OrdinaryDiffEqTsit5/OrdinaryDiffEqCore (adaptive Tsit5 + PI controller +
 * automatic initial dt; compat "1.5"/"3", no Manifest) and FluidProperties.jl (rev "main").
 * Their published algorithms are restated below from the packages' documented behaviour.
 * What IS pinned (tests/test_oracle_golden.py): Goff-Gratch vapour pressure, Campbell-de
 * Vries soil properties, Campbell retention curve, sky temperature and the Monin-Obukhov
 * profile against internally consistent columns of the reference's NicheMapR golden CSVs,
 * and the end-to-end soil temperature / moisture of configs 1-2 at the reference's own
 * (loose) test tolerances.
 *
 * Units: unit-stripped SI (K, Pa, m, s, W/m2, kg/m2, J/kg); ODE time in minutes and dT/dt in
 * K/min exactly as the reference (src/simulation.jl:856, heat_transport_1d.jl:279).
 */
#include "../include/microclimate_b200.h"
#include "../include/mc_math.h"   /* bit-reproducible exp/log/pow/atan/cos shared with the CUDA engine (see its header) */

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

/* region tags for the operation-counting build (oracle/opcount/opcount.cpp); no-ops in the normal build */
#ifndef MCO_REGION
#define MCO_REGION(r) ((void)0)
#endif

namespace mco {

static const int MAXN = MC_MAX_SOIL_NODES;
static const int MAXM = MC_MAX_ODE_NODES;
static const double T0C = 273.15;
static const double WATER_HEAT_CAPACITY = 4184.0;      /* src/constants.jl:1 */
static const double WATER_DENSITY = 1000.0;            /* src/constants.jl:2 */
static const double VAPOUR_MOLAR_DENSITY_STP = 44.65;  /* src/constants.jl:3 */
static const double VAPOUR_DIFFUSIVITY_STP = 2.12e-5;  /* src/constants.jl:4 */
static const double LATENT_HEAT_FUSION = 333550.0;     /* src/constants.jl:12 */
static const double T_LO = T0C - 81.0;                 /* heat_transport_1d.jl:159 */

/* Julia's max/min propagate NaN; std::fmax does not. */
static inline double jmax(double a, double b) { return (a != a || b != b) ? std::numeric_limits<double>::quiet_NaN() : (a > b ? a : b); }
static inline double jmin(double a, double b) { return (a != a || b != b) ? std::numeric_limits<double>::quiet_NaN() : (a < b ? a : b); }
static inline double jclamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

/* ------------------------------------------------------------------------------------------
 * FluidProperties.jl (third party, un-vendored; reference Project.toml:29).  Restated from the
 * NicheMapR WETAIR/VAPPRS lineage the package documents; constants are data (mc_fluid_constants).
 * Call sites in the reference: heat_transport_1d.jl:239, campbell_devries.jl:117-119,
 * bulk_transfer.jl:27-28,67, campbell.jl:308,416-420, view_factor.jl:42, monin_obukhov.jl:196-197.
 * ------------------------------------------------------------------------------------------ */
/* vapour_pressure(GoffGratch(), T) [Pa]; ice branch below 273.16 K. */
static double goff_gratch(double T)
{
    const double Tp = T + 0.01;
    double loge;
    if (Tp <= 273.16) {
        const double r = 273.16 / Tp;
        loge = -9.09718 * (r - 1.0) - 3.56654 * mcm::log10(r) + 0.876793 * (1.0 - Tp / 273.16) + 0.7858350313586662 /* log10(6.1071) */;
    } else {
        const double r = 373.16 / Tp;
        loge = -7.90298 * (r - 1.0) + 5.02808 * mcm::log10(r)
             - 1.3816e-7 * (mcm::exp10(11.344 * (1.0 - Tp / 373.16)) - 1.0)
             + 8.1328e-3 * (mcm::exp10(-3.49149 * (r - 1.0)) - 1.0) + 3.0057148979490314 /* log10(1013.246) */;
    }
    return 100.0 * mcm::exp10(loge);
}

struct WetAir { double vapour_pressure, vapour_density, specific_heat, density, mixing_ratio; };

/* wet_air_properties(T, rh, P) */
static WetAir wet_air(const mc_fluid_constants& fc, double T, double rh, double P)
{
    WetAir w;
    const double esat = goff_gratch(T);
    const double e = esat * rh;
    const double eps = fc.molar_mass_water / fc.molar_mass_air;
    w.vapour_pressure = e;
    w.mixing_ratio = (eps * fc.ideal_gas_departure * e) / (P - fc.ideal_gas_departure * e);
    w.vapour_density = e * fc.molar_mass_water / (0.998 * fc.gas_constant * T);
    const double T_vir = T * ((1.0 + w.mixing_ratio / eps) / (1.0 + w.mixing_ratio));
    w.density = (fc.molar_mass_air / fc.gas_constant) * P / (0.999 * T_vir);
    w.specific_heat = (fc.cp_dry_air + w.mixing_ratio * fc.cp_water_vapour) / (1.0 + w.mixing_ratio);
    return w;
}
static inline double enthalpy_of_vaporisation(const mc_fluid_constants& fc, double T) { return fc.latent_heat_a - fc.latent_heat_b * (T - T0C); }
static inline double molar_enthalpy_of_vaporisation(const mc_fluid_constants& fc, double T) { return fc.molar_latent_heat_a - fc.molar_latent_heat_b * (T - T0C); }

/* ------------------------------------------------------------------------------------------
 * Forcing: 9 linear interpolants over 25 knots at 0,60,...,1440 min
 * (src/forcing.jl:8-37, src/simulation.jl:970-1021; Interpolations.jl BSpline(Linear())).
 * ------------------------------------------------------------------------------------------ */
enum { F_SOLAR = 0, F_ZENITH, F_SLOPE_ZENITH, F_TEMPERATURE, F_WIND, F_HUMIDITY, F_CLOUD, F_PRESSURE, F_DIFFUSE, F_COUNT };
struct DayForcing { double v[F_COUNT][25]; };

static inline double lerp25(const double* c, double t_min)
{
    double x = t_min / 60.0;           /* knot coordinate, 0-based */
    int i = (int)std::floor(x);
    if (i < 0) i = 0;
    if (i > 23) i = 23;
    const double d = x - (double)i;
    return (1.0 - d) * c[i] + d * c[i + 1];
}

struct Instant9 { double pressure, air_temperature, wind_speed, zenith, solar, cloud, humidity, slope_zenith, diffuse_fraction; };

/* interpolate_forcings, src/forcing.jl:24-37 (floors and clamps are part of the model) */
static Instant9 interpolate_forcings(const DayForcing& f, double t)
{
    Instant9 o;
    o.pressure = lerp25(f.v[F_PRESSURE], t);
    o.air_temperature = lerp25(f.v[F_TEMPERATURE], t);
    o.wind_speed = jmax(lerp25(f.v[F_WIND], t), 0.1);
    o.zenith = jmin(90.0, std::nearbyint(lerp25(f.v[F_ZENITH], t) * 1000.0) / 1000.0);
    o.solar = jmax(0.0, lerp25(f.v[F_SOLAR], t));
    o.cloud = jclamp(lerp25(f.v[F_CLOUD], t), 0.0, 1.0);
    o.humidity = jclamp(lerp25(f.v[F_HUMIDITY], t), 0.0, 1.0);
    o.slope_zenith = jmin(90.0, lerp25(f.v[F_SLOPE_ZENITH], t));
    o.diffuse_fraction = jclamp(lerp25(f.v[F_DIFFUSE], t), 0.0, 1.0);
    return o;
}

/* ------------------------------------------------------------------------------------------
 * Campbell-de Vries soil properties, src/soil_properties/campbell_devries.jl:79-142
 * ------------------------------------------------------------------------------------------ */
struct SoilProps { double conductivity, heat_capacity, density; };

static SoilProps soil_properties(const mc_model_desc& m, double P, double T, double theta,
                                 double bulk_density, double mineral_density,
                                 double mineral_conductivity, double mineral_heat_capacity)
{
    const mc_fluid_constants& fc = m.fluid;
    const double ga = m.de_vries_shape_factor;
    const double gc = 1.0 - 2.0 * ga;
    auto weight = [&](double lam_c, double lam_f) {
        return 2.0 / (3.0 * (1.0 + ga * (lam_c / lam_f - 1.0))) + 1.0 / (3.0 * (1.0 + gc * (lam_c / lam_f - 1.0)));
    };
    const double Tc = T - T0C;
    SoilProps o;
    o.heat_capacity = bulk_density / mineral_density * mineral_heat_capacity + theta * WATER_HEAT_CAPACITY;
    o.density = theta * WATER_DENSITY + bulk_density;
    const double lam_w = 0.554 + 2.24e-3 * Tc - 9.87e-6 * Tc * Tc;     /* :104 */
    const double lam_a = 0.024 + 7.73e-5 * Tc - 2.6e-8 * Tc * Tc;      /* :106 */
    const double Dv = VAPOUR_DIFFUSIVITY_STP * (fc.standard_pressure / P) * mcm::pow(T / 273.15, 1.75);  /* :109 */
    const double rho_hat = VAPOUR_MOLAR_DENSITY_STP * (P / fc.standard_pressure) * (273.15 / T);        /* :111 */
    const double Hv = molar_enthalpy_of_vaporisation(fc, T);                                            /* :112 */
    const double e0 = goff_gratch(T) * 0.99;                                                            /* :117 */
    const double em = goff_gratch(T - 1.0) * 0.99;
    const double ep = goff_gratch(T + 1.0) * 0.99;
    const double dedT = (ep - em) / 2.0;                                                                /* :122 */
    const double phi_m = bulk_density / mineral_density;
    const double phi_g = jmax(0.0, 1.0 - theta - phi_m);
    const double fw = 1.0 / (1.0 + mcm::pow(theta / m.return_flow_threshold, -m.recirculation_power));  /* :127 */
    const double lam_g = lam_a + Hv * dedT * fw * rho_hat * Dv / (P - e0);                              /* :130 */
    const double lam_f = lam_g + fw * (lam_w - lam_g);                                                  /* :132 */
    const double ww = weight(lam_w, lam_f), wm = weight(mineral_conductivity, lam_f), wg = weight(lam_g, lam_f);
    o.conductivity = (theta * ww * lam_w + phi_m * wm * mineral_conductivity + phi_g * wg * lam_g) /
                     (theta * ww + phi_m * wm + phi_g * wg);                                            /* :134-139 */
    return o;
}

/* ------------------------------------------------------------------------------------------
 * Monin-Obukhov, src/boundary_layer/monin_obukhov.jl
 * ------------------------------------------------------------------------------------------ */
static const double RHOCPTKG_SI = 6.003e-8 * 4.184 * 3600.0 * 1.0e8;   /* :135,:234  cal min2/cm4 -> J s2/m4 */

static inline double calc_rho_cp(double Tmean) { return (0.08472 / Tmean) * 4.184e6; }                 /* :280-282 */
static inline double sublayer_stanton(double z0, double ustar) { return 0.62 / mcm::pow(z0 * 100.0 * (ustar * 6000.0) / 12.0, 0.45); } /* :414-416 */
static inline double bulk_stanton(double lzr) { return 0.64 / lzr; }                                    /* :423-425 */
static inline double bulk_stanton(double lzr, double z, double L) { return (0.64 / lzr) * (1.0 - 0.1 * z / L); } /* :432-434 */
static inline double convective_flux(double rho_cp, double dT, double ustar, double Stb, double Sts) { return rho_cp * dT * ustar * Stb / (1.0 + Stb / Sts); } /* :386-388 */
static inline double calc_phi_m(double z, double gamma, double L) { return std::sqrt(std::sqrt(1.0 - gamma * z / L)); } /* :464-468 */
static inline double calc_psi_m(double x) { return 2.0 * mcm::log((1.0 + x) / 2.0) + mcm::log((1.0 + x * x) / 2.0) - 2.0 * mcm::atan(x) + M_PI / 2.0; } /* :491-493 */
static inline double calc_psi_h(double x) { return 2.0 * mcm::log((1.0 + x * x) / 2.0); }   /* :516-518 */

struct ObukhovOut { double L, Sts, Stb, ustar, psi_h, Q, T_z0; int iters; };

/* calc_Obukhov_length, :527-577 */
static ObukhovOut calc_obukhov(double T_ref, double T_surf, double v_ref, double z0, double z, double kappa,
                               double dT, double rho_cp, double gamma, int max_iter, double tol, double L0)
{
    ObukhovOut o; o.L = L0; o.Q = 0; o.Stb = 0; o.Sts = 0; o.psi_h = 0; o.ustar = 0; o.iters = 0;
    double rel = 1.0; int count = 0; const double tiny = 1.0e-6;
    while (rel > tol && count < max_iter) {
        MCO_REGION(R_OBUKHOV);
        count++;
        const double x = calc_phi_m(z, gamma, o.L);
        const double psi_m = calc_psi_m(x);
        o.psi_h = calc_psi_h(x);
        double lrc = mcm::log(z / z0) - psi_m;
        if (lrc <= 0.0) lrc = tiny;
        o.ustar = kappa * v_ref / lrc;
        if (o.ustar < tiny) o.ustar = tiny;
        o.Sts = sublayer_stanton(z0, o.ustar);
        o.Stb = bulk_stanton(lrc, z, o.L);
        o.Q = convective_flux(rho_cp, dT, o.ustar, o.Stb, o.Sts);
        const double Lnew = RHOCPTKG_SI * o.ustar * o.ustar * o.ustar / o.Q;
        rel = std::fabs((Lnew - o.L) / o.L);
        o.L = Lnew;
    }
    o.iters = count;
    o.T_z0 = (T_ref * o.Stb + T_surf * o.Sts) / (o.Stb + o.Sts);
    return o;
}

/* analysis hook (tools only): Newton iterations per (site, hour, sub-step) */
static int32_t* g_trace_sub = nullptr; static thread_local int64_t tl_trace_row = -1;
struct Counters { uint64_t n_rhs = 0, n_accept = 0, n_reject = 0, n_obukhov = 0, n_newton = 0, n_infil = 0, n_day_passes = 0; };

/* surface_fluxes, :208-245 -> (Q_conv, dT) */
static void surface_fluxes(const mc_model_desc& m, double T_surf, double T_air, double wind, double zenith,
                           double z0, double zref, double L_prev, double& Qc, double& dT, Counters& cn)
{
    const double kappa = m.karman_constant;
    const double lzr = mcm::log(zref / z0);
    dT = T_air - T_surf;
    const double rho_cp = calc_rho_cp((T_surf + T_air) / 2.0);
    if (T_air >= T_surf || zenith >= 90.0) {
        const double ustar = jmax(kappa * wind / lzr, 1.0e-6);
        Qc = convective_flux(rho_cp, dT, ustar, bulk_stanton(lzr), sublayer_stanton(z0, ustar));
    } else {
        const double L0 = L_prev >= 0.0 ? -0.3 : L_prev;
        ObukhovOut o = calc_obukhov(T_air, T_surf, wind, z0, zref, kappa, dT, rho_cp, 16.0, 30, 1e-2, L0);
        cn.n_obukhov += o.iters;
        Qc = o.Q;
    }
}

/* evaporation(::BulkTransferEvaporation), src/evaporation/bulk_transfer.jl:11-53 -> Q_evap [W/m2] */
static double evaporation_Q(const mc_fluid_constants& fc, double T_surf, double T_air, double rh, double rh_surf,
                            double hd, double P, double wetness, bool saturated)
{
    const double Tc = T_surf < T_LO ? T_LO : T_surf;
    const double rho_s = wet_air(fc, Tc, rh_surf, P).vapour_density;
    const double rho_a = wet_air(fc, T_air, rh, P).vapour_density;
    const double w = saturated ? 1.0 : wetness;
    const double flux = w * hd * (rho_s - rho_a);
    return flux * enthalpy_of_vaporisation(fc, Tc);
}

static inline double heat_transfer_coefficient(double Qc, double dT) { return dT == 0.0 ? 0.5 : jmax(std::fabs(Qc / dT), 0.5); } /* monin_obukhov.jl:397-398 */
static inline double mass_transfer_coefficient(double hc, double cp, double rho) { return (hc / (cp * rho)) * mcm::pow(0.71 / 0.60, 0.666); } /* :406-407 */

/* surface_convection_evaporation, bulk_transfer.jl:55-81 */
static void surface_convection_evaporation(const mc_model_desc& m, double T_surf, double T_air, double wind, double rh,
                                           double P, double z0, double zref, double zenith, double wetness,
                                           double L_prev, double& Qe, double& Qc, Counters& cn)
{
    double dT;
    surface_fluxes(m, T_surf, T_air, wind, zenith, z0, zref, L_prev, Qc, dT, cn);
    const double hc = heat_transfer_coefficient(Qc, dT);
    const WetAir wa = wet_air(m.fluid, T_air, rh, P);
    const double hd = mass_transfer_coefficient(hc, wa.specific_heat, wa.density);
    Qe = evaporation_Q(m.fluid, T_surf, T_air, rh, 1.0, hd, P, wetness, false);
}

/* atmospheric_radiation, src/radiation/atmospheric/{campbell_norman,swinbank}.jl */
static double atmospheric_radiation(const mc_model_desc& m, double e_Pa, double T_air)
{
    const double sigma = m.fluid.stefan_boltzmann;
    if (m.atmospheric_radiation == MC_ATMRAD_SWINBANK)
        return (9.2e-6 * T_air * T_air) * sigma * mcm::pow4(T_air);
    const double Tp = T_air + 0.01;
    return (1.72 * mcm::pow((e_Pa / 1000.0) / Tp, 1.0 / 7.0)) * sigma * mcm::pow4(Tp);
}

/* precompute_longwave_sky -> sky temperature only (the RHS recomputes the fluxes itself),
 * src/radiation/longwave/view_factor.jl:23-67 */
static double sky_temperature(const mc_model_desc& m, double svf, double P, double rh, double T_ref, double eps_cloud,
                              double cloud, double shade, bool has_lw, double supplied_lw)
{
    const double sigma = m.fluid.stefan_boltzmann;
    const double cloud_rad = sigma * eps_cloud * mcm::pow4(T_ref - 2.0);
    const double hill_rad = sigma * eps_cloud * mcm::pow4(T_ref);
    double lw_sky;
    if (has_lw) lw_sky = supplied_lw * (1.0 - shade);
    else {
        const WetAir wa = wet_air(m.fluid, T_ref, rh, P);
        const double atm = atmospheric_radiation(m, wa.vapour_pressure, T_ref);
        lw_sky = (atm * (1.0 - cloud) + cloud_rad * cloud) * (1.0 - shade);
    }
    const double lw_veg = shade * hill_rad;
    const double incoming = (lw_sky + lw_veg) * svf + hill_rad * (1.0 - svf);
    return mcm::pow(incoming / sigma, 0.25);
}


/* ------------------------------------------------------------------------------------------
 * Snow: SnowModel{8} (src/snow/snow_model.jl) and the apparent-heat-capacity closures
 * (src/apparent_heat_capacity/ *.jl).  Depths in cm, densities in g/cm3, as upstream.
 * The reference's `snow_depth_hourly[nsteps]` scratch array is only ever read at [step] and
 * [step-1]; it is carried as two scalars (sd_cur, sd_prev).
 * ------------------------------------------------------------------------------------------ */
static const int NSN = MC_SNOW_NODES;
struct SnowState { double depth, age, days_since, density, prev_density, cum_melt, extra, sum_phase; int active; };
struct SnowScratch { double node_depths[NSN]; double sd_prev, sd_cur; double meanT[NSN], meanTpast[NSN]; };

/* apparent_heat_capacity(::BonacinaStep / TanhSmoothed / Gaussian / WestermannSigmoid, T, cp_base, L_fusion) */
static inline double mc_tanh(double x) { const double e = mcm::exp(2.0 * x); return x > 20.0 ? 1.0 : (x < -20.0 ? -1.0 : (e - 1.0) / (e + 1.0)); }
static double apparent_heat_capacity(const mc_snow_desc& sm, double T, double cp_base)
{
    const double Tc = T - T0C, Lf = LATENT_HEAT_FUSION;
    switch (sm.ahc_kind) {
    case MC_AHC_TANH: {                                                         /* tanh.jl:40-49 */
        const double lo = (Tc - sm.ahc_T_low_C) / sm.ahc_smoothing_K, hi = (Tc - sm.ahc_T_high_C) / sm.ahc_smoothing_K;
        return cp_base + (Lf / sm.ahc_window_K) * (0.5 * (mc_tanh(lo) - mc_tanh(hi)));
    }
    case MC_AHC_GAUSSIAN: {                                                     /* gaussian.jl:31-38 */
        const double x = (Tc - sm.ahc_T_melt_C) / sm.ahc_width_K;
        return cp_base + (Lf / (sm.ahc_width_K * std::sqrt(2.0 * M_PI))) * mcm::exp(-0.5 * x * x);
    }
    case MC_AHC_WESTERMANN: {                                                   /* westermann.jl:35-43 */
        const double x = (Tc - sm.ahc_T_melt_C) / sm.ahc_width_K, th = mc_tanh(x);
        return cp_base + (Lf / (2.0 * sm.ahc_width_K)) * (1.0 - th * th);
    }
    default:                                                                    /* bonacina.jl:30-36 */
        if (Tc > sm.ahc_T_low_C && Tc <= sm.ahc_T_high_C) return cp_base + Lf / sm.ahc_window_K;
        return cp_base;
    }
}

static SnowState initial_snow_state(const mc_snow_desc& sm, double depth, double density_or_neg)     /* snow_model.jl:63-74 */
{
    const double dflt = sm.density_function[0] > 0.0 ? sm.density_function[1] : sm.snow_density;
    const double d = density_or_neg > 0.0 ? density_or_neg : dflt;
    SnowState st; st.depth = depth; st.age = 0.0; st.days_since = 0.3; st.density = d; st.prev_density = d;
    st.cum_melt = 0.0; st.extra = 0.0; st.active = 0; st.sum_phase = 0.0;
    return st;
}
static double evolve_snow_density(const mc_snow_desc& sm, const SnowState& st)                        /* :277-288 */
{
    const double* df = sm.density_function;
    if (!(df[0] > 0.0)) return st.density;
    if (df[2] > 0.0) return (df[0] - df[1]) * (1.0 - mcm::exp(-df[2] * st.depth - df[3] * st.age)) + df[1];
    return jmin(0.9167, df[0] * st.age + df[1]);
}
static inline double snow_albedo(double days_since) { return (-9.874 * mcm::log(jmax(days_since, 0.3)) + 78.3434) / 100.0; }   /* :268-270 */
static double snow_specific_heat(const mc_fluid_constants& fc, double dens, double P)                 /* :323-329 */
{
    const double mr = wet_air(fc, 273.15, 1.0, P).mixing_ratio;
    return 2100.0 * dens + (1005.0 + 1820.0 * (mr / 1.0 + mr)) * (1.0 - dens);
}
/* activate_snow_nodes!, :108-142.  `sd` is snow_depth_hourly[step]. */
static void activate_snow_nodes(const mc_snow_desc& sm, SnowState& st, SnowScratch& sc, double& sd)
{
    const double cursnow = sd;
    int active = st.active;
    if (cursnow > 300.0) active = 0;
    if (cursnow < sm.min_snow_depth_cm) {
        for (int i = 0; i < NSN; i++) sc.node_depths[i] = 0.0;
        sd = 0.0; st.active = 0; st.depth = 0.0;
        return;
    }
    int maxsnode = 0;
    for (int i = 1; i <= NSN; i++) if (cursnow > sm.node_thresholds_cm[i - 1]) maxsnode = i;
    maxsnode = std::min(maxsnode, NSN - 1);
    for (int i = 1; i <= maxsnode; i++) sc.node_depths[i + (NSN - maxsnode) - 1] = sm.node_thresholds_cm[i - 1];
    if (maxsnode < active) for (int i = 1; i <= NSN - maxsnode; i++) sc.node_depths[i - 1] = 0.0;
    st.depth = cursnow; st.active = maxsnode;
}
/* compute_effective_depths!, :155-167 (metres) */
static void compute_effective_depths(const mc_snow_desc& sm, const SnowState& st, const SnowScratch& sc, const double* soil_depths, int N, double* out)
{
    const double off = st.depth >= sm.min_snow_depth_cm ? st.depth : 0.0;
    for (int i = 0; i < NSN; i++) out[i] = sc.node_depths[i] / 100.0;
    for (int i = 0; i < N; i++) out[NSN + i] = soil_depths[i] + off / 100.0;
}
/* snow_layer_geometry, :304-317 (1-based layer_index; returns 0-based node index) */
static void snow_layer_geometry(const mc_snow_desc& sm, const SnowState& st, const SnowScratch& sc, int layer_index, int& node_index, double& thickness)
{
    const int maxsnode = st.active;
    const int ni = std::max(1, std::min(NSN, layer_index + NSN - 1 - maxsnode));
    if (layer_index != maxsnode + 1) {
        const int upper = std::max(1, layer_index + NSN - maxsnode);
        thickness = sc.node_depths[upper - 1] - sc.node_depths[ni - 1];
    } else {
        thickness = jmax(0.0, st.depth - sm.node_thresholds_cm[std::max(1, maxsnode) - 1]);
    }
    node_index = ni - 1;
}
/* snow_thermal_melt!, :333-365 -> cm of water; fills meanT / meanTpast (degC) */
static double snow_thermal_melt(const mc_model_desc& m, const SnowState& st, SnowScratch& sc, const double* Ts, const double* Ts_before,
                                double P, double T_soil, double T_soil_before)
{
    const mc_snow_desc& sm = m.snow;
    for (int i = 0; i < NSN; i++) {
        const double below = i < NSN - 1 ? Ts[i + 1] : T_soil, below_past = i < NSN - 1 ? Ts_before[i + 1] : T_soil_before;
        sc.meanT[i] = (Ts[i] + below) / 2.0 - T0C;
        sc.meanTpast[i] = (Ts_before[i] + below_past) / 2.0 - T0C;
    }
    const double cp = snow_specific_heat(m.fluid, st.density, P);
    double total = 0.0;                                                       /* kg/m2 */
    for (int j = 1; j <= st.active + 1; j++) {
        int ni; double thick; snow_layer_geometry(sm, st, sc, j, ni, thick);
        const double mass = thick * st.density * 10.0;                        /* cm * g/cm3 -> kg/m2 */
        if (sc.meanT[ni] > sm.melt_threshold_C) {
            const double dT = sc.meanT[ni] - sm.melt_threshold_C;
            const double sensible = dT * cp * mass, latent = LATENT_HEAT_FUSION * mass;
            total += (sensible + latent) / LATENT_HEAT_FUSION;
        }
    }
    return total / 10.0;                                                      /* kg/m2 of water / (1 g/cm3) -> cm */
}
/* snow_phase_transition, :369-467 */
static void snow_phase_transition(const mc_model_desc& m, SnowState& st, SnowScratch& sc, double* Ts, double P, double& T_soil_surface)
{
    const mc_snow_desc& sm = m.snow;
    if (st.depth >= 200.0) return;
    if (st.depth < sm.min_snow_depth_cm) return;
    const double cp = snow_specific_heat(m.fluid, st.density, P);
    bool clamp_soil = false; bool clamp[NSN]; double phase_heat[NSN], layer_mass[NSN];
    for (int i = 0; i < NSN; i++) { clamp[i] = false; phase_heat[i] = 0.0; layer_mass[i] = 0.0; }
    for (int j = 1; j <= st.active + 1; j++) {
        if (j > NSN) continue;
        int ni; double thick; snow_layer_geometry(sm, st, sc, j, ni, thick);
        if (!(sc.node_depths[ni] > 0.0)) continue;
        const double mass = thick * st.density * 10.0;
        if (sc.meanTpast[ni] >= 0.0 && sc.meanT[ni] <= 0.0) {
            layer_mass[ni] = jmax(0.0, mass);
            phase_heat[ni] = (sc.meanTpast[ni] - sc.meanT[ni]) * cp * layer_mass[ni];
            clamp[ni] = true;
            if (ni == NSN - 1) clamp_soil = true; else clamp[ni + 1] = true;
        }
    }
    double new_sum = st.sum_phase, total_mass = 0.0;
    { double sp = 0.0; for (int i = 0; i < NSN; i++) sp += phase_heat[i]; new_sum = st.sum_phase + sp; }
    for (int i = 0; i < NSN; i++) total_mass += layer_mass[i];
    if (total_mass > 0.0 && new_sum > LATENT_HEAT_FUSION * total_mass) {
        bool over[NSN]; for (int i = 0; i < NSN; i++) over[i] = false;
        for (int i = 0; i < NSN; i++) if (clamp[i] && phase_heat[i] > 0.0) { over[i] = true; phase_heat[i] = 0.0; if (i < NSN - 1) over[i + 1] = true; }
        for (int i = 0; i < NSN; i++) if (over[i]) Ts[i] = T0C - 0.5;
        if (clamp_soil) T_soil_surface = T0C - 0.5;
        new_sum = 0.0;
    } else {
        for (int i = 0; i < NSN; i++) if (clamp[i]) Ts[i] = T0C;
        if (clamp_soil) T_soil_surface = T0C;
    }
    st.sum_phase = new_sum;
}
struct SnowUpdateOut { double snowfall, thermal_melt, rain_melt; };
/* update_snow!, :471-594.  `step` is the 1-based global hour index. */
static SnowUpdateOut update_snow(const mc_model_desc& m, SnowState& st, SnowScratch& sc, const double* Ts, const double* Ts_before,
                                 double T_air, double rainfall, double P, int64_t step, double T_soil, double T_soil_before, double Q_evap,
                                 bool hourly, double shade, int day_of_year)
{
    const mc_snow_desc& sm = m.snow;
    const bool first = step == 1, midnight = ((step - 1) % 24) == 0;
    const double prev_dens = st.density;
    const double new_dens = evolve_snow_density(sm, st);
    const double ratio = new_dens / prev_dens;
    if (step > 1 && ratio != 0.0 && std::isfinite(ratio)) sc.sd_prev /= ratio;
    double snowfall = 0.0, days_since = st.days_since;
    if (T_air - T0C <= sm.snow_temperature_threshold_C && rainfall > 0.0) {
        if (midnight || hourly) {
            snowfall = rainfall * sm.rain_multiplier * sm.undercatch / new_dens / 10.0;        /* kg/m2 / (g/cm3) -> cm */
            if (shade > 0.0) snowfall *= (1.0 - sm.canopy_interception);
            days_since = 0.3;
        } else days_since += 1.0 / 25.0;
    } else days_since += 1.0 / 25.0;
    days_since = jmax(days_since, 0.3);
    double sublimation = 0.0;
    if (T_soil - T0C > 0.0 && Q_evap > 0.0) {
        const double gw = Q_evap / enthalpy_of_vaporisation(m.fluid, T_soil);                   /* kg/s/m2 */
        sublimation = gw * 3600.0 / new_dens / 10.0;
    }
    const double previous_depth = step > 1 ? sc.sd_prev : st.depth;
    double rain_melt = 0.0;
    if (T_air - T0C >= sm.snow_temperature_threshold_C && rainfall > 0.0 && previous_depth >= sm.min_snow_depth_cm) {
        if (midnight || hourly) {
            const double water_depth = rainfall / WATER_DENSITY * 100.0;                         /* cm */
            rain_melt = sm.rain_melt_factor * water_depth * (T_air - T0C) * 24.0 / new_dens;
            if (hourly) rain_melt /= 24.0;
            rain_melt = jmax(0.0, rain_melt);
        }
    }
    SnowState for_melt = st; for_melt.density = new_dens;
    double thermal_melt;
    if (first || previous_depth < sm.min_snow_depth_cm) {
        snow_thermal_melt(m, for_melt, sc, Ts, Ts_before, P, T_soil, T_soil_before);
        thermal_melt = 0.0;
    } else {
        const double raw = snow_thermal_melt(m, for_melt, sc, Ts, Ts_before, P, T_soil, T_soil_before);
        thermal_melt = day_of_year > 1 ? raw * sm.snow_melt_factor : raw;
    }
    double cum_melt = st.cum_melt + thermal_melt;
    if (step % 24 == 0) cum_melt = 0.0;
    const double carried = st.extra / new_dens / 10.0;
    const double net_acc = jmax(0.0, snowfall - sublimation) + carried;
    double extra = 0.0;
    double new_depth = jmax(0.0, previous_depth + net_acc - rain_melt - thermal_melt);
    if (new_depth < sm.min_snow_depth_cm && (step <= 1 || sc.sd_prev < 1e-8)) { extra = new_depth * new_dens * 10.0; new_depth = 0.0; }
    sc.sd_cur = new_depth;
    double age;
    if (first) age = 0.0;
    else if (sc.sd_cur > 0.0 && sc.sd_prev > 0.0) age = st.age + 1.0 / 25.0;
    else age = 0.0;
    st.depth = new_depth; st.age = age; st.days_since = days_since; st.density = new_dens; st.prev_density = new_dens;
    st.cum_melt = cum_melt; st.extra = extra;
    SnowUpdateOut o; o.snowfall = snowfall; o.thermal_melt = thermal_melt; o.rain_melt = rain_melt;
    return o;
}
/* adjust_snow_near_nodes!, :598-629 */
static void adjust_snow_near_nodes(const mc_snow_desc& sm, SnowState& st, SnowScratch& sc, int64_t step)
{
    double snowout = sc.sd_cur;
    double threshold = sm.node_thresholds_cm[std::max(1, st.active) - 1];
    double extra = st.extra;
    if (step > 1 && snowout > 0.0) {
        if ((snowout - threshold < 0.5) && (snowout <= sc.sd_prev)) {
            extra += (snowout - threshold) * st.density * 10.0;
            sc.sd_cur = jmax(0.0, threshold - 0.1);
        }
    }
    st.extra = extra;
    activate_snow_nodes(sm, st, sc, sc.sd_cur);
    snowout = sc.sd_cur;
    if (snowout > 0.0) {
        threshold = sm.node_thresholds_cm[std::max(1, st.active) - 1];
        if (snowout < sm.min_snow_depth_cm + 0.5 || snowout < threshold + 0.5) {
            extra = 0.5 * st.density * 10.0;
            sc.sd_cur = threshold + 0.5;
        }
    }
    st.extra = extra; st.depth = sc.sd_cur;
}

/* ------------------------------------------------------------------------------------------
 * Per-site problem data gathered from the SoA buffers
 * ------------------------------------------------------------------------------------------ */
struct Site {
    double z0, slope, svf, albedo;
    double bd[MAXN], md[MAXN], lam_m[MAXN], c_m[MAXN], psi_e[MAXN], ks[MAXN], bb[MAXN], rootd[MAXN];
};

/* environment_instant (get_day + get_instant, src/simulation.jl:1025-1070) */
struct EnvInstant {
    double lai, shade, eps_surf, eps_cloud, soil_wetness_daily, T_deep, rain_daily, albedo_daily; bool has_albedo;
    double P, T_ref, wind, rh, zenith, cloud, solar, lw; bool has_lw;
};

/* parameters of the ODE right-hand side: SoilEnergyInputs, heat_transport_1d.jl:41-63 */
struct RhsParams {
    const mc_model_desc* m; const Site* site; const DayForcing* forcing;
    double depths[MAXM + 1];      /* effective ODE depths (snow + soil), metres */
    double theta[MAXN];           /* p.soil_moisture (aliases the evolving state in the reference) */
    double shade, eps_surf, eps_cloud;
    double soil_wetness, albedo, Q_freeze;
    double L_prev;                /* buffers.soil_energy_balance.obukhov_length_prev */
    int M, n_snow;
    const SnowState* snow; const SnowScratch* snow_scratch;   /* p.snow_state, buffers.snow (n_snow > 0) */
};

/* soil_energy_balance, heat_transport_1d.jl:137-302.  u, du in K and K/min. */
static void soil_energy_balance(const double* u, const RhsParams& p, double t, double* du, Counters& cn)
{
    MCO_REGION(R_RHS);
    cn.n_rhs++;
    const mc_model_desc& m = *p.m; const Site& s = *p.site; const int M = p.M; const int ns = p.n_snow; const int N = M - ns;
    const double sigma = m.fluid.stefan_boltzmann;
    const Instant9 f = interpolate_forcings(*p.forcing, t);
    const double T_hi = m.maximum_surface_temperature_K;
    double T[MAXM], lam[MAXM], cp[MAXM], rho[MAXM], C[MAXM], G[MAXM], ld[MAXM + 2];
    for (int i = 0; i < M; i++) T[i] = jclamp(u[i], T_LO, T_hi);                          /* :159-161 */
    for (int i = 0; i < N; i++) {                                                          /* :165 -> soil_properties! */
        SoilProps sp = soil_properties(m, f.pressure, T[ns + i], p.theta[i], s.bd[i], s.md[i], s.lam_m[i], s.c_m[i]);
        lam[ns + i] = sp.conductivity; cp[ns + i] = sp.heat_capacity; rho[ns + i] = sp.density;
    }
    if (ns > 0) {                                                                          /* write_snow_properties!, snow_model.jl:171-208 */
        const mc_snow_desc& sm = m.snow; const SnowState& st = *p.snow; const SnowScratch& sc = *p.snow_scratch;
        const double snowdens = evolve_snow_density(sm, st);
        if (st.depth >= sm.min_snow_depth_cm) {
            const double cps = snow_specific_heat(m.fluid, snowdens, f.pressure);
            double k;
            if (sm.snow_conductivity > 0.0) k = sm.snow_conductivity;
            else { const double r = snowdens * 1000.0; k = 0.00395 + 0.00084 * r - 1.7756e-6 * (r * r) + 3.80635e-9 * (r * r * r); }
            for (int i = 0; i < ns; i++) {
                lam[i] = k;
                cp[i] = apparent_heat_capacity(sm, T[i], cps);
                const bool active = sc.node_depths[i] > 0.0 || (i < ns - 1 && sc.node_depths[std::min(ns - 1, i + 1)] > 0.0);
                rho[i] = active ? snowdens * 1000.0 : 0.0;
            }
        } else {
            for (int i = 0; i < ns; i++) { lam[i] = lam[ns]; cp[i] = cp[ns]; rho[i] = 0.0; }
        }
    }
    for (int i = 0; i < M; i++) ld[i] = p.depths[i];
    ld[M] = 0.0;                                                                           /* :127,173 sentinel */
    for (int i = 0; i < M; i++) {                                                          /* :175-210 */
        const double vhc = rho[i] * cp[i];
        if (i == 0) {
            const double d2 = ld[1];
            C[0] = vhc * d2 / 2.0;
            G[0] = d2 == 0.0 ? 0.0 : lam[0] / d2;
        } else {
            const double dd = ld[i + 1] - ld[i];
            if (i < ns && ld[i] == 0.0 && ld[i + 1] != 0.0) C[i] = vhc * ld[i + 1];
            else C[i] = vhc * (ld[i + 1] - ld[i - 1]) / 2.0;
            const double pdd = ld[i] - ld[i - 1];
            G[i] = dd == 0.0 ? (pdd == 0.0 ? 0.0 : lam[i] / pdd) : lam[i] / dd;
        }
    }
    int j = 0; while (j < M - 1 && C[j] == 0.0) j++;                                       /* :216 */

    double Q_solar = (1.0 - p.albedo) * f.solar * (1.0 - p.shade);                         /* :223 */
    if (s.slope > 0.0 && f.zenith < 90.0) {                                                /* :224-232 */
        const double cz = mcm::cos(f.zenith * M_PI / 180.0), csz = mcm::cos(f.slope_zenith * M_PI / 180.0);
        const double direct = (1.0 - f.diffuse_fraction) * Q_solar, diffuse = f.diffuse_fraction * Q_solar;
        Q_solar = jmax(0.0, (direct / cz) * csz) + diffuse * s.svf;
    }
    const WetAir wa = wet_air(m.fluid, f.air_temperature, f.humidity, f.pressure);         /* :239 */
    const double atm = atmospheric_radiation(m, wa.vapour_pressure, f.air_temperature);
    const double cloud_rad = sigma * p.eps_cloud * mcm::pow4(f.air_temperature - 2.0);
    const double hill_rad = sigma * p.eps_cloud * mcm::pow4(f.air_temperature);
    const double lw_sky = (atm * (1.0 - f.cloud) + cloud_rad * f.cloud) * (1.0 - p.shade);
    const double lw_veg = p.shade * hill_rad;
    const double incoming = (lw_sky + lw_veg) * s.svf + hill_rad * (1.0 - s.svf);
    const double out_coeff = (1.0 - p.shade) * sigma * p.eps_surf;
    const double Q_ir = incoming - out_coeff * mcm::pow4(T[j]) - p.shade * hill_rad;     /* :250 */

    const double cond_ref = (t == 0.0 && j > 0) ? T[0] : T[j];                             /* :255 */
    const double Q_cond = j < M - 1 ? G[j] * (T[j + 1] - cond_ref) : 0.0;

    double Ts = T[j];                                                                      /* :260-266 */
    if (std::fabs(Ts - T0C) > 100.0) Ts = f.air_temperature + 1.0;
    if (Ts == f.air_temperature) Ts = Ts + 0.1;
    double Qe, Qc;
    surface_convection_evaporation(m, Ts, f.air_temperature, f.wind_speed, f.humidity, f.pressure, s.z0,
                                   m.reference_height_m, f.zenith, p.soil_wetness, p.L_prev, Qe, Qc, cn);
    const double surf = 60.0 * (Q_solar + Q_ir + Q_cond + Qc + p.Q_freeze - Qe) / C[j];    /* :279 */
    for (int i = 0; i < M; i++) {                                                          /* :287-299 */
        if (i <= j) du[i] = surf;
        else if (i == M - 1) du[i] = 0.0;
        else du[i] = C[i] == 0.0 ? 0.0 : 60.0 * (G[i - 1] * (T[i - 1] - T[i]) + G[i] * (T[i + 1] - T[i])) / C[i];
    }
    (void)N;
}

/* ------------------------------------------------------------------------------------------
 * OrdinaryDiffEqTsit5 + OrdinaryDiffEqCore (third party, un-vendored; compat 1.5 / 3).
 * Restatement of: Tsit5ConstantCache perform_step!, PIController (beta1=7/50, beta2=2/25,
 * gamma=9/10, qmin=1/5, qmax=10, qoldinit=1e-4) with FastPower.fastpower, Hairer-Wanner
 * ode_determine_initdt (order 5), tstop clipping, FSAL handling incl. reinit!(reinit_cache=false)
 * leaving fsalfirst stale (SURVEY §8c hazard H1).  Call sites: src/simulation.jl:853-903,736-741.
 * ------------------------------------------------------------------------------------------ */
namespace tsit5 {
static const double c1 = 0.161, c2 = 0.327, c3 = 0.9, c4 = 0.9800255409045097;
static const double a21 = 0.161;
static const double a31 = -0.008480655492356989, a32 = 0.335480655492357;
static const double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
static const double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365, a54 = -0.09249506636175525;
static const double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159, a64 = -0.071584973281401, a65 = -0.028269050394068383;
static const double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996, a74 = 1.379008574103742, a75 = -3.290069515436081, a76 = 2.324710524099774;
static const double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469, bt3 = 0.007880878010261995,
                    bt4 = -0.1447110071732629, bt5 = 0.5823571654525552, bt6 = -0.45808210592918697, bt7 = 0.015151515151515152;
}

/* FastPower.fastpower(x::Float64, y::Float64): Float32 log2 by Goldberg's rational fit, exp2 */
static inline float fastlog2f(float x)
{
    const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
    uint32_t ux; std::memcpy(&ux, &x, 4);
    const uint32_t ex = (ux & 0x7F800000u) >> 23;
    float signif, fexp; uint32_t u2;
    if (ux & 0x00400000u) { u2 = (ux & 0x007FFFFFu) | 0x3f000000u; std::memcpy(&signif, &u2, 4); fexp = (float)ex - 126.0f; }
    else                  { u2 = (ux & 0x007FFFFFu) | 0x3f800000u; std::memcpy(&signif, &u2, 4); fexp = (float)ex - 127.0f; }
    signif = signif - 1.0f;
    return fexp + signif * (a * signif + b) / (signif + c);
}
static inline double fastpow(double x, double y)
{
    const float l = fastlog2f((float)x) * (float)y;
    return (double)(float)mcm::exp2((double)l);   /* correctly rounded Float32 exp2 */
}

struct Integrator {
    int M;
    double u[MAXM], uprev[MAXM], fsalfirst[MAXM], fsallast[MAXM];
    double t, dt, dtpropose, qold, q11, EEst;
    int iter; bool accept_step, u_modified;
    double tstop_next;       /* first(tstops) */
    double reltol, abstol;
    int flags;
};

static inline double rms_norm(const double* x, int n) { double s = 0; for (int i = 0; i < n; i++) s += x[i] * x[i]; return std::sqrt(s / n); }

/* ode_determine_initdt (out-of-place), OrdinaryDiffEqCore initdt.jl */
static double determine_initdt(const Integrator& I, const RhsParams& p, Counters& cn)
{
    MCO_REGION(R_STEP);
    const int M = I.M; const double dtmax = 1440.0;
    double sk[MAXM], tmp[MAXM], f0[MAXM], f1[MAXM], u1[MAXM];
    for (int i = 0; i < M; i++) sk[i] = I.abstol + std::fabs(I.u[i]) * I.reltol;
    for (int i = 0; i < M; i++) tmp[i] = I.u[i] / sk[i];
    const double d0 = rms_norm(tmp, M);
    soil_energy_balance(I.u, p, I.t, f0, cn);
    for (int i = 0; i < M; i++) if (!std::isfinite(f0[i])) return 1e-12;
    for (int i = 0; i < M; i++) tmp[i] = f0[i] / sk[i];
    const double d1 = rms_norm(tmp, M);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
    dt0 = std::min(dt0, dtmax);
    for (int i = 0; i < M; i++) u1[i] = I.u[i] + dt0 * f0[i];
    soil_energy_balance(u1, p, I.t + dt0, f1, cn);
    for (int i = 0; i < M; i++) tmp[i] = (f1[i] - f0[i]) / sk[i];
    const double d2 = rms_norm(tmp, M) / dt0;
    const double md = std::max(d1, d2);
    double dt1;
    if (md <= 1e-15) dt1 = std::max(1e-6, dt0 * 1e-3);
    else dt1 = mcm::exp10(-(2.0 + mcm::log10(md)) / 5.0);
    return std::min(std::min(100.0 * dt0, dt1), dtmax);
}

/* SciMLBase.reinit!(integrator, u0; t0=0, erase_sol=false, reset_dt=true, reinit_cache=false),
 * src/simulation.jl:868-880.  fsalfirst is deliberately NOT recomputed (H1). */
static void reinit(Integrator& I, const double* u0, const RhsParams& p, Counters& cn)
{
    for (int i = 0; i < I.M; i++) { I.u[i] = u0[i]; I.uprev[i] = u0[i]; }
    I.t = 0.0; I.iter = 0; I.u_modified = false; I.accept_step = false;
    I.qold = 1e-4; I.q11 = 1.0; I.tstop_next = 1440.0;
    I.dt = determine_initdt(I, p, cn);
}

/* one accepted step: SciMLBase.step!(integrator) = loopheader!/perform_step!/loopfooter! until accepted */
static void step(Integrator& I, const RhsParams& p, Counters& cn)
{
    MCO_REGION(R_STEP);
    using namespace tsit5;
    const int M = I.M;
    double k2[MAXM], k3[MAXM], k4[MAXM], k5[MAXM], k6[MAXM], tmp[MAXM], unew[MAXM];
    int guard = 0;
    for (;;) {
        /* loopheader! */
        if (I.iter > 0) {
            if (!I.accept_step) {
                I.dt = I.dt / std::min(5.0, I.q11 / 0.9);                 /* step_reject_controller! */
            } else {
                for (int i = 0; i < M; i++) I.uprev[i] = I.u[i];          /* apply_step!: update_uprev! */
                I.dt = I.dtpropose;
                if (I.u_modified) soil_energy_balance(I.u, p, I.t, I.fsalfirst, cn);   /* reset_fsal! */
                else for (int i = 0; i < M; i++) I.fsalfirst[i] = I.fsallast[i];
            }
        } else if (I.u_modified) {
            for (int i = 0; i < M; i++) I.uprev[i] = I.u[i];
        }
        I.iter++;
        I.dt = std::min(1440.0, I.dt);                                    /* fix_dt_at_bounds! */
        I.dt = std::min(I.dt, I.tstop_next - I.t);                        /* modify_dt_for_tstops! */
        if (!(I.dt > 1e-12) || ++guard > 100000) {                          /* dt<=dtmin / NaN / maxiters: flag and bail out of the hour */
            I.flags |= (I.dt != I.dt) ? MC_STATUS_NAN : (guard > 100000 ? MC_STATUS_STEP_LIMIT : MC_STATUS_DT_UNDERFLOW);
            I.t = I.tstop_next; I.accept_step = true; I.dtpropose = 1.0;
            for (int i = 0; i < M; i++) I.fsallast[i] = I.fsalfirst[i];
            return;
        }
        /* perform_step!(integrator, ::Tsit5ConstantCache) */
        const double t = I.t, dt = I.dt; const double* k1 = I.fsalfirst; const double* up = I.uprev;
        for (int i = 0; i < M; i++) tmp[i] = up[i] + dt * a21 * k1[i];
        soil_energy_balance(tmp, p, t + c1 * dt, k2, cn);
        for (int i = 0; i < M; i++) tmp[i] = up[i] + dt * (a31 * k1[i] + a32 * k2[i]);
        soil_energy_balance(tmp, p, t + c2 * dt, k3, cn);
        for (int i = 0; i < M; i++) tmp[i] = up[i] + dt * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
        soil_energy_balance(tmp, p, t + c3 * dt, k4, cn);
        for (int i = 0; i < M; i++) tmp[i] = up[i] + dt * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
        soil_energy_balance(tmp, p, t + c4 * dt, k5, cn);
        for (int i = 0; i < M; i++) tmp[i] = up[i] + dt * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
        soil_energy_balance(tmp, p, t + dt, k6, cn);
        for (int i = 0; i < M; i++) unew[i] = up[i] + dt * (a71 * k1[i] + a72 * k2[i] + a73 * k3[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
        soil_energy_balance(unew, p, t + dt, I.fsallast, cn);
        const double* k7 = I.fsallast;
        for (int i = 0; i < M; i++) {
            const double ut = dt * (bt1 * k1[i] + bt2 * k2[i] + bt3 * k3[i] + bt4 * k4[i] + bt5 * k5[i] + bt6 * k6[i] + bt7 * k7[i]);
            tmp[i] = ut / (I.abstol + std::max(std::fabs(up[i]), std::fabs(unew[i])) * I.reltol);
        }
        I.EEst = rms_norm(tmp, M);
        for (int i = 0; i < M; i++) I.u[i] = unew[i];
        /* loopfooter! */
        I.u_modified = false;
        double q;
        if (I.EEst == 0.0) q = 0.1;                                       /* stepsize_controller!(PIController) */
        else {
            I.q11 = fastpow(I.EEst, 7.0 / 50.0);
            q = I.q11 / fastpow(I.qold, 2.0 / 25.0);
            q = std::max(0.1, std::min(5.0, q / 0.9));
        }
        I.accept_step = I.EEst <= 1.0;
        if (I.EEst != I.EEst) { I.flags |= MC_STATUS_NAN; I.accept_step = true; q = 1.0; }
        if (I.accept_step) {
            cn.n_accept++;
            if (q == 1.0) q = 1.0;                                         /* qsteady_min = qsteady_max = 1 */
            I.qold = std::max(I.EEst, 1e-4);                              /* step_accept_controller! */
            const double dtnew = I.dt / q;
            double ttmp = I.t + I.dt;
            if (std::fabs(ttmp - I.tstop_next) < 100.0 * std::numeric_limits<double>::epsilon() * std::max(I.t, I.tstop_next)) ttmp = I.tstop_next; /* fixed_t_for_floatingpoint_error! */
            I.t = ttmp;
            I.dtpropose = std::min(1440.0, dtnew);                        /* calc_dt_propose! */
            return;
        }
        cn.n_reject++;
        for (int i = 0; i < M; i++) I.u[i] = I.uprev[i];                  /* rejected: u is recomputed from uprev next attempt */
    }
}

/* step_to_hour!, src/simulation.jl:888-903 */
static void step_to_hour(Integrator& I, int hour_index, const RhsParams& p, Counters& cn)
{
    const double target = hour_index * 60.0;
    I.tstop_next = target;                                                /* add_tstop! (1440 stays queued behind it) */
    const double T_hi = p.m->maximum_surface_temperature_K;
    while (I.t < target - 1e-10) {
        step(I, p, cn);
        bool changed = false;
        for (int i = 0; i < I.M; i++) { const double c = jclamp(I.u[i], T_LO, T_hi); if (c != I.u[i]) { I.u[i] = c; changed = true; } }
        if (changed) { for (int i = 0; i < I.M; i++) I.uprev[i] = I.u[i]; I.u_modified = true; }
    }
}

/* ------------------------------------------------------------------------------------------
 * phase_transition!(::PhaseTransitionLatentHeat), src/soil_balance/freezing/latent_heat.jl:21-112
 * ------------------------------------------------------------------------------------------ */
static void phase_transition(int N, const double* depths, double* T, const double* Tpast, double* sum_phase, const double* theta)
{
    const double tol = 1.0e-4;
    for (int i = 0; i < N; i++) {
        double mean, mean_past;
        if (i < N - 1) { mean = 0.5 * (T[i] + T[i + 1]); mean_past = 0.5 * (Tpast[i] + Tpast[i + 1]); }
        else { mean = T[i]; mean_past = Tpast[i]; }
        double mass = 0.0;
        if (theta[i] > 0.0) {
            if (i < N - 1) mass = (depths[i + 1] - depths[i]) * 1000.0 * theta[i];
            else mass = (depths[i] + 1.0 - depths[i]) * 1000.0 * theta[i];
        }
        const double max_latent = LATENT_HEAT_FUSION * mass;
        if (theta[i] <= 0.0 || mass <= 0.0) { sum_phase[i] = 0.0; continue; }
        if ((mean_past - T0C > tol) && (mean - T0C <= -tol)) {
            double q = (mean_past - mean) * mass * WATER_HEAT_CAPACITY;
            sum_phase[i] += q;
            if (sum_phase[i] >= max_latent) sum_phase[i] = max_latent;
            T[i] = T0C; if (i < N - 1) T[i + 1] = T0C;
        } else if ((mean_past - T0C < -tol) && (mean - T0C >= tol)) {
            double q = (mean - mean_past) * mass * WATER_HEAT_CAPACITY;
            sum_phase[i] -= q;
            if (sum_phase[i] <= 0.0) sum_phase[i] = 0.0;
            else { T[i] = T0C; if (i < N - 1) T[i + 1] = T0C; }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Campbell (1985) water balance.  infiltration_step!, src/soil_hydraulics/campbell.jl:129-361 with
 * both infiltration algorithms (infiltration_algorithm/{matric_potential,matric_flux_potential}.jl).
 * Buffers are 1-based in the reference; arrays here have N+2 entries and use the same 1-based
 * indices (entry 0 unused) so the index shift (buffer i <-> soil node i-1) reads identically.
 * The scratch persists across calls exactly like the reference's pre-allocated buffers.
 * ------------------------------------------------------------------------------------------ */
struct InfilBuffers {
    double psi[MAXN + 2], depth[MAXN + 2], mass[MAXN + 2], wc[MAXN + 2], wcn[MAXN + 2], K[MAXN + 2], h[MAXN + 2], Ts[MAXN + 2],
           psi_e[MAXN + 2], binv[MAXN + 2], nexp[MAXN + 2], ncomp[MAXN + 2], sat[MAXN + 2], rootR[MAXN + 2], rootZ[MAXN + 2],
           jv[MAXN + 2], djv[MAXN + 2], sub[MAXN + 2], dia[MAXN + 2], sup[MAXN + 2], cc[MAXN + 2], F[MAXN + 2], rr[MAXN + 2],
           dpsi[MAXN + 2], soilR[MAXN + 2], uptake[MAXN + 2], phi[MAXN + 2], fexp[MAXN + 2], phi_e[MAXN + 2];
    double psi_out[MAXN], h_out[MAXN];
    InfilBuffers() { std::memset(this, 0, sizeof(*this)); }
};
struct InfilOut { double evaporation, transpiration, surface_water_flux; int newton_iters; bool hit_max; };

static InfilOut infiltration_step(InfilBuffers& B, const mc_model_desc& m, const Site& s, int N, const double* depths,
                                  double P, double* theta, double rh_loc, double lai, double ET, const double* Tin, Counters& cn)
{
    MCO_REGION(R_INFIL);
    cn.n_infil++;
    const double Mw = 0.01801528, Dv = 2.4e-5, R = m.fluid.gas_constant, g = m.fluid.standard_gravity;
    const double dt = m.moisture_timestep_s; const int max_it = m.moisture_max_iterations; const double tolr = m.moisture_tolerance;
    const bool flux_alg = m.infiltration_algorithm == MC_INFIL_MATRIC_FLUX_POTENTIAL;
    const int L = N;   /* num_layers */
    for (int i = 1; i <= L; i++) {                                                         /* :180-199 */
        B.psi_e[i] = -std::fabs(s.psi_e[i - 1]);
        B.sat[i] = 1.0 - s.bd[i - 1] / s.md[i - 1];
        B.binv[i] = 1.0 / s.bb[i - 1];
        B.nexp[i] = 2.0 + 3.0 / s.bb[i - 1];
        B.ncomp[i] = 1.0 - B.nexp[i];
        B.fexp[i] = (2.0 * s.bb[i - 1] + 3.0) / (s.bb[i - 1] + 3.0);
        B.phi_e[i] = -s.ks[i - 1] * B.psi_e[i] * s.bb[i - 1] / (s.bb[i - 1] + 3.0);
    }
    B.psi_e[L + 1] = B.psi_e[L]; B.sat[L + 1] = B.sat[L]; B.binv[L + 1] = B.binv[L]; B.nexp[L + 1] = B.nexp[L];
    B.ncomp[L + 1] = B.ncomp[L]; B.fexp[L + 1] = B.fexp[L]; B.phi_e[L + 1] = B.phi_e[L];
    for (int i = 1; i <= L; i++) B.depth[i + 1] = depths[i - 1];                            /* :202-206 */
    B.depth[1] = 0.0; B.depth[2] = 0.0;
    for (int i = 1; i <= L; i++) B.Ts[i] = Tin[i - 1];                                      /* :209-212 */
    B.Ts[L + 1] = Tin[L - 1];
    for (int i = 2; i <= L; i++) { B.wcn[i] = theta[i - 2]; B.wc[i] = theta[i - 2]; }      /* :215-218 */
    auto bb = [&](int i) { return s.bb[i - 1]; };
    auto ks = [&](int i) { return s.ks[i - 1]; };
    /* initial_state_from_water_content! + state_to_water_potential! (:220-221) */
    if (!flux_alg) {
        for (int i = 2; i <= L; i++) {
            B.psi[i] = B.psi_e[i] * mcm::pow(B.sat[i] / B.wcn[i], bb(i));
            B.K[i] = ks(i) * mcm::pow(B.psi_e[i] / B.psi[i], B.nexp[i]);
        }
        B.psi[L + 1] = B.psi_e[L];
    } else {
        for (int i = 2; i <= L; i++) {
            B.phi[i] = B.phi_e[i] * mcm::pow(B.wcn[i] / B.sat[i], bb(i) + 3.0);
            B.K[i] = ks(i) * mcm::pow(B.phi[i] / B.phi_e[i], B.fexp[i]);
        }
        B.phi[L + 1] = B.phi_e[L];
        for (int i = 2; i <= L; i++) B.psi[i] = B.psi_e[i] * mcm::pow(B.sat[i] / B.wcn[i], bb(i));
        B.psi[L + 1] = B.psi_e[L];
    }
    for (int i = 2; i <= L; i++) B.h[i] = mcm::exp(Mw * B.psi[i] / (R * B.Ts[i - 1]));      /* :223-225 */
    for (int i = 2; i <= L; i++) B.mass[i] = WATER_DENSITY * (B.depth[i + 1] - B.depth[i - 1]) / 2.0;   /* :228-230 */
    B.h[L + 1] = 1.0; B.wc[L + 1] = B.sat[L + 1]; B.wcn[L + 1] = B.sat[L + 1];              /* :232-237 */
    B.depth[1] = -1e10; B.depth[L + 1] = 1e20; B.K[L + 1] = ks(L);
    for (int i = 2; i <= L; i++) {                                                          /* :240-248 */
        const double rd = s.rootd[i - 1];
        if (rd > 0.0) {
            B.rootR[i] = m.root_resistance / (rd * (B.depth[i + 1] - B.depth[i - 1]) / 2.0);
            B.rootZ[i] = B.ncomp[i] * mcm::log(M_PI * m.root_radius * m.root_radius * rd) / (4.0 * M_PI * rd * (B.depth[i + 1] - B.depth[i - 1]) / 2.0);
        } else { B.rootR[i] = 1e20; B.rootZ[i] = 0.0; }
    }
    B.psi[1] = B.psi[2]; B.K[1] = 0.0;                                                      /* :250-251 */
    const double EP = mcm::exp(-0.82 * lai) * ET;                                           /* :254-255 */
    const double TP = ET - EP;
    double psum = 0.0, rsum = 0.0, psi_leaf = 0.0;                                          /* :258-267 */
    for (int i = 2; i <= L; i++) {
        B.soilR[i] = B.rootZ[i] / B.K[i];
        psum += B.psi[i] / (B.soilR[i] + B.rootR[i]);
        rsum += 1.0 / (B.soilR[i] + B.rootR[i]);
    }
    const double psi_bar = psum / rsum, R_bar = 1.0 / rsum;
    const double RL = m.leaf_resistance, PC = m.stomatal_closure_potential, SP = m.stomatal_stability_parameter;
    int counter = 0;
    while (counter < max_it) {                                                              /* :270-284 */
        MCO_REGION(R_STOMATAL);
        if (psi_leaf > psi_bar) psi_leaf = psi_bar - TP * (R_bar + RL);
        const double xp = mcm::pow(psi_leaf / PC, SP);
        const double slope = TP * (R_bar + RL) * SP * xp / (psi_leaf * (1.0 + xp) * (1.0 + xp)) - 1.0;
        const double resid = psi_bar - psi_leaf - TP * (R_bar + RL) / (1.0 + xp);
        psi_leaf = psi_leaf - resid / slope;
        counter++;
        if (std::fabs(resid) <= 10.0) break;
    }
    const double xp = mcm::pow(psi_leaf / PC, SP);                                          /* :285-289 */
    const double TR = TP / (1.0 + xp);
    for (int i = 2; i <= L; i++) B.uptake[i] = (B.psi[i] - psi_leaf - RL * TR) / (B.rootR[i] + B.soilR[i]);

    InfilOut out; out.hit_max = true; out.newton_iters = 0;
    counter = 0;
    while (counter < max_it) {                                                              /* :293-336 */
        MCO_REGION(R_NEWTON);
        double err = 0.0;
        counter++; cn.n_newton++;
        for (int i = 2; i <= L; i++) {
            if (!flux_alg) B.K[i] = ks(i) * mcm::pow(B.psi_e[i] / B.psi[i], B.nexp[i]);
            else B.K[i] = ks(i) * mcm::pow(B.phi[i] / B.phi_e[i], B.fexp[i]);
        }
        if (flux_alg) { for (int i = 2; i <= L; i++) B.psi[i] = B.psi_e[i] * mcm::pow(B.sat[i] / B.wcn[i], bb(i)); B.psi[L + 1] = B.psi_e[L]; }
        B.jv[1] = EP * (B.h[2] - rh_loc) / (1.0 - rh_loc);                                  /* :301-302 */
        B.djv[1] = EP * Mw * B.h[2] / (R * B.Ts[1] * (1.0 - rh_loc));
        for (int i = 2; i <= L; i++) {                                                      /* :307-312 */
            const double rhov = wet_air(m.fluid, B.Ts[i], 1.0, P).vapour_density;
            const double kv = 0.66 * Dv * rhov * (B.sat[i] - (B.wcn[i] + B.wcn[i + 1]) / 2.0) / (B.depth[i + 1] - B.depth[i]);
            B.jv[i] = kv * (B.h[i + 1] - B.h[i]);
            B.djv[i] = Mw * B.h[i] * kv / (R * B.Ts[i - 1]);
        }
        for (int i = 2; i <= L; i++) {                                                      /* assemble_tridiagonal_system! */
            if (!flux_alg) {
                const double cap = -1.0 * B.mass[i] * B.wcn[i] / (bb(i) * B.psi[i] * dt);
                B.sub[i] = -1.0 * B.K[i - 1] / (B.depth[i] - B.depth[i - 1]) + g * B.nexp[i] * B.K[i - 1] / B.psi[i - 1];
                B.sup[i] = -1.0 * B.K[i + 1] / (B.depth[i + 1] - B.depth[i]);
                B.dia[i] = B.K[i] / (B.depth[i] - B.depth[i - 1]) + B.K[i] / (B.depth[i + 1] - B.depth[i]) + cap
                         - g * B.nexp[i] * B.K[i] / B.psi[i] + B.djv[i - 1] + B.djv[i];
                B.F[i] = ((B.psi[i] * B.K[i] - B.psi[i - 1] * B.K[i - 1]) / (B.depth[i] - B.depth[i - 1])
                        - (B.psi[i + 1] * B.K[i + 1] - B.psi[i] * B.K[i]) / (B.depth[i + 1] - B.depth[i])) / B.ncomp[i]
                       + B.mass[i] * (B.wcn[i] - B.wc[i]) / dt - g * (B.K[i - 1] - B.K[i]) + B.jv[i - 1] - B.jv[i] + B.uptake[i];
            } else {
                const double b3 = B.fexp[i];
                const double cap = B.mass[i] * B.wcn[i] / ((bb(i) + 3.0) * B.phi[i] * dt);
                B.sub[i] = -1.0 / (B.depth[i] - B.depth[i - 1]) - g * b3 * B.K[i - 1] / B.phi[i - 1];
                B.sup[i] = -1.0 / (B.depth[i + 1] - B.depth[i]);
                B.dia[i] = 1.0 / (B.depth[i] - B.depth[i - 1]) + 1.0 / (B.depth[i + 1] - B.depth[i]) + cap + g * b3 * B.K[i] / B.phi[i];
                B.F[i] = (B.phi[i] - B.phi[i - 1]) / (B.depth[i] - B.depth[i - 1]) - (B.phi[i + 1] - B.phi[i]) / (B.depth[i + 1] - B.depth[i])
                       + B.mass[i] * (B.wcn[i] - B.wc[i]) / dt - g * (B.K[i - 1] - B.K[i]) + B.jv[i - 1] - B.jv[i] + B.uptake[i];
            }
            err += std::fabs(B.F[i]);
        }
        for (int i = 2; i <= L - 1; i++) {                                                  /* thomas_forward_elimination!, abstract.jl:70-82 */
            B.cc[i] = B.sup[i] / B.dia[i];
            if (std::fabs(B.cc[i]) < 1e-8) B.cc[i] = 0.0;
            B.rr[i] = B.F[i] / B.dia[i];
            B.dia[i + 1] -= B.sub[i + 1] * B.cc[i];
            B.F[i + 1] -= B.sub[i + 1] * B.rr[i];
        }
        double* st = flux_alg ? B.phi : B.psi;
        const double* st_e = flux_alg ? B.phi_e : B.psi_e;
        auto clampflux = [&](int i) { if (B.phi[i] < 0.0) B.phi[i] = 1e-6; else if (B.phi[i] > B.phi_e[i]) B.phi[i] = B.phi_e[i]; };
        B.dpsi[L] = B.F[L] / B.dia[L];                                                      /* apply_lower_boundary_clamp! */
        st[L] -= B.dpsi[L];
        if (!flux_alg) st[L] = jmin(st[L], st_e[L]); else clampflux(L);
        for (int i = L - 1; i >= 2; i--) {                                                  /* apply_interior_clamp! */
            B.dpsi[i] = B.rr[i] - B.cc[i] * B.dpsi[i + 1];
            st[i] -= B.dpsi[i];
            if (!flux_alg) { if (st[i] > st_e[i]) st[i] = (st[i] + B.dpsi[i] + st_e[i]) / 2.0; }
            else clampflux(i);
        }
        if (!flux_alg) {                                                                    /* recover_water_content! */
            for (int i = 2; i <= L; i++) {
                B.wcn[i] = jmax(B.sat[i] * mcm::pow(B.psi_e[i] / B.psi[i], B.binv[i]), 1e-7);
                B.psi[i] = B.psi_e[i] * mcm::pow(B.sat[i] / B.wcn[i], bb(i));
            }
        } else {
            for (int i = 2; i <= L; i++) B.wcn[i] = jmax(B.sat[i] * mcm::pow(B.phi[i] / B.phi_e[i], 1.0 / (bb(i) + 3.0)), 1e-7);
            for (int i = 2; i <= L; i++) B.psi[i] = B.psi_e[i] * mcm::pow(B.sat[i] / B.wcn[i], bb(i));
            B.psi[L + 1] = B.psi_e[L];
        }
        for (int i = 2; i <= L; i++) B.h[i] = mcm::exp(Mw * B.psi[i] / (R * B.Ts[i - 1]));  /* :329-332 */
        B.h[L + 1] = B.h[L];
        if (err <= tolr) { out.hit_max = false; break; }
    }
    out.newton_iters = counter;
    out.surface_water_flux = ((B.psi[2] * B.K[2] - B.psi[3] * B.K[3]) / (B.ncomp[2] * (B.depth[3] - B.depth[2])) + g * B.K[2] + TR) * dt;  /* :338 */
    for (int i = 1; i <= L + 1; i++) B.wc[i] = B.wcn[i];                                    /* :339 */
    for (int i = 2; i <= L + 1; i++) theta[i - 2] = B.wcn[i];                               /* :340-342 */
    out.evaporation = EP * (B.h[2] - rh_loc) / (1.0 - rh_loc) * dt;                         /* :348 */
    out.transpiration = TR;
    for (int i = 0; i < N; i++) { B.psi_out[i] = B.psi[i + 2]; B.h_out[i] = B.h[i + 2]; }   /* :355-357 */
    return out;
}

/* _wet_surface_node!, campbell.jl:369-376 */
static double wet_surface_node(double* theta, double pool, double sat, double half_thickness)
{
    if (pool <= 0.0) return pool;
    const double cap_node = half_thickness * 1000.0;
    const double capacity = jmax(0.0, (sat - theta[0]) * cap_node);
    const double delivered = jmin(pool, capacity);
    theta[0] = jmin(sat, theta[0] + delivered / cap_node);
    return pool - delivered;
}

/* atmospheric_surface_profile! restricted to what soil_water_balance! consumes: air temperature at the
 * first height (heights[1]) plus the convective flux; monin_obukhov.jl:85-206.  L_prof is the buffer's own
 * warm start (:26,166-168). */
static void surface_profile(const mc_model_desc& m, double z0, const EnvInstant& e, double T_surf, double& L_prof,
                            double& T_loc, double& Qc, Counters& cn)
{
    const double z = m.reference_height_m, h = m.local_height_m, kappa = m.karman_constant, gamma = m.dyer_constant;
    const double T_ref = e.T_ref;
    const double lzr = mcm::log(z / z0);
    const double dT = T_ref - T_surf;
    const double rho_cp = calc_rho_cp((T_surf + T_ref) / 2.0);
    if (T_ref >= T_surf || e.zenith >= 90.0) {
        const double ustar = jmax(kappa * e.wind / lzr, 1.0e-6);
        Qc = convective_flux(rho_cp, dT, ustar, bulk_stanton(lzr), sublayer_stanton(z0, ustar));
        const double Tz0 = (T_ref * bulk_stanton(lzr) + T_surf * sublayer_stanton(z0, ustar)) / (bulk_stanton(lzr) + sublayer_stanton(z0, ustar));
        if (h < z0) T_loc = T_surf + (Tz0 - T_surf) * (h / z0);
        else if (h == z) T_loc = T_ref;
        else T_loc = Tz0 + (T_ref - Tz0) * mcm::log(h / z0) / lzr;
    } else {
        ObukhovOut o = calc_obukhov(T_ref, T_surf, e.wind, z0, z, kappa, dT, rho_cp, 16.0, 30, 1e-2, L_prof);
        cn.n_obukhov += o.iters;
        L_prof = o.L; Qc = o.Q;
        if (h < z0) T_loc = T_surf + (o.T_z0 - T_surf) * (h / z0);
        else if (h == z) T_loc = T_ref;
        else {
            const double x = calc_phi_m(h, gamma, o.L);
            const double psi_h2 = calc_psi_h(x);
            const double tl = jmax(mcm::log(h / z0) - psi_h2, 1e-6);
            T_loc = o.T_z0 + (T_ref - o.T_z0) * tl / (mcm::log(z / z0) - o.psi_h);
        }
    }
}

/* soil_water_balance!, campbell.jl:378-495 */
struct WaterOut { double soil_wetness, pool; bool have_infil; };
static WaterOut soil_water_balance(InfilBuffers& B, const mc_model_desc& m, const Site& s, int N, const EnvInstant& e,
                                   const double* T0, double pool, double soil_wetness, double* theta, double& L_prof,
                                   bool snow_present, int& flags, Counters& cn)
{
    const mc_fluid_constants& fc = m.fluid;
    const double T_surf = T0[0], T_air = e.T_ref, P = e.P, rh = e.rh;
    double T_loc, Qc;
    surface_profile(m, s.z0, e, T_surf, L_prof, T_loc, Qc, cn);                              /* :408 */
    /* relative humidity at the reference height as the profile returns it (monin_obukhov.jl:196-197) */
    const double e_ref = wet_air(fc, T_air, rh, P).vapour_pressure;
    const double rh_ref_prof = jclamp(e_ref / goff_gratch(T_air) * 1.0, 0.0, 1.0);
    const WetAir wref = wet_air(fc, T_air, rh_ref_prof, P);                                  /* :416 */
    const WetAir wloc = wet_air(fc, T_loc, 1.0, P);                                          /* :417 */
    const double rh_loc = jclamp(wref.vapour_pressure / wloc.vapour_pressure, 0.0, 0.99);    /* :418 */
    const double hc = jmax(std::fabs(Qc / (T_surf - T_air)), 0.5);                           /* :419 */
    const WetAir wa = wet_air(fc, T_air, rh, P);
    const double hd = (hc / (wa.specific_heat * wa.density)) * mcm::pow(0.71 / 0.60, 0.666); /* :423 */
    const double Qe = evaporation_Q(fc, T_surf, T_air, rh, 1.0, hd, P, soil_wetness, true);  /* :424-434 */
    const double Lv = enthalpy_of_vaporisation(fc, T_surf);
    double EP = jmax(1e-7, Qe / Lv);                                                         /* :436 */
    if (snow_present) EP = 1e-7;
    const double sat = 1.0 - s.bd[0] / s.md[0];                                              /* :444-446 */
    const double half = (m.depths_m[1] - m.depths_m[0]) / 2.0;
    pool = wet_surface_node(theta, pool, sat, half);
    const int niter = (int)(3600.0 / m.moisture_timestep_s);
    InfilOut io{};
    for (int k = 0; k < niter; k++) {                                                        /* :449-484 */
        const uint64_t nw0 = cn.n_newton;
        io = infiltration_step(B, m, s, N, m.depths_m, P, theta, rh_loc, e.lai, EP, T0, cn);
        if (g_trace_sub && tl_trace_row >= 0 && k < 10) g_trace_sub[tl_trace_row * 10 + k] += (int32_t)(cn.n_newton - nw0);
        if (io.hit_max) flags |= MC_STATUS_NEWTON_MAXITER;
        const double surf_evap = jmax(0.0, io.evaporation);
        const double wflux = jmax(0.0, io.surface_water_flux);
        pool = jclamp(pool - wflux - surf_evap, 0.0, m.max_surface_pool);
        pool = wet_surface_node(theta, pool, sat, half);
    }
    WaterOut w;
    const double raw = jmax(0.0, io.evaporation);                                            /* :491-492 */
    w.soil_wetness = jclamp(raw / (EP * m.moisture_timestep_s), 0.0, 1.0);
    w.pool = pool; w.have_infil = true;
    return w;
}

/* ------------------------------------------------------------------------------------------
 * Per-site persistent state and driver: solve_soil!, src/simulation.jl:355-819
 * ------------------------------------------------------------------------------------------ */
struct SiteState {
    double T[MAXN], theta[MAXN], sum_phase[MAXN];
    double pool, soil_wetness, L_ode, L_prof, tdeep_prev;
    double fsal[MAXM];
    int flags;
    InfilBuffers infil;
    /* SnowModel only */
    double Ts[NSN]; SnowState snow; SnowScratch snow_scratch; double sd_last, Q_freeze;
    double L_air;        /* buffers.air_profile.obukhov_length_prev (solve_air!) */
};

static int32_t* g_trace_attempts = nullptr; static int32_t* g_trace_newton = nullptr; static int64_t g_trace_rows = 0;

struct SiteIO {
    /* inputs for the chunk, per-site gathered views (strided access through lambdas below) */
    int64_t S, s; int ndays;
    const mc_forcing_soa* f;
    mc_output_soa* o;
    int N;
    double hourly(const double* a, int step) const { return a[(int64_t)step * S + s]; }
    double daily(const double* a, int day) const { return a[(int64_t)day * S + s]; }
    void put_node(double* a, int row, int node, double v) const { if (a) a[((int64_t)row * N + node) * S + s] = v; }
    void put_scalar(double* a, int row, double v) const { if (a) a[(int64_t)row * S + s] = v; }
};

static EnvInstant get_instant(const SiteIO& io, int day, int step, const Site& site)
{
    const mc_forcing_soa& f = *io.f; EnvInstant e;
    e.lai = io.daily(f.leaf_area_index, day); e.shade = io.daily(f.shade, day);
    e.eps_surf = io.daily(f.surface_emissivity, day); e.eps_cloud = io.daily(f.cloud_emissivity, day);
    e.soil_wetness_daily = io.daily(f.soil_wetness, day); e.T_deep = io.daily(f.deep_soil_temperature, day);
    e.rain_daily = io.daily(f.rainfall_daily, day);
    e.has_albedo = f.albedo_daily != nullptr; e.albedo_daily = e.has_albedo ? io.daily(f.albedo_daily, day) : site.albedo;
    e.P = io.hourly(f.pressure, step); e.T_ref = io.hourly(f.reference_temperature, step);
    e.wind = io.hourly(f.reference_wind_speed, step); e.rh = io.hourly(f.reference_humidity, step);
    e.zenith = io.hourly(f.zenith_angle, step); e.cloud = io.hourly(f.cloud_cover, step);
    e.solar = io.hourly(f.global_radiation, step);
    e.has_lw = f.longwave_radiation != nullptr; e.lw = e.has_lw ? io.hourly(f.longwave_radiation, step) : 0.0;
    return e;
}

/* update_forcing_day!, src/simulation.jl:1006-1021 (25th knot repeats hour 24, :990) */
static void update_forcing_day(DayForcing& F, const SiteIO& io, int day)
{
    const mc_forcing_soa& f = *io.f;
    const double* src[F_COUNT];
    src[F_SOLAR] = f.global_radiation; src[F_ZENITH] = f.zenith_angle; src[F_SLOPE_ZENITH] = f.zenith_slope_angle;
    src[F_TEMPERATURE] = f.reference_temperature; src[F_WIND] = f.reference_wind_speed; src[F_HUMIDITY] = f.reference_humidity;
    src[F_CLOUD] = f.cloud_cover; src[F_PRESSURE] = f.pressure; src[F_DIFFUSE] = f.diffuse_fraction;
    for (int v = 0; v < F_COUNT; v++) {
        for (int h = 0; h < 24; h++) F.v[v][h] = io.hourly(src[v], day * 24 + h);
        F.v[v][24] = F.v[v][23];
    }
}

/* _tins_mean_K, src/simulation.jl:120-128 */
static double tins_mean(const SiteIO& io, int day)
{
    double ssum = 0.0;
    for (int h = 0; h < 24; h++) ssum += io.hourly(io.f->reference_temperature, day * 24 + h);
    ssum += io.hourly(io.f->reference_temperature, day * 24 + 23);
    return ssum / 25.0;
}

static void write_soil_properties(const SiteIO& io, const mc_model_desc& m, const Site& site, int N, int row,
                                  const double* T, const double* theta, double P)
{
    if (!io.o->soil_thermal_conductivity && !io.o->soil_heat_capacity && !io.o->soil_bulk_density) return;
    for (int i = 0; i < N; i++) {
        SoilProps sp = soil_properties(m, P, T[i], theta[i], site.bd[i], site.md[i], site.lam_m[i], site.c_m[i]);
        io.put_node(io.o->soil_thermal_conductivity, row, i, sp.conductivity);
        io.put_node(io.o->soil_heat_capacity, row, i, sp.heat_capacity);
        io.put_node(io.o->soil_bulk_density, row, i, sp.density);
    }
}

static bool is_converged(const mc_model_desc& m, int iter, int niter, const double* T0, const double* T0_prev, int N)
{
    if (m.convergence_kind == MC_CONV_FIXED) return iter >= niter;                           /* fixed.jl:12 */
    if (iter >= niter) return true;                                                          /* tolerance.jl:16-22 */
    if (iter <= 1) return niter <= 1;
    double mx = 0.0; for (int i = 0; i < N; i++) mx = std::max(mx, std::fabs(T0[i] - T0_prev[i]));
    return mx < m.convergence_tolerance_K;
}

/* fresh state == what init() + the top of solve_soil! establish (src/simulation.jl:141-243,383-452) */
static void init_site_state(SiteState& st, const mc_model_desc& m, const Site& site, const mc_initial_soa& init,
                            int64_t S, int64_t s)
{
    const int N = m.n_soil;
    for (int i = 0; i < N; i++) {
        st.T[i] = init.initial_soil_temperature ? init.initial_soil_temperature[(int64_t)i * S + s] : std::numeric_limits<double>::quiet_NaN();
        st.theta[i] = init.initial_soil_moisture[(int64_t)i * S + s];
        st.sum_phase[i] = 0.0;
    }
    st.pool = 0.0; st.soil_wetness = 0.0; st.L_ode = -0.3; st.L_prof = -0.3;
    st.tdeep_prev = std::numeric_limits<double>::quiet_NaN();
    for (int i = 0; i < MAXM; i++) st.fsal[i] = std::numeric_limits<double>::quiet_NaN();
    st.flags = 0;
    for (int i = 0; i < NSN; i++) { st.Ts[i] = m.snow.initial_snow_temperature_K; st.snow_scratch.node_depths[i] = 0.0; st.snow_scratch.meanT[i] = 0.0; st.snow_scratch.meanTpast[i] = 0.0; }
    st.snow = initial_snow_state(m.snow, m.snow.initial_snow_depth_cm, m.snow.initial_snow_density);
    st.snow_scratch.sd_prev = 0.0; st.snow_scratch.sd_cur = 0.0; st.sd_last = 0.0; st.Q_freeze = 0.0;
    st.L_air = -0.3;
    (void)site;
}

/*
 * solve_soil! for days [day0, day0+ndays) of one site.  day0 == 0 runs the pre-loop block (src/simulation.jl:395-452) and the
 * prototype RHS evaluation that SciMLBase.init performs to seed fsalfirst (src/simulation.jl:224-238, 853-860).
 * With SnowModel the ODE state is [T_snow(8); T_soil(N)] over hourly-changing effective depths (src/simulation.jl:595-741).
 */
static void solve_soil_site(const mc_model_desc& m, const Site& site, const mc_initial_soa& init, SiteState& st,
                            SiteIO& io, int64_t day0, Counters& cn)
{
    const int N = m.n_soil; const int ns = m.n_snow; const int M = N + ns;
    const bool snow = ns > 0; const mc_snow_desc& sm = m.snow;
    const int ndays = io.ndays; const int nrows = ndays * 24;
    const bool dynamic = m.moisture_strategy == MC_MOIST_DYNAMIC;
    const bool nonconsec = m.time_mode == MC_TIME_NONCONSECUTIVE;
    const bool hourly_rain = m.rainfall_schedule == MC_RAIN_HOURLY;
    const mc_forcing_soa& f = *io.f; mc_output_soa& o = *io.o;
    const bool has_init_T = init.initial_soil_temperature != nullptr;
    double T0[MAXN]; for (int i = 0; i < N; i++) T0[i] = st.T[i];
    double* theta = st.theta; double* sum_phase = st.sum_phase;
    double* Ts = st.Ts; SnowState& sst = st.snow; SnowScratch& ssc = st.snow_scratch;
    double sum_phase_day_start[MAXN], sat[MAXN];
    for (int i = 0; i < N; i++) sat[i] = 1.0 - site.bd[i] / site.md[i];
    double pool = st.pool, Q_freeze = st.Q_freeze;
    DayForcing F;
    RhsParams p; p.m = &m; p.site = &site; p.forcing = &F; p.M = M; p.n_snow = ns; p.snow = &sst; p.snow_scratch = &ssc;
    for (int i = 0; i < N; i++) p.depths[ns + i] = m.depths_m[i];
    for (int i = 0; i < ns; i++) p.depths[i] = 0.0;
    Integrator I; I.M = M; I.reltol = m.reltol; I.abstol = m.abstol; I.flags = 0; I.iter = 0; I.accept_step = false; I.u_modified = false;
    I.t = 0; I.dt = 0; I.dtpropose = 0; I.qold = 1e-4; I.q11 = 1; I.EEst = 1; I.tstop_next = 1440.0;
    for (int i = 0; i < M; i++) { I.fsalfirst[i] = st.fsal[i]; I.fsallast[i] = st.fsal[i]; }
    auto put_water_potential_row = [&](int row) {
        for (int i = 0; i < N; i++) io.put_node(o.soil_water_potential, row, i, site.psi_e[i] * mcm::pow(sat[i] / theta[i], site.bb[i]));
    };
    auto init_soil_humidity = [&](const double* Tn) {   /* initialise_soil_humidity!, dynamic.jl:32-35 — always row 1 of the buffer */
        if (!dynamic) return;
        for (int i = 0; i < N; i++) {
            const double psi = site.psi_e[i] * mcm::pow(sat[i] / theta[i], site.bb[i]);
            io.put_node(o.soil_humidity, 0, i, jclamp(mcm::exp(0.01801528 * psi / (m.fluid.gas_constant * Tn[i])), 0.0, 1.0));
        }
    };
    auto combine = [&](double* u) { for (int i = 0; i < ns; i++) u[i] = Ts[i]; for (int i = 0; i < N; i++) u[ns + i] = T0[i]; };
    auto sync_inactive = [&]() { const double sy = Ts[0]; for (int k = 0; k < ns; k++) if (!(ssc.node_depths[k] > 0.0)) Ts[k] = sy; };   /* simulation.jl:93-98 */
    EnvInstant env;
    if (day0 == 0) {
        /* init(): prototype RHS evaluation seeds fsalfirst (SciMLBase.init -> initialize!), src/simulation.jl:196-238 */
        update_forcing_day(F, io, 0);
        env = get_instant(io, 0, 0, site);
        double Tproto[MAXM];
        for (int i = 0; i < ns; i++) Tproto[i] = sm.initial_snow_temperature_K;
        for (int i = 0; i < N; i++) { Tproto[ns + i] = has_init_T ? st.T[i] : 280.0; p.theta[i] = theta[i]; }
        p.shade = env.shade; p.eps_surf = env.eps_surf; p.eps_cloud = env.eps_cloud;
        p.soil_wetness = 0.0; p.albedo = site.albedo; p.Q_freeze = 0.0; p.L_prev = st.L_ode;
        SnowState proto_state = initial_snow_state(sm, 0.0, -1.0);
        if (snow) {                                            /* depths_proto = combined_depths(initial snow state) but p.snow_state = initial_snow_state(model) */
            compute_effective_depths(sm, sst, ssc, m.depths_m, N, p.depths);
            p.snow = &proto_state;
        }
        Counters scratch;
        soil_energy_balance(Tproto, p, 0.0, I.fsalfirst, scratch);
        p.snow = &sst;
        /* solve_soil! pre-loop, :395-452 */
        if (!has_init_T) { const double t_init = tins_mean(io, 0); for (int i = 0; i < N; i++) T0[i] = t_init; }
        for (int i = 0; i < N; i++) sum_phase[i] = 0.0;
        write_soil_properties(io, m, site, N, 0, T0, theta, 101325.0);
        put_water_potential_row(0);
        for (int i = 0; i < N; i++) { io.put_node(o.soil_temperature, 0, i, T0[i]); io.put_node(o.soil_moisture, 0, i, theta[i]); }
        init_soil_humidity(T0);
        st.tdeep_prev = env.T_deep;
        pool = 0.0; Q_freeze = 0.0;
    }
    double tdeep_stale = st.tdeep_prev;
    double soil_wetness_dyn = st.soil_wetness;
    double sky_T = 0.0;
    double sd_last = st.sd_last;      /* snow_depth_hourly[last processed step] */

    for (int jl = 0; jl < ndays; jl++) {
        const int64_t j = day0 + jl;   /* 0-based global day */
        const int doy = f.day_of_year ? f.day_of_year[jl] : (int)(j + 1);
        update_forcing_day(F, io, jl);
        int niter;
        if (nonconsec) niter = m.iterations_per_day;                                         /* non_consecutive.jl:46-47 */
        else niter = (m.spinup_first_day && j == 0) ? m.convergence_iterations : 1;          /* consecutive.jl:21-23 */
        int noon_i = 0; double best = io.hourly(f.zenith_angle, jl * 24);                    /* :463-473 */
        for (int i = 1; i < 24; i++) { const double z = io.hourly(f.zenith_angle, jl * 24 + i); if (z < best) { best = z; noon_i = i; } }
        const int midnight_i = (noon_i + 12) % 24;   /* 0-based hour index */
        if (nonconsec) {                                                                     /* :480-516 */
            for (int i = 0; i < N; i++) sum_phase[i] = 0.0;
            if (snow) sst.sum_phase = 0.0;
            if (!dynamic && f.prescribed_soil_moisture) { for (int i = 0; i < N; i++) theta[i] = f.prescribed_soil_moisture[((int64_t)jl * N + i) * io.S + io.s]; }
            else for (int i = 0; i < N; i++) theta[i] = init.initial_soil_moisture[(int64_t)i * io.S + io.s];
            /* reset_snow_per_day is false in BOTH time modes (non_consecutive.jl:40): the snowpack persists from one output day to the next */
            const double t_reset = tins_mean(io, jl);
            if (!has_init_T) for (int i = 0; i < N; i++) T0[i] = t_reset;
            else for (int i = 0; i < N; i++) T0[i] = init.initial_soil_temperature[(int64_t)i * io.S + io.s];
            for (int i = 0; i < ns; i++) Ts[i] = t_reset;
        }
        T0[N - 1] = tdeep_stale;                                                             /* :517 */
        int iter = 0; bool is_last = false;
        double T0_prev_start[MAXN], T0_day_start[MAXN], T0_iter_start[MAXN], T0_before[MAXN], T0_output[MAXN];
        double Ts_day_start[NSN], Ts_before[NSN];
        for (int i = 0; i < N; i++) { T0_prev_start[i] = T0[i]; T0_day_start[i] = T0[i]; sum_phase_day_start[i] = sum_phase[i]; }
        for (int i = 0; i < ns; i++) Ts_day_start[i] = Ts[i];
        while (!is_last) {                                                                   /* :530 */
            iter++; cn.n_day_passes++;
            for (int i = 0; i < N; i++) T0_iter_start[i] = T0[i];
            is_last = is_converged(m, iter, niter, T0, T0_prev_start, N);
            for (int i = 0; i < N; i++) T0_prev_start[i] = T0_iter_start[i];
            if (!nonconsec) { for (int i = 0; i < N; i++) T0[i] = T0_day_start[i]; for (int i = 0; i < ns; i++) Ts[i] = Ts_day_start[i]; }   /* iter_resets_T */
            for (int i = 0; i < N; i++) { sum_phase[i] = sum_phase_day_start[i]; T0_output[i] = T0[i]; }
            if (snow && nonconsec) sst.sum_phase = 0.0;
            const int day_row = jl * 24;
            if (is_last) {                                                                   /* :557-592 */
                for (int i = 0; i < N; i++) { io.put_node(o.soil_temperature, day_row, i, T0[i]); io.put_node(o.soil_moisture, day_row, i, theta[i]); }
                io.put_scalar(o.surface_water, day_row, pool);
                put_water_potential_row(day_row);
                init_soil_humidity(T0);
                EnvInstant e0 = get_instant(io, jl, day_row, site);
                io.put_scalar(o.sky_temperature, day_row, sky_temperature(m, site.svf, e0.P, e0.rh, e0.T_ref, e0.eps_cloud, e0.cloud, e0.shade, e0.has_lw, e0.lw));
                write_soil_properties(io, m, site, N, day_row, T0, theta, e0.P);
            }
            int step = jl * 24;                                                              /* chunk-relative 0-based; global 1-based = day0*24 + step + 1 */
            env = get_instant(io, jl, step, site);
            T0[N - 1] = env.T_deep;
            /* apply_snow_overrides (simulation.jl:46-81) + per-hour parameter update */
            auto set_hour_params = [&](const EnvInstant& e, int64_t gstep /*1-based global*/, double depth_prev_hour) {
                p.shade = e.shade; p.eps_surf = e.eps_surf; p.eps_cloud = e.eps_cloud;
                const bool snow_active = snow && is_last && ((gstep > 1 ? depth_prev_hour : sst.depth) > 0.0);
                if (snow_active) { p.albedo = snow_albedo(sst.days_since); p.soil_wetness = 1.0; }
                else { p.albedo = e.has_albedo ? e.albedo_daily : site.albedo; p.soil_wetness = dynamic ? soil_wetness_dyn : e.soil_wetness_daily; }
                p.Q_freeze = Q_freeze;
                sky_T = sky_temperature(m, site.svf, e.P, e.rh, e.T_ref, e.eps_cloud, e.cloud, e.shade, e.has_lw, e.lw);
            };
            const int64_t g0 = day0 * 24 + step + 1;
            set_hour_params(env, g0, sd_last);
            if (snow) {
                if (!is_last) { double tmp = sst.depth; activate_snow_nodes(sm, sst, ssc, tmp); }
                compute_effective_depths(sm, sst, ssc, m.depths_m, N, p.depths);
                int fa = -1; for (int k = 0; k < ns; k++) if (ssc.node_depths[k] > 0.0) { fa = k; break; }
                const double sync_temp = fa < 0 ? T0[0] : Ts[fa];
                for (int k = 0; k < ns; k++) if (fa < 0 || k < fa) Ts[k] = sync_temp;     /* _sync_snow_to_active */
            }
            p.L_prev = st.L_ode;
            for (int i = 0; i < N; i++) p.theta[i] = theta[i];
            double u0[MAXM]; combine(u0);
            reinit(I, u0, p, cn);
            for (int i = 0; i < 24; i++) {                                                   /* :635 */
                step = jl * 24 + i;
                const int64_t gstep = day0 * 24 + step + 1;
                for (int k = 0; k < N; k++) { T0_before[k] = T0[k]; p.theta[k] = theta[k]; }
                for (int k = 0; k < ns; k++) Ts_before[k] = Ts[k];
                p.L_prev = st.L_ode;
                const uint64_t att0 = cn.n_accept + cn.n_reject, new0 = cn.n_newton;
                tl_trace_row = (g_trace_sub && step < g_trace_rows) ? io.s * g_trace_rows + step : -1;
                step_to_hour(I, i + 1, p, cn);
                if (g_trace_attempts && step < g_trace_rows) g_trace_attempts[io.s * g_trace_rows + step] += (int32_t)(cn.n_accept + cn.n_reject - att0);
                for (int k = 0; k < ns; k++) Ts[k] = I.u[k];
                for (int k = 0; k < N; k++) T0[k] = I.u[ns + k];
                if (is_last) {                                                               /* :645-650 */
                    phase_transition(N, m.depths_m, T0, T0_before, sum_phase, theta);
                    for (int k = 0; k < N; k++) T0_output[k] = T0[k];
                }
                double snowfall_hour = 0.0, thermal_melt = 0.0, rain_melt_snow = 0.0;
                if (snow && is_last) {                                                       /* :656-708 */
                    ssc.sd_prev = sd_last;
                    double Qe, Qc;                                                           /* zenith defaults to 90 deg: neutral branch */
                    surface_convection_evaporation(m, T0[0], env.T_ref, env.wind, env.rh, env.P, site.z0, m.reference_height_m, 90.0, p.soil_wetness, -0.3, Qe, Qc, cn);
                    const double rain_for_snow = env.rain_daily;                             /* environment_instant.rainfall is the DAILY total (simulation.jl:1031) */
                    SnowUpdateOut so = update_snow(m, sst, ssc, Ts, Ts_before, env.T_ref, rain_for_snow, env.P, gstep, T0[0], T0_before[0], Qe, hourly_rain, env.shade, doy);
                    snowfall_hour = so.snowfall; thermal_melt = so.thermal_melt; rain_melt_snow = so.rain_melt;
                    if (doy > 1 && sm.snow_melt_factor <= 1.0) Q_freeze = thermal_melt * 10.0 * LATENT_HEAT_FUSION / 3600.0 * (1.0 - sm.snow_melt_factor);
                    else Q_freeze = 0.0;
                    int prev_active = sst.active;
                    activate_snow_nodes(sm, sst, ssc, ssc.sd_cur);
                    if (sst.active < prev_active) sync_inactive();
                    double soil_surf = T0[0];
                    snow_phase_transition(m, sst, ssc, Ts, env.P, soil_surf);               /* soil-surface clamp deliberately NOT applied (simulation.jl:689-702) */
                    for (int k = 0; k < ns; k++) if (ssc.node_depths[k] > 0.0 && Ts[k] > T0C) Ts[k] = T0C;   /* clamp_snow_temperatures */
                    /* freeze_new_snow (snow_model.jl:235-264) selects the index range (N+1-m):(N-m), which is empty: a no-op upstream */
                    adjust_snow_near_nodes(sm, sst, ssc, gstep);
                    prev_active = sst.active;
                    activate_snow_nodes(sm, sst, ssc, ssc.sd_cur);
                    if (sst.active < prev_active) sync_inactive();
                    sd_last = ssc.sd_cur;
                }
                if (i < 23) {                                                                /* :712-735 */
                    env = get_instant(io, jl, step + 1, site);
                    T0[N - 1] = env.T_deep;
                    set_hour_params(env, gstep + 1, sd_last);
                    if (snow) {
                        if (!is_last) { double tmp = sst.depth; activate_snow_nodes(sm, sst, ssc, tmp); }
                        compute_effective_depths(sm, sst, ssc, m.depths_m, N, p.depths);
                    }
                }
                combine(I.u);                                                                /* :736-741 */
                I.u_modified = true;
                {                                                                            /* init_surface_fluxes!, monin_obukhov.jl:247-264 */
                    const Instant9 fi = interpolate_forcings(F, (double)(i * 60));
                    const double Tsurf = T0[0];
                    if (fi.air_temperature < Tsurf && fi.zenith < 90.0) {
                        const double dT = fi.air_temperature - Tsurf;
                        const double rho_cp = calc_rho_cp((Tsurf + fi.air_temperature) / 2.0);
                        const double L0 = st.L_ode >= 0.0 ? -0.3 : st.L_ode;
                        ObukhovOut ob = calc_obukhov(fi.air_temperature, Tsurf, fi.wind_speed, site.z0, m.reference_height_m, m.karman_constant, dT, rho_cp, 16.0, 30, 1e-2, L0);
                        cn.n_obukhov += ob.iters;
                        st.L_ode = ob.L;
                    }
                }
                bool have_infil = false;
                if (is_last) {                                                               /* :745-777 */
                    double rain;
                    if (hourly_rain) rain = io.hourly(f.rainfall_hourly, step);
                    else rain = (i == midnight_i) ? env.rain_daily : 0.0;
                    bool snow_present = false;
                    if (snow) {
                        rain = rain * sm.rain_multiplier;
                        snow_present = ssc.sd_cur > 0.0;
                        const double melted_water = thermal_melt * 10.0, rain_melt_water = rain_melt_snow * sst.density * 10.0;
                        if (snow_present && env.T_ref - T0C < sm.snow_temperature_threshold_C) pool = jclamp(pool + melted_water, 0.0, m.max_surface_pool);
                        else pool = jclamp(pool + rain + rain_melt_water + melted_water, 0.0, m.max_surface_pool);
                    } else {
                        pool = jclamp(pool + rain + 0.0 + 0.0, 0.0, m.max_surface_pool);
                    }
                    if (dynamic) {
                        WaterOut w = soil_water_balance(st.infil, m, site, N, env, T0, pool, soil_wetness_dyn, theta, st.L_prof, snow_present, st.flags, cn);
                        soil_wetness_dyn = w.soil_wetness; pool = w.pool; have_infil = true;
                    }
                }
                if (g_trace_newton && step < g_trace_rows) g_trace_newton[io.s * g_trace_rows + step] += (int32_t)(cn.n_newton - new0);
                const int out_row = step + 1;                                                /* :782-798 */
                const bool in_bounds = out_row < nrows;   /* rows past the chunk belong to the next chunk's day pre-fill */
                const bool crosses_day = (i == 23);
                const bool next_day_resets = crosses_day && nonconsec;   /* (j < ndays_total is implied by in_bounds) */
                if (is_last && in_bounds && !next_day_resets) {
                    io.put_scalar(o.surface_water, out_row, pool);
                    for (int k = 0; k < N; k++) { io.put_node(o.soil_temperature, out_row, k, T0_output[k]); io.put_node(o.soil_moisture, out_row, k, theta[k]); }
                    io.put_scalar(o.sky_temperature, out_row, sky_T);
                    if (have_infil) for (int k = 0; k < N; k++) { io.put_node(o.soil_water_potential, out_row, k, st.infil.psi_out[k]); io.put_node(o.soil_humidity, out_row, k, st.infil.h_out[k]); }
                    const double P_out = io.hourly(f.pressure, out_row);
                    write_soil_properties(io, m, site, N, out_row, T0, theta, P_out);
                }
                if (is_last && snow) {                                                       /* snow outputs are recorded at `step`, :803-813 */
                    io.put_scalar(o.snow_fall, step, snowfall_hour);
                    io.put_scalar(o.snow_depth, step, ssc.sd_cur);
                    io.put_scalar(o.snow_density, step, evolve_snow_density(sm, sst));
                    if (o.snow_temperature) for (int k = 0; k < ns; k++) o.snow_temperature[((int64_t)step * ns + k) * io.S + io.s] = Ts[k];
                }
            }
        }
        tdeep_stale = env.T_deep;
    }
    for (int i = 0; i < N; i++) st.T[i] = T0[i];
    st.pool = pool; st.soil_wetness = soil_wetness_dyn; st.tdeep_prev = tdeep_stale; st.sd_last = sd_last; st.Q_freeze = Q_freeze;
    for (int i = 0; i < M; i++) st.fsal[i] = I.fsalfirst[i];
    st.flags |= I.flags;
    for (int i = 0; i < N; i++) if (!(std::isfinite(st.T[i]) && std::isfinite(st.theta[i]))) st.flags |= MC_STATUS_NAN;
}


/* solve_air! (src/simulation.jl:822-851): per output row, atmospheric_surface_profile! (monin_obukhov.jl:85-206) at every height of
 * MicroModel.heights from the row's final surface temperature (top snow node when snow covers the ground).  Rows are processed in
 * order: the Obukhov warm start of the air-profile buffer carries from row to row. */
static void solve_air_site(const mc_model_desc& m, const Site& site, SiteState& st, const SiteIO& io, Counters& cn)
{
    const mc_forcing_soa& f = *io.f; const mc_output_soa& o = *io.o;
    if (!o.air_temperature && !o.wind_speed && !o.relative_humidity && !o.convective_heat_flux && !o.friction_velocity) return;
    const int N = m.n_soil, ns = m.n_snow, nh = m.n_heights, H = io.ndays * 24;
    const double z = m.reference_height_m, z0 = site.z0, kappa = m.karman_constant, gamma = m.dyer_constant;
    const double lzr = mcm::log(z / z0);
    int ref_idx = -1; for (int k = 0; k < nh; k++) if (m.heights_m[k] == z) { ref_idx = k; break; }
    auto put_h = [&](double* a, int row, int k, double v) { if (a) a[((int64_t)row * nh + k) * io.S + io.s] = v; };
    for (int row = 0; row < H; row++) {
        const bool has_snow = ns > 0 && o.snow_depth && o.snow_depth[(int64_t)row * io.S + io.s] > 0.0;
        const double T_surf = has_snow ? o.snow_temperature[((int64_t)row * ns + 0) * io.S + io.s] : o.soil_temperature[((int64_t)row * N + 0) * io.S + io.s];
        const double T_ref = io.hourly(f.reference_temperature, row), wind = io.hourly(f.reference_wind_speed, row);
        const double rh = io.hourly(f.reference_humidity, row), zen = io.hourly(f.zenith_angle, row);
        const double dT = T_ref - T_surf;
        const double rho_cp = calc_rho_cp((T_surf + T_ref) / 2.0);
        double Tair[MC_MAX_HEIGHTS], Qc, ustar;
        if (T_ref >= T_surf || zen >= 90.0) {
            ustar = jmax(kappa * wind / lzr, 1.0e-6);
            const double Stb = bulk_stanton(lzr), Sts = sublayer_stanton(z0, ustar);
            Qc = convective_flux(rho_cp, dT, ustar, Stb, Sts);
            const double Tz0 = (T_ref * Stb + T_surf * Sts) / (Stb + Sts);
            for (int k = 0; k < nh; k++) {
                const double h = m.heights_m[k]; double v;
                if (h < z0) { v = 0.0; Tair[k] = T_surf + (Tz0 - T_surf) * (h / z0); }
                else if (k == ref_idx) { v = wind; Tair[k] = T_ref; }
                else { v = (ustar / kappa) * mcm::log(h / z0 + 0.0); Tair[k] = Tz0 + (T_ref - Tz0) * mcm::log(h / z0) / lzr; }
                put_h(o.wind_speed, row, k, v);
            }
        } else {
            ObukhovOut ob = calc_obukhov(T_ref, T_surf, wind, z0, z, kappa, dT, rho_cp, 16.0, 30, 1e-2, st.L_air);
            cn.n_obukhov += ob.iters;
            st.L_air = ob.L; Qc = ob.Q; ustar = ob.ustar;
            for (int k = 0; k < nh; k++) {
                const double h = m.heights_m[k]; double v;
                if (h < z0) { v = 0.0; Tair[k] = T_surf + (ob.T_z0 - T_surf) * (h / z0); }
                else if (k == ref_idx) { v = wind; Tair[k] = T_ref; }
                else {
                    const double x = calc_phi_m(h, gamma, ob.L);
                    const double psi_m1 = calc_psi_m(x), psi_h2 = calc_psi_h(x);
                    const double hr = h / z0;
                    v = (ob.ustar / kappa) * jmax(mcm::log(hr) - psi_m1, 1e-6);
                    Tair[k] = ob.T_z0 + (T_ref - ob.T_z0) * jmax(mcm::log(hr) - psi_h2, 1e-6) / (mcm::log(z / z0) - ob.psi_h);
                }
                put_h(o.wind_speed, row, k, v);
            }
        }
        const double e_ref = goff_gratch(T_ref) * rh;
        for (int k = 0; k < nh; k++) {
            put_h(o.air_temperature, row, k, Tair[k]);
            put_h(o.relative_humidity, row, k, jclamp(e_ref / goff_gratch(Tair[k]) * 1.0, 0.0, 1.0));
        }
        io.put_scalar(o.convective_heat_flux, row, Qc);
        io.put_scalar(o.friction_velocity, row, ustar);
    }
}

static void gather_site(Site& site, const mc_model_desc& m, const mc_sites_soa& ss, int64_t S, int64_t s)
{
    const int N = m.n_soil;
    site.z0 = ss.roughness_height[s]; site.slope = ss.slope[s]; site.svf = ss.sky_view_fraction[s]; site.albedo = ss.albedo[s];
    for (int i = 0; i < N; i++) {
        const int64_t k = (int64_t)i * S + s;
        site.bd[i] = ss.bulk_density[k]; site.md[i] = ss.mineral_density[k]; site.lam_m[i] = ss.mineral_conductivity[k];
        site.c_m[i] = ss.mineral_heat_capacity[k]; site.psi_e[i] = ss.air_entry_water_potential[k];
        site.ks[i] = ss.saturated_hydraulic_conductivity[k]; site.bb[i] = ss.campbell_b_parameter[k]; site.rootd[i] = ss.root_density[k];
    }
}

struct Oracle {
    mc_model_desc model; int64_t S;
    std::vector<Site> sites; std::vector<SiteState> states;
    std::vector<double> init_T, init_theta; bool has_init_T;
    int64_t days_done; Counters cn;
};

} // namespace mco

/* ------------------------------------------------------------------------------------------
 * C entry points (ctypes / bench).  mco_* mirrors mc_* of the product ABI so tests feed
 * both sides the very same buffers.
 * ------------------------------------------------------------------------------------------ */
extern "C" {

void* mco_create(const mc_model_desc* model, int64_t n_sites)
{
    if (!model || n_sites <= 0 || model->n_soil < 3 || model->n_soil > MC_MAX_SOIL_NODES || (model->n_snow != 0 && model->n_snow != MC_SNOW_NODES)) return nullptr;
    mco::Oracle* o = new mco::Oracle();
    o->model = *model; o->S = n_sites; o->sites.resize(n_sites); o->states.resize(n_sites); o->days_done = 0; o->has_init_T = false;
    return o;
}
void mco_destroy(void* h) { delete (mco::Oracle*)h; }

int mco_set_sites(void* h, const mc_sites_soa* sites, const mc_initial_soa* init)
{
    mco::Oracle* o = (mco::Oracle*)h; const int N = o->model.n_soil; const int64_t S = o->S;
    o->has_init_T = init->initial_soil_temperature != nullptr;
    if (o->has_init_T) o->init_T.assign(init->initial_soil_temperature, init->initial_soil_temperature + (size_t)N * S);
    o->init_theta.assign(init->initial_soil_moisture, init->initial_soil_moisture + (size_t)N * S);
    mc_initial_soa ii; ii.initial_soil_temperature = o->has_init_T ? o->init_T.data() : nullptr; ii.initial_soil_moisture = o->init_theta.data();
    for (int64_t s = 0; s < S; s++) { mco::gather_site(o->sites[s], o->model, *sites, S, s); mco::init_site_state(o->states[s], o->model, o->sites[s], ii, S, s); }
    o->days_done = 0; o->cn = mco::Counters();
    return 0;
}

/* n_threads <= 0: all cores (OpenMP) */
int mco_step_days(void* h, int64_t day0, int32_t ndays, const mc_forcing_soa* forcing, mc_output_soa* out, int n_threads)
{
    mco::Oracle* o = (mco::Oracle*)h; const int64_t S = o->S;
    if (day0 != o->days_done) return -2;
    mc_initial_soa ii; ii.initial_soil_temperature = o->has_init_T ? o->init_T.data() : nullptr; ii.initial_soil_moisture = o->init_theta.data();
    mco::Counters total;
    int nbad = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
#endif
    {
        mco::Counters cn;
        int bad = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t s = 0; s < S; s++) {
            mco::SiteIO io; io.S = S; io.s = s; io.ndays = ndays; io.f = forcing; io.o = out; io.N = o->model.n_soil;
            mco::solve_soil_site(o->model, o->sites[s], ii, o->states[s], io, day0, cn);
            mco::solve_air_site(o->model, o->sites[s], o->states[s], io, cn);
            if (out->status) out->status[s] = o->states[s].flags;
            if (o->states[s].flags) bad++;
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            total.n_rhs += cn.n_rhs; total.n_accept += cn.n_accept; total.n_reject += cn.n_reject; total.n_obukhov += cn.n_obukhov;
            total.n_newton += cn.n_newton; total.n_infil += cn.n_infil; total.n_day_passes += cn.n_day_passes; nbad += bad;
        }
    }
    o->cn.n_rhs += total.n_rhs; o->cn.n_accept += total.n_accept; o->cn.n_reject += total.n_reject; o->cn.n_obukhov += total.n_obukhov;
    o->cn.n_newton += total.n_newton; o->cn.n_infil += total.n_infil; o->cn.n_day_passes += total.n_day_passes;
    o->days_done += ndays;
    return nbad;
}

int mco_get_counters(void* h, mc_counters* c)
{
    mco::Oracle* o = (mco::Oracle*)h; std::memset(c, 0, sizeof(*c));
    c->n_rhs = o->cn.n_rhs; c->n_accept = o->cn.n_accept; c->n_reject = o->cn.n_reject; c->n_obukhov_iter = o->cn.n_obukhov;
    c->n_newton_iter = o->cn.n_newton; c->n_infil_steps = o->cn.n_infil; c->n_day_passes = o->cn.n_day_passes;
    c->n_site_days = (uint64_t)(o->days_done * o->S);
    return 0;
}

int mco_get_state(void* h, mc_state_soa* out)
{
    mco::Oracle* o = (mco::Oracle*)h; const int N = o->model.n_soil; const int64_t S = o->S;
    for (int64_t s = 0; s < S; s++) {
        const mco::SiteState& st = o->states[s];
        for (int i = 0; i < N; i++) {
            if (out->soil_temperature) out->soil_temperature[(int64_t)i * S + s] = st.T[i];
            if (out->soil_moisture) out->soil_moisture[(int64_t)i * S + s] = st.theta[i];
            if (out->sum_phase) out->sum_phase[(int64_t)i * S + s] = st.sum_phase[i];
            if (out->fsal) out->fsal[(int64_t)i * S + s] = st.fsal[i];
        }
        if (out->pool) out->pool[s] = st.pool;
        if (out->soil_wetness) out->soil_wetness[s] = st.soil_wetness;
        if (out->obukhov_ode) out->obukhov_ode[s] = st.L_ode;
        if (out->obukhov_profile) out->obukhov_profile[s] = st.L_prof;
        if (out->deep_temperature_prev) out->deep_temperature_prev[s] = st.tdeep_prev;
        if (out->obukhov_air) out->obukhov_air[s] = st.L_air;
        if (o->model.n_snow > 0) {
            for (int i = 0; i < o->model.n_snow; i++) {
                if (out->snow_temperature) out->snow_temperature[(int64_t)i * S + s] = st.Ts[i];
                if (out->snow_node_depths) out->snow_node_depths[(int64_t)i * S + s] = st.snow_scratch.node_depths[i];
                if (out->fsal) out->fsal[(int64_t)(N + i) * S + s] = st.fsal[N + i];
            }
            if (out->snow_state) {
                const double v[MC_SNOW_STATE_FIELDS] = {st.snow.depth, st.snow.age, st.snow.days_since, st.snow.density, st.snow.prev_density, st.snow.cum_melt,
                                                        st.snow.extra, (double)st.snow.active, st.snow.sum_phase, st.sd_last};
                for (int k = 0; k < MC_SNOW_STATE_FIELDS; k++) out->snow_state[(int64_t)k * S + s] = v[k];
            }
            if (out->q_freeze) out->q_freeze[s] = st.Q_freeze;
        }
    }
    if (out->days_done) out->days_done[0] = o->days_done;
    if (out->status) for (int64_t s = 0; s < S; s++) out->status[s] = o->states[s].flags;
    return 0;
}

/* ---- leaf functions exported for known-answer tests against the reference's golden CSVs ---- */
double mco_goff_gratch(double T) { return mco::goff_gratch(T); }
void mco_wet_air(const mc_fluid_constants* fc, double T, double rh, double P, double* out5)
{
    mco::WetAir w = mco::wet_air(*fc, T, rh, P);
    out5[0] = w.vapour_pressure; out5[1] = w.vapour_density; out5[2] = w.specific_heat; out5[3] = w.density; out5[4] = w.mixing_ratio;
}
void mco_soil_properties(const mc_model_desc* m, double P, double T, double theta, double bd, double md, double lam_m, double c_m, double* out3)
{
    mco::SoilProps sp = mco::soil_properties(*m, P, T, theta, bd, md, lam_m, c_m);
    out3[0] = sp.conductivity; out3[1] = sp.heat_capacity; out3[2] = sp.density;
}
double mco_sky_temperature(const mc_model_desc* m, double svf, double P, double rh, double T_ref, double eps_cloud, double cloud, double shade)
{
    return mco::sky_temperature(*m, svf, P, rh, T_ref, eps_cloud, cloud, shade, false, 0.0);
}
/* profile at the local height: returns T_loc, wind_loc, Q_conv, L (atmospheric_surface_profile!) */
void mco_surface_profile(const mc_model_desc* m, double z0, double T_ref, double wind, double zenith, double T_surf, double L_prev, double* out4)
{
    mco::EnvInstant e{}; e.T_ref = T_ref; e.wind = wind; e.zenith = zenith;
    double L = L_prev, T_loc, Qc; mco::Counters cn;
    mco::surface_profile(*m, z0, e, T_surf, L, T_loc, Qc, cn);
    /* wind at the local height, monin_obukhov.jl:161,183-190 */
    const double z = m->reference_height_m, h = m->local_height_m, kappa = m->karman_constant;
    double v;
    if (T_ref >= T_surf || zenith >= 90.0) {
        const double ustar = mco::jmax(kappa * wind / mcm::log(z / z0), 1e-6);
        v = h < z0 ? 0.0 : (ustar / kappa) * mcm::log(h / z0 + 0.0);
    } else {
        mco::ObukhovOut o = mco::calc_obukhov(T_ref, T_surf, wind, z0, z, kappa, T_ref - T_surf, mco::calc_rho_cp((T_surf + T_ref) / 2.0), 16.0, 30, 1e-2, L_prev);
        const double x = mco::calc_phi_m(h, m->dyer_constant, o.L);
        v = h < z0 ? 0.0 : (o.ustar / kappa) * mco::jmax(mcm::log(h / z0) - mco::calc_psi_m(x), 1e-6);
    }
    out4[0] = T_loc; out4[1] = v; out4[2] = Qc; out4[3] = L;
}
/* white-box infiltration_step! (reference test/infiltration_algorithm_comparison.jl) */
int mco_infiltration_steps(const mc_model_desc* m, const mc_sites_soa* site1, double P, double rh_loc, double lai, double ET,
                           const double* T, double* theta, int nsteps)
{
    mco::Site s; mco::gather_site(s, *m, *site1, 1, 0);
    mco::InfilBuffers B; mco::Counters cn; int iters = 0;
    for (int k = 0; k < nsteps; k++) { mco::InfilOut io = mco::infiltration_step(B, *m, s, m->n_soil, m->depths_m, P, theta, rh_loc, lai, ET, T, cn); iters += io.newton_iters; }
    return iters;
}
double mco_fastpow(double x, double y) { return mco::fastpow(x, y); }
/* analysis hook: per-site per-hour RK attempts / Newton iterations of the NEXT mco_step_days call ([S][rows] int32, zero-initialised by the caller) */
void mco_set_trace_sub(int32_t* sub) { mco::g_trace_sub = sub; }
void mco_set_trace(int32_t* attempts, int32_t* newton, int64_t rows) { mco::g_trace_attempts = attempts; mco::g_trace_newton = newton; mco::g_trace_rows = rows; }
/* phase_transition! white box */
void mco_phase_transition(int N, const double* depths, double* T, const double* Tpast, double* sum_phase, const double* theta)
{
    mco::phase_transition(N, depths, T, Tpast, sum_phase, theta);
}

} /* extern "C" */
