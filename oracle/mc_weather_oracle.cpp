/*
 * mc_weather_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code) for the widening rows f2 / f3 of SURVEY 8:
 *     populate_weather!(output, mp, solar_radiation_out, scratch)      /root/reference/src/simulation.jl:20-36
 *       interpolate_minmax!                                            src/simulation.jl:306-331
 *         populate_quantities! / evaluate! / sample / covering_shape / endpoint_height / shape_value / frac
 *                                                                      src/inputs/diel.jl:316-408
 *       adjust_for_cloud_cover! -> shortwave_radiation!(AngstromMaxwellShortwave)
 *                                                                      src/simulation.jl:340-352, src/radiation/shortwave/angstrom_maxwell.jl:21-46
 *         sunshine_fraction(Angstrom)                                  src/radiation/sunshine_fraction/angstrom.jl:14
 *         calc_diffuse_fraction(ErbsDiffuseFraction)                   src/radiation/diffuse/erbs.jl:14-23
 *     solve_solar! (geometry outputs only)                             src/simulation.jl:291-304
 * one site, one day, one hour at a time, written the way the Julia is written (the diel engine re-resolves every endpoint for every
 * hour and recurses into seams exactly as diel.jl does).  Only tests/ may load this.
 *
 * PARITY STATUS
 *   diel engine, Angstrom-Maxwell, Erbs: restated from the reference's own source, line by line.  Not executable against Julia here
 *     (no Julia in the image), so "vs C++ restatement"; pinned against the reference's NicheMapR goldens (metout_monthly.csv TAREF / RH /
 *     VREF) at the reference's own test tolerances by tests/test_weather.py.
 *   solar geometry: PARITY UNPINNED AGAINST THE PACKAGE.  SolarRadiation.jl is an un-vendored dependency (reference Project.toml), so the
 *     geometry is the published McCullough & Porter (1971) algorithm as NicheMapR's SOLRAD implements it:
 *         omega = 2 pi / 365,  zeta = omega (d - 80) + 2 e (sin(omega d) - sin(80 omega)),  sin(delta) = 0.39784993 sin(zeta),
 *         cos Z = cos(phi) cos(delta) cos(h) + sin(phi) sin(delta),  h = 15 deg (t - t_noon),  H_sunrise = acos(-tan(delta) tan(phi)),
 *     with SOLRAD's near-horizon refraction (apparent zenith = Z - (31' - 7.5' per degree of altitude) for 88 < Z <= 90 degrees).  The
 *     constants and the refraction rule are ANCHORED ON THE REFERENCE'S GOLDEN VECTORS: the ZEN column of test/data/metout_monthly.csv
 *     (288 rows), metout_FordDryLake.csv (720 rows) and scan329/metout.csv (8760 rows) is reproduced to <= 3e-6 degrees
 *     (tests/test_weather.py).  The slope-zenith angle has no golden with slope != 0 and is the textbook spherical formula.
 *     The spectral clear-sky irradiance (SOLRAD's 111-wavelength integration) is NOT restated: its source and tables are absent.
 */
#include "../include/microclimate_b200.h"
#include "../include/mc_math.h"

#include <cmath>
#include <cstdint>
#include <limits>

namespace mcw_oracle {

static const double NaN = std::numeric_limits<double>::quiet_NaN();

/* ---- diel engine (src/inputs/diel.jl) ------------------------------------------------------------------------------------------- */
struct Day { double sunrise, sunset, solar_noon; };                                     /* solar_day, diel.jl:319-324 */

static double resolve_time(const mc_time_of_day& t, const Day& day)                    /* diel.jl:79-83 */
{
    switch (t.kind) {
    case MC_TOD_CLOCK: return t.value;
    case MC_TOD_SUNRISE: return day.sunrise + t.value;
    case MC_TOD_SUNSET: return day.sunset + t.value;
    case MC_TOD_MIDDAY: return day.solar_noon + t.value;
    default: return t.value;                                                            /* Midnight: the offset itself */
    }
}
static bool same_time(const mc_time_of_day& a, const mc_time_of_day& b) { return a.kind == b.kind && a.value == b.value; }   /* struct == */

static bool in_span(double h, double t1, double t2)                                     /* diel.jl:371-375 */
{
    const double t2c = t2 <= t1 ? t2 + 24 : t2;
    const double hc = h < t1 ? h + 24 : h;
    return t1 <= hc && hc < t2c;
}
static double frac(double h, double t1, double t2)                                      /* diel.jl:403-407 */
{
    const double t2c = t2 <= t1 ? t2 + 24 : t2;
    const double hc = h < t1 ? h + 24 : h;
    return (hc - t1) / (t2c - t1);
}
/* Julia mod(x, 24.0) for the arguments that occur here (|x| < 48): exact remainder, then + 24 when negative */
static double mod24(double x)
{
    double r = x;
    while (r >= 24.0) r -= 24.0;
    while (r <= -24.0) r += 24.0;
    if (r < 0.0) r += 24.0;
    return r;
}
static double cospi(double f) { return mcm::cos_r(3.141592653589793 * f); }

static double shape_value(int kind, double h, double t1, double v1, double t2, double v2)   /* diel.jl:398-400 */
{
    const double f = frac(h, t1, t2);
    if (kind == MC_SHAPE_LINEAR) return v1 + (v2 - v1) * f;
    if (kind == MC_SHAPE_DECAY) return v2 + (v1 - v2) * mcm::exp(-3 * f);
    return v1 + (v2 - v1) * (1 - cospi(f)) / 2;
}

static const mc_diel_shape* covering_shape(const mc_diel_curve& c, double h, const Day& day)   /* diel.jl:356-368 */
{
    const mc_diel_shape* filler = nullptr;
    for (int i = 0; i < c.n_shapes; i++) {
        const mc_diel_shape& s = c.shapes[i];
        if (s.kind == MC_SHAPE_SINE) filler = &s;
        else if (in_span(h, resolve_time(s.a, day), resolve_time(s.b, day))) return &s;
    }
    return filler;
}

static double sample_shape(const mc_diel_shape* s, double h, const Day& day, const mc_diel_curve& c, const double* vals, int depth);

static double endpoint_height(const mc_time_of_day& t, double th, const Day& day, const mc_diel_curve& c, const double* vals, int depth)   /* diel.jl:387-392 */
{
    for (int i = 0; i < c.n_inputs; i++) if (same_time(c.inputs[i], t)) return vals[i];
    if (depth > 2 * MC_DIEL_MAX_SHAPES) return NaN;                /* a seam cycle recurses for ever in the reference */
    const mc_diel_shape* prev = covering_shape(c, mod24(th - 1e-6), day);
    return sample_shape(prev, th, day, c, vals, depth + 1);
}
static double sample_shape(const mc_diel_shape* s, double h, const Day& day, const mc_diel_curve& c, const double* vals, int depth)   /* diel.jl:377-385 */
{
    if (!s) return NaN;
    const double ta = resolve_time(s->a, day), tb = resolve_time(s->b, day);
    const double va = endpoint_height(s->a, ta, day, c, vals, depth);
    const double vb = endpoint_height(s->b, tb, day, c, vals, depth);
    return shape_value(s->kind, h, ta, va, tb, vb);
}
static double sample(const mc_diel_curve& c, double h, const Day& day, const double* vals)    /* diel.jl:352-353 */
{
    return sample_shape(covering_shape(c, h, day), h, day, c, vals, 0);
}

/* vapour_pressure(GoffGratch(), T) [Pa] of the un-vendored FluidProperties.jl (the functor of Derived, diel.jl:285-287): same published formula as
 * mc_oracle.cpp's goff_gratch (pinned there against the reference's RHLOC goldens), ice branch at or below 273.16 K */
static double saturation_vapour_pressure(double T)
{
    const double Tp = T + 0.01;
    double loge;
    if (Tp <= 273.16) {
        const double r = 273.16 / Tp;
        loge = -9.09718 * (r - 1.0) - 3.56654 * mcm::log10(r) + 0.876793 * (1.0 - Tp / 273.16) + 0.7858350313586662;
    } else {
        const double r = 373.16 / Tp;
        loge = -7.90298 * (r - 1.0) + 5.02808 * mcm::log10(r) - 1.3816e-7 * (mcm::exp10(11.344 * (1.0 - Tp / 373.16)) - 1.0)
             + 8.1328e-3 * (mcm::exp10(-3.49149 * (r - 1.0)) - 1.0) + 3.0057148979490314;
    }
    return 100.0 * mcm::exp10(loge);
}

/* ---- shortwave (angstrom_maxwell.jl:21-46) ------------------------------------------------------------------------------------- */
static double erbs(double ci)                                                            /* erbs.jl:14-23 */
{
    if (ci <= 0.22) return 1 - 0.09 * ci;
    if (ci <= 0.80) return 0.9511 - 0.1604 * ci + 4.388 * (ci * ci) - 16.638 * (ci * ci * ci) + 12.336 * ((ci * ci) * (ci * ci));
    return 0.165;
}
static double jl_pow(double x, double y) { return y == 1.0 ? x : mcm::pow(x, y); }      /* pow(x, 1.0) is x in any correctly rounded pow */
static double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

static void shortwave_hour(const mc_weather_desc& w, int doy, double cloud, double diffuse_clear, double direct_clear, double* global_out, double* fraction_out)
{
    const double theta1 = 360.0 * (doy - 1) / 365.0;
    const double theta2 = 2.0 * theta1;
    double s1, c1, s2, c2;
    mcm::sincosd(theta1, &s1, &c1);
    mcm::sincosd(theta2, &s2, &c2);
    const double ec = 1.00011 + 0.034221 * c1 + 0.00128 * s1 + 0.000719 * c2 + 0.000077 * s2;
    const double eth = w.solar_constant * ec;
    const double sf = w.angstrom_a + w.angstrom_b * jl_pow(1.0 - cloud, w.angstrom_gamma);
    const double gcs = diffuse_clear + direct_clear;
    const double p = sf * gcs;
    const double gr = (p != p) ? p : (p > 0.0 ? p : 0.0);                               /* max(sf * gcs, 0) */
    *global_out = gr;
    const double den = eth > 1e-9 ? eth : 1e-9;
    const double ci = clampd(gr / den, 0.0, 1.2);
    *fraction_out = clampd(erbs(ci), 0.0, 1.0);
}

/* ---- solar geometry (see the header comment) ----------------------------------------------------------------------------------- */
struct SunDay { double sinde, cosde, sinphi, cosphi, ha_sunrise_h; };

static SunDay sun_day(const mc_weather_desc& w, double lat_deg, int doy)
{
    const double pi = 3.141592653589793;
    SunDay g;
    const double omega = 2.0 * pi / w.days_per_year;
    const double d = (double)doy;
    const double zeta = omega * (d - w.equinox_day) + 2.0 * w.orbital_eccentricity * (mcm::sin_r(omega * d) - mcm::sin_r(omega * w.equinox_day));
    g.sinde = w.sin_obliquity * mcm::sin_r(zeta);
    g.cosde = std::sqrt(1.0 - g.sinde * g.sinde);
    mcm::sincosd(lat_deg, &g.sinphi, &g.cosphi);
    /* sunrise hour angle H: cos H = -tan(delta) tan(phi); polar night H = 0, polar day H = pi */
    const double tt = -(g.sinde * g.sinphi) / (g.cosde * g.cosphi);
    double H;
    if (!(g.cosde * g.cosphi > 0.0)) H = (g.sinde * g.sinphi > 0.0) ? pi : 0.0;
    else if (tt >= 1.0) H = 0.0;
    else if (tt <= -1.0) H = pi;
    else H = mcm::acos(tt);
    g.ha_sunrise_h = H * 12.0 / pi;
    return g;
}

static void sun_hour(const mc_weather_desc& w, const SunDay& g, double hour, double noon_h, double slope_deg, double aspect_deg, double* zen_out, double* szen_out)
{
    const double rad2deg = 57.29577951308232, deg2rad = 1.74532925199432957692e-02;
    const double ha_deg = 15.0 * (hour - noon_h);
    double sh, ch;
    mcm::sincosd(ha_deg, &sh, &ch);
    double cz = g.cosphi * g.cosde * ch + g.sinphi * g.sinde;
    cz = clampd(cz, -1.0, 1.0);
    const double zgeo = mcm::acos(cz) * rad2deg;
    double z = zgeo;
    if (w.refraction && z > 88.0 && z <= 90.0) z = z - (31.0 - 7.5 * (90.0 - z)) / 60.0;
    *zen_out = z < 90.0 ? z : 90.0;                                                      /* src/simulation.jl:301 */
    if (slope_deg == 0.0) { *szen_out = *zen_out; return; }
    /* sun azimuth clockwise from north: 180 -/+ the azimuth from south, morning / afternoon */
    const double sz = std::sqrt(1.0 - cz * cz);
    const double den = g.cosphi * sz;
    double caz = den > 1e-12 ? (g.sinphi * cz - g.sinde) / den : 1.0;
    caz = clampd(caz, -1.0, 1.0);
    const double az_south = mcm::acos(caz) * rad2deg;
    const double az = sh < 0.0 ? 180.0 - az_south : 180.0 + az_south;
    double ss, cs, sd, cd;
    mcm::sincosd(slope_deg, &ss, &cs);
    double rel = az - aspect_deg;
    while (rel > 360.0) rel -= 360.0;
    while (rel < -360.0) rel += 360.0;
    mcm::sincosd(rel, &sd, &cd);
    const double czs = clampd(cz * cs + sz * ss * cd, -1.0, 1.0);
    const double zs = mcm::acos(czs) * rad2deg;
    *szen_out = zs < 90.0 ? zs : 90.0;                                                   /* src/simulation.jl:302 */
    (void)deg2rad;
}

}  // namespace mcw_oracle

extern "C" {

/* host arrays in the C-ABI layout ([hour or day][site]); returns 0, or -1 when something required is missing */
int mco_populate_weather(int64_t S, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                         const mc_weather_in_soa* in, mc_weather_out_soa* out)
{
    using namespace mcw_oracle;
    if (!w || !in || !out || S <= 0 || ndays <= 0 || !in->day_of_year) return -1;
    const bool need_geometry = !in->hour_solar_noon || !in->hour_angle_sunrise || out->zenith_angle || out->zenith_slope_angle;
    if (need_geometry && (!sites || !sites->latitude)) return -1;
    const bool want_sw = out->global_radiation || out->diffuse_fraction;
    if (want_sw && (!in->direct_horizontal || !in->diffuse_horizontal)) return -1;
    const mc_diel_curve* curves[4] = {&w->reference_temperature, &w->reference_wind_speed, &w->reference_humidity, &w->cloud_cover};
    const double* const* series[4] = {in->reference_temperature_values, in->reference_wind_speed_values, in->reference_humidity_values, in->cloud_cover_values};
    double* dst[4] = {out->reference_temperature, out->reference_wind_speed, out->reference_humidity, out->cloud_cover};
    for (int q = 0; q < 4; q++)
        if (dst[q] || (q == 3 && want_sw) || (q == 0 && dst[2] && w->humidity_source == MC_HUMIDITY_FROM_VAPOUR_PRESSURE))
            for (int i = 0; i < curves[q]->n_inputs; i++) if (!series[q][i]) return -1;
    if (out->pressure && !in->pressure_hourly && !in->pressure_site) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; s++) {
        for (int d = 0; d < ndays; d++) {
            const int doy = in->day_of_year[d];
            const size_t ds = (size_t)d * (size_t)S + (size_t)s;
            SunDay g{};
            if (need_geometry) g = sun_day(*w, sites->latitude[s], doy);
            const double noon = in->hour_solar_noon ? in->hour_solar_noon[ds] : 12.0 + (sites && sites->solar_noon_offset_h ? sites->solar_noon_offset_h[s] : 0.0);
            const double ha_rise = in->hour_angle_sunrise ? in->hour_angle_sunrise[ds] : g.ha_sunrise_h;
            if (out->hour_solar_noon) out->hour_solar_noon[ds] = noon;
            if (out->hour_angle_sunrise) out->hour_angle_sunrise[ds] = ha_rise;
            Day day;                                                                     /* diel.jl:319-324 */
            day.solar_noon = noon;
            day.sunrise = noon - ha_rise;
            day.sunset = 2 * noon - day.sunrise;
            double vals[4][MC_DIEL_MAX_INPUTS];
            for (int q = 0; q < 4; q++) for (int i = 0; i < curves[q]->n_inputs; i++) vals[q][i] = series[q][i] ? series[q][i][ds] : NaN;
            const bool derived_rh = w->humidity_source == MC_HUMIDITY_FROM_VAPOUR_PRESSURE;
            for (int h = 0; h < 24; h++) {
                const size_t hs = ((size_t)d * 24 + (size_t)h) * (size_t)S + (size_t)s;
                double cloud = NaN, T_hour = NaN;
                for (int q = 0; q < 4; q++) {
                    const bool feeds_rh = q == 0 && derived_rh && dst[2];          /* the derived humidity reads the hour's temperature (populate_quantities!: forcings first) */
                    if (!dst[q] && !(q == 3 && want_sw) && !feeds_rh) continue;
                    double v = sample(*curves[q], (double)h, day, vals[q]);
                    if (q == 0) T_hour = v;
                    if (q == 2 && derived_rh) v = clampd(v / saturation_vapour_pressure(T_hour), 0.0, 1.0);   /* fill_derived!, diel.jl:285-287,438-444 */
                    if ((q == 2 || q == 3) && v > 1.0) v = 1.0;                          /* src/simulation.jl:317-326 */
                    if (dst[q]) dst[q][hs] = v;
                    if (q == 3) cloud = v;
                }
                if (out->pressure) out->pressure[hs] = in->pressure_hourly ? in->pressure_hourly[hs] : in->pressure_site[s];   /* :328 */
                if (want_sw) {
                    double gr, fr;
                    shortwave_hour(*w, doy, cloud, in->diffuse_horizontal[hs], in->direct_horizontal[hs], &gr, &fr);
                    if (out->diffuse_fraction) out->diffuse_fraction[hs] = fr;           /* src/simulation.jl:29 */
                    if (out->global_radiation) out->global_radiation[hs] = in->global_radiation_hourly ? in->global_radiation_hourly[hs] : gr;   /* :30-34 */
                }
                if (out->zenith_angle || out->zenith_slope_angle) {
                    double z, zs;
                    sun_hour(*w, g, (double)h, noon, sites->slope ? sites->slope[s] : 0.0, sites->aspect ? sites->aspect[s] : 0.0, &z, &zs);
                    if (out->zenith_angle) out->zenith_angle[hs] = z;
                    if (out->zenith_slope_angle) out->zenith_slope_angle[hs] = zs;
                }
            }
        }
    }
    return 0;
}

}  // extern "C"
