/*
 * mc_weather.cuh — device code of the widening rows f2 / f3: populate_weather! and the solar geometry it needs,
 * one thread per (site, day), all hourly arrays in the soil kernel's own [hour][site] layout.
 *
 * Reference: src/simulation.jl:20-36 (populate_weather!), :306-331 (interpolate_minmax!), :340-352 (adjust_for_cloud_cover!),
 * src/inputs/diel.jl:316-408 (the declarative diel engine), src/radiation/shortwave/angstrom_maxwell.jl:21-46,
 * src/radiation/diffuse/erbs.jl:14-23, src/radiation/sunshine_fraction/angstrom.jl:14; solar geometry: see include/microclimate_b200.h.
 *
 * Not a translation of the Julia control flow.  diel.jl resolves the two endpoints of the covering shape — anchor times from the day's
 * solar geometry, heights from the input series or, for a seam, by recursing into the shape that covers the instant just before it — again
 * for every one of the 24 hours.  None of that depends on the hour, so a thread resolves every shape of a curve ONCE per day (seams by
 * fixed-point passes instead of recursion: the device has no use for a recursive call graph) and the hour loop is one span test and one
 * shape evaluation.  The floating-point operations that produce a value are the reference's, in the reference's order, so the result is
 * bit-identical to the literal restatement in oracle/mc_weather_oracle.cpp (tests/test_weather.py).
 *
 * Written against the same tiny platform shim as the soil code so tests/hostemu can run the kernel body on the CPU.
 */
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/microclimate_b200.h"
#include "../../include/mc_math.h"

#if defined(__CUDACC__)
#define MCW_DEV __device__ __forceinline__
#define MCW_LDG(p) __ldg(p)
#else
#define MCW_DEV static inline
#define MCW_LDG(p) (*(p))
#endif

namespace mcw {

/* kernel arguments: every pointer is a device pointer (host pointer in the host emulation); NULL = absent / not wanted */
struct WArgs {
    int64_t S;
    int32_t ndays;
    const int32_t* doy;                                          /* [ndays] */
    const double *lat, *slope, *aspect, *noon_off;               /* [S] */
    const double* vals[4][MC_DIEL_MAX_INPUTS];                   /* [ndays][S] per curve input: T, wind, RH, cloud */
    const double *noon_in, *ha_in;                               /* [ndays][S] */
    const double *direct, *diffuse, *gr_hourly, *p_hourly;       /* [24*ndays][S] */
    const double* p_site;                                        /* [S] */
    double *o_P, *o_T, *o_wind, *o_rh, *o_cloud, *o_gr, *o_df, *o_zen, *o_szen;   /* [24*ndays][S] */
    double *o_noon, *o_ha;                                       /* [ndays][S] */
};

MCW_DEV double qnan() { return mcm::u2d(0x7ff8000000000000ULL); }
MCW_DEV double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

/* ---- diel engine ---------------------------------------------------------------------------------------------------------------- */
struct Day { double sunrise, sunset, solar_noon; };

MCW_DEV double resolve_time(const mc_time_of_day& t, const Day& day)                     /* diel.jl:79-83 */
{
    const double base = t.kind == MC_TOD_SUNRISE ? day.sunrise : (t.kind == MC_TOD_SUNSET ? day.sunset : day.solar_noon);
    if (t.kind == MC_TOD_CLOCK || t.kind == MC_TOD_MIDNIGHT) return t.value;
    return base + t.value;
}
MCW_DEV bool in_span(double h, double t1, double t2)                                     /* diel.jl:371-375 */
{
    const double t2c = t2 <= t1 ? t2 + 24 : t2;
    const double hc = h < t1 ? h + 24 : h;
    return t1 <= hc && hc < t2c;
}
MCW_DEV double frac(double h, double t1, double t2)                                      /* diel.jl:403-407 */
{
    const double t2c = t2 <= t1 ? t2 + 24 : t2;
    const double hc = h < t1 ? h + 24 : h;
    return (hc - t1) / (t2c - t1);
}
MCW_DEV double mod24(double x)                                                           /* Julia mod(x, 24.0), |x| < 48 */
{
    double r = x;
    while (r >= 24.0) r -= 24.0;
    while (r <= -24.0) r += 24.0;
    if (r < 0.0) r += 24.0;
    return r;
}
MCW_DEV double shape_value(int kind, double h, double t1, double v1, double t2, double v2)   /* diel.jl:398-400 */
{
    const double f = frac(h, t1, t2);
    if (kind == MC_SHAPE_LINEAR) return v1 + (v2 - v1) * f;
    if (kind == MC_SHAPE_DECAY) return v2 + (v1 - v2) * mcm::exp(-3 * f);
    return v1 + (v2 - v1) * (1 - mcm::cos_r(3.141592653589793 * f)) / 2;
}

/* a curve resolved for one (site, day): anchor times and endpoint heights of every shape */
struct Resolved {
    int n, filler;
    int kind[MC_DIEL_MAX_SHAPES];
    double ta[MC_DIEL_MAX_SHAPES], tb[MC_DIEL_MAX_SHAPES], va[MC_DIEL_MAX_SHAPES], vb[MC_DIEL_MAX_SHAPES];
};

MCW_DEV int covering(const Resolved& r, double h)                                        /* diel.jl:356-368 on resolved times */
{
    for (int i = 0; i < r.n; i++)
        if (r.kind[i] != MC_SHAPE_SINE && in_span(h, r.ta[i], r.tb[i])) return i;
    return r.filler;
}
MCW_DEV double sample(const Resolved& r, double h)                                       /* diel.jl:352-353 */
{
    const int i = covering(r, h);
    if (i < 0) return qnan();
    return shape_value(r.kind[i], h, r.ta[i], r.va[i], r.tb[i], r.vb[i]);
}

MCW_DEV void resolve_curve(const mc_diel_curve& c, const Day& day, const double* vals, Resolved& r)
{
    r.n = c.n_shapes; r.filler = -1;
    unsigned known = 0;                                            /* bit 2i: height of shape i's first endpoint known, bit 2i+1: of its second */
    for (int i = 0; i < c.n_shapes; i++) {
        r.kind[i] = c.shapes[i].kind;
        if (r.kind[i] == MC_SHAPE_SINE) r.filler = i;              /* the last space-filling shape wins, as in covering_shape's loop */
        r.ta[i] = resolve_time(c.shapes[i].a, day);
        r.tb[i] = resolve_time(c.shapes[i].b, day);
        r.va[i] = r.vb[i] = qnan();
        for (int k = 0; k < c.n_inputs; k++) {                     /* input_index, diel.jl:394-399: the first matching input */
            if (!(known & (1u << (2 * i))) && c.inputs[k].kind == c.shapes[i].a.kind && c.inputs[k].value == c.shapes[i].a.value) { r.va[i] = vals[k]; known |= 1u << (2 * i); }
            if (!(known & (2u << (2 * i))) && c.inputs[k].kind == c.shapes[i].b.kind && c.inputs[k].value == c.shapes[i].b.value) { r.vb[i] = vals[k]; known |= 2u << (2 * i); }
        }
    }
    /* seams (endpoint_height, diel.jl:387-392): handed off from the shape covering the instant just before the endpoint */
    const unsigned all = (1u << (2 * c.n_shapes)) - 1u;
    for (int pass = 0; pass < 2 * MC_DIEL_MAX_SHAPES && known != all; pass++) {
        unsigned progress = 0;
        for (int e = 0; e < 2 * c.n_shapes; e++) {
            if (known & (1u << e)) continue;
            const int i = e >> 1;
            const double th = (e & 1) ? r.tb[i] : r.ta[i];
            const int p = covering(r, mod24(th - 1e-6));
            double v;
            if (p < 0) v = qnan();
            else if ((known >> (2 * p) & 3u) == 3u) v = shape_value(r.kind[p], th, r.ta[p], r.va[p], r.tb[p], r.vb[p]);
            else continue;
            if (e & 1) r.vb[i] = v; else r.va[i] = v;
            progress |= 1u << e;
        }
        if (!progress) break;                                      /* a seam cycle: the reference recurses for ever; heights stay NaN */
        known |= progress;
    }
}

/* vapour_pressure(GoffGratch(), T) [Pa] (FluidProperties.jl; the functor of Derived(RelativeHumidityFromVapourPressureAndTemperature), diel.jl:285-287) */
MCW_DEV double saturation_vapour_pressure(double T)
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

/* ---- shortwave (angstrom_maxwell.jl:21-46) -------------------------------------------------------------------------------------- */
MCW_DEV double erbs(double ci)                                                           /* erbs.jl:14-23 */
{
    if (ci <= 0.22) return 1 - 0.09 * ci;
    if (ci <= 0.80) return 0.9511 - 0.1604 * ci + 4.388 * (ci * ci) - 16.638 * (ci * ci * ci) + 12.336 * ((ci * ci) * (ci * ci));
    return 0.165;
}
/* the day's extraterrestrial irradiance factor depends on the day of year only: once per (site, day) */
MCW_DEV double extraterrestrial(const mc_weather_desc& w, int doy)
{
    const double theta1 = 360.0 * (doy - 1) / 365.0;
    const double theta2 = 2.0 * theta1;
    double s1, c1, s2, c2;
    mcm::sincosd(theta1, &s1, &c1);
    mcm::sincosd(theta2, &s2, &c2);
    const double ec = 1.00011 + 0.034221 * c1 + 0.00128 * s1 + 0.000719 * c2 + 0.000077 * s2;
    return w.solar_constant * ec;
}
MCW_DEV void shortwave_hour(const mc_weather_desc& w, double eth, double cloud, double diffuse_clear, double direct_clear, double* global_out, double* fraction_out)
{
    const double clear = 1.0 - cloud;
    const double sf = w.angstrom_a + w.angstrom_b * (w.angstrom_gamma == 1.0 ? clear : mcm::pow(clear, w.angstrom_gamma));
    const double gcs = diffuse_clear + direct_clear;
    const double p = sf * gcs;
    const double gr = (p != p) ? p : (p > 0.0 ? p : 0.0);
    *global_out = gr;
    const double den = eth > 1e-9 ? eth : 1e-9;
    const double ci = clampd(gr / den, 0.0, 1.2);
    *fraction_out = clampd(erbs(ci), 0.0, 1.0);
}

/* ---- solar geometry (McCullough & Porter 1971 / NicheMapR SOLRAD form; include/microclimate_b200.h) -------------------------------- */
struct SunDay { double sinde, cosde, sinphi, cosphi, ha_sunrise_h; };

MCW_DEV SunDay sun_day(const mc_weather_desc& w, double lat_deg, int doy)
{
    const double pi = 3.141592653589793;
    SunDay g;
    const double omega = 2.0 * pi / w.days_per_year;
    const double d = (double)doy;
    const double zeta = omega * (d - w.equinox_day) + 2.0 * w.orbital_eccentricity * (mcm::sin_r(omega * d) - mcm::sin_r(omega * w.equinox_day));
    g.sinde = w.sin_obliquity * mcm::sin_r(zeta);
    g.cosde = sqrt(1.0 - g.sinde * g.sinde);
    mcm::sincosd(lat_deg, &g.sinphi, &g.cosphi);
    const double tt = -(g.sinde * g.sinphi) / (g.cosde * g.cosphi);
    double H;
    if (!(g.cosde * g.cosphi > 0.0)) H = (g.sinde * g.sinphi > 0.0) ? pi : 0.0;
    else if (tt >= 1.0) H = 0.0;
    else if (tt <= -1.0) H = pi;
    else H = mcm::acos(tt);
    g.ha_sunrise_h = H * 12.0 / pi;
    return g;
}

MCW_DEV void sun_hour(const mc_weather_desc& w, const SunDay& g, double hour, double noon_h, double slope_deg, double aspect_deg, double* zen_out, double* szen_out)
{
    const double rad2deg = 57.29577951308232;
    const double ha_deg = 15.0 * (hour - noon_h);
    double sh, ch;
    mcm::sincosd(ha_deg, &sh, &ch);
    double cz = g.cosphi * g.cosde * ch + g.sinphi * g.sinde;
    cz = clampd(cz, -1.0, 1.0);
    double z = mcm::acos(cz) * rad2deg;
    if (w.refraction && z > 88.0 && z <= 90.0) z = z - (31.0 - 7.5 * (90.0 - z)) / 60.0;
    *zen_out = z < 90.0 ? z : 90.0;
    if (slope_deg == 0.0) { *szen_out = *zen_out; return; }
    const double sz = sqrt(1.0 - cz * cz);
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
    *szen_out = zs < 90.0 ? zs : 90.0;
}

/* ---- one (site, day) ------------------------------------------------------------------------------------------------------------- */
MCW_DEV void site_day(const mc_weather_desc& w, const WArgs& a, int64_t s, int d)
{
    const size_t S = (size_t)a.S;
    const size_t ds = (size_t)d * S + (size_t)s;
    const size_t h0 = (size_t)d * 24 * S + (size_t)s;             /* hour h of this day sits at h0 + h * S */
    const int doy = MCW_LDG(a.doy + d);
    const bool need_geometry = !a.noon_in || !a.ha_in || a.o_zen || a.o_szen;
    SunDay g; g.sinde = g.cosde = g.sinphi = g.cosphi = g.ha_sunrise_h = 0.0;
    if (need_geometry) g = sun_day(w, MCW_LDG(a.lat + s), doy);
    const double noon = a.noon_in ? MCW_LDG(a.noon_in + ds) : 12.0 + (a.noon_off ? MCW_LDG(a.noon_off + s) : 0.0);
    const double ha_rise = a.ha_in ? MCW_LDG(a.ha_in + ds) : g.ha_sunrise_h;
    if (a.o_noon) a.o_noon[ds] = noon;
    if (a.o_ha) a.o_ha[ds] = ha_rise;
    Day day;
    day.solar_noon = noon;
    day.sunrise = noon - ha_rise;
    day.sunset = 2 * noon - day.sunrise;

    const bool want_sw = a.o_gr || a.o_df;
    const bool derived_rh = w.humidity_source == MC_HUMIDITY_FROM_VAPOUR_PRESSURE && a.o_rh;
    const mc_diel_curve* curves[4] = {&w.reference_temperature, &w.reference_wind_speed, &w.reference_humidity, &w.cloud_cover};
    double* dst[4] = {a.o_T, a.o_wind, a.o_rh, a.o_cloud};
    double T_hour[24];                                             /* only touched when the humidity is derived from vapour pressure and temperature */
    for (int q = 0; q < 4; q++) {
        if (!dst[q] && !(q == 3 && want_sw) && !(q == 0 && derived_rh)) continue;
        const mc_diel_curve& c = *curves[q];
        double vals[MC_DIEL_MAX_INPUTS];
        for (int i = 0; i < MC_DIEL_MAX_INPUTS; i++) vals[i] = (i < c.n_inputs && a.vals[q][i]) ? MCW_LDG(a.vals[q][i] + ds) : qnan();
        Resolved r;
        resolve_curve(c, day, vals, r);
        const double eth = (q == 3 && want_sw) ? extraterrestrial(w, doy) : 0.0;
        for (int h = 0; h < 24; h++) {
            const size_t hs = h0 + (size_t)h * S;
            double v = sample(r, (double)h);
            if (q == 0 && derived_rh) T_hour[h] = v;
            if (q == 2 && derived_rh) v = clampd(v / saturation_vapour_pressure(T_hour[h]), 0.0, 1.0);      /* fill_derived!, diel.jl:285-287,438-444 */
            if (q >= 2 && v > 1.0) v = 1.0;                        /* humidity and cloud cover, src/simulation.jl:317-326 */
            if (dst[q]) dst[q][hs] = v;
            if (q == 3 && want_sw) {
                double gr, fr;
                shortwave_hour(w, eth, v, MCW_LDG(a.diffuse + hs), MCW_LDG(a.direct + hs), &gr, &fr);
                if (a.o_df) a.o_df[hs] = fr;                                             /* src/simulation.jl:29 */
                if (a.o_gr) a.o_gr[hs] = a.gr_hourly ? MCW_LDG(a.gr_hourly + hs) : gr;   /* src/simulation.jl:30-34 */
            }
        }
    }
    if (a.o_P) {
        const double ps = a.p_hourly ? 0.0 : MCW_LDG(a.p_site + s);
        for (int h = 0; h < 24; h++) { const size_t hs = h0 + (size_t)h * S; a.o_P[hs] = a.p_hourly ? MCW_LDG(a.p_hourly + hs) : ps; }   /* :328 */
    }
    if (a.o_zen || a.o_szen) {
        const double slope = a.slope ? MCW_LDG(a.slope + s) : 0.0, aspect = a.aspect ? MCW_LDG(a.aspect + s) : 0.0;
        for (int h = 0; h < 24; h++) {
            const size_t hs = h0 + (size_t)h * S;
            double z, zs;
            sun_hour(w, g, (double)h, noon, slope, aspect, &z, &zs);
            if (a.o_zen) a.o_zen[hs] = z;
            if (a.o_szen) a.o_szen[hs] = zs;
        }
    }
}

}  // namespace mcw
