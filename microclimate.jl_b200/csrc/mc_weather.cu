/*
 * mc_weather.cu — second translation unit of libmicroclimate_b200.so: the widening rows f2 / f3 of SURVEY 8 on the device
 * (populate_weather! + solar geometry, reference src/simulation.jl:20-36,291-352; device code in mc_weather.cuh) and the C-ABI entry
 * points around them (include/microclimate_b200.h, "Widening").  A separate unit so that the machine code of the soil kernels
 * (mc_api.cu) — which the committed ncu captures are tied to by SASS hash — does not move when this part changes.
 *
 * weather_kernel: one thread per (site, day); grid = (ceil(S / 128), ndays).  A thread reads its ~13 daily inputs and 48 clear-sky values
 * and writes up to 9 x 24 hourly values; for fixed (hour, variable) the 32 lanes of a warp touch 32 consecutive doubles, so every load
 * and store is one coalesced 256-byte request.  HBM bound by construction: ~0.5 KB in and 1.7 KB out per site-day.
 */
#ifndef MCW_HOST_EMULATION             /* tests/hostemu/weather_api_hostemu.cpp compiles this file with g++ over a stand-in CUDA runtime (CPU tier) */
#include <cuda_runtime.h>
#endif
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

/* each translation unit carries its own copy of the math kernel (its out-of-line functions and its constant table are not `static`, and this
 * build has no relocatable device code): same source, different namespace name */
#define mcm mcm_weather
#include "mc_weather.cuh"
#include "mc_internal.h"

namespace {

constexpr int WX_BLOCK = 128;

__global__ void __launch_bounds__(WX_BLOCK) weather_kernel(const __grid_constant__ mc_weather_desc w, const __grid_constant__ mcw::WArgs a)
{
    const int64_t s = (int64_t)blockIdx.x * WX_BLOCK + threadIdx.x;
    if (s >= a.S) return;
    mcw::site_day(w, a, s, (int)blockIdx.y);
}

#define WX_CUDA(h, call)                                                                                       \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) {                                                                               \
            char b_[256];                                                                                      \
            std::snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return mc_i_fail(h, b_, -100 - (int)e_);                                                           \
        }                                                                                                      \
    } while (0)

/* per-handle device staging, grown on demand, released by mc_destroy through mc_i_weather_release */
struct WxStage {
    double* buf = nullptr; size_t doubles = 0;
    int32_t* doy = nullptr; size_t doy_cap = 0;
};
std::mutex g_stage_mutex;
std::map<mc_handle*, WxStage> g_stage;

WxStage* stage_of(mc_handle* h)
{
    std::lock_guard<std::mutex> g(g_stage_mutex);
    return &g_stage[h];
}

int ensure(mc_handle* h, WxStage* st, size_t doubles, size_t ndays)
{
    if (st->doubles < doubles) {
        WX_CUDA(h, cudaStreamSynchronize(mc_i_stream(h)));
        if (st->buf) { WX_CUDA(h, cudaFree(st->buf)); st->buf = nullptr; st->doubles = 0; }
        WX_CUDA(h, cudaMalloc(&st->buf, doubles * sizeof(double)));
        st->doubles = doubles;
    }
    if (st->doy_cap < ndays) {
        WX_CUDA(h, cudaStreamSynchronize(mc_i_stream(h)));
        if (st->doy) { WX_CUDA(h, cudaFree(st->doy)); st->doy = nullptr; st->doy_cap = 0; }
        WX_CUDA(h, cudaMalloc(&st->doy, ndays * sizeof(int32_t)));
        st->doy_cap = ndays;
    }
    return 0;
}

const mc_diel_curve* curve_of(const mc_weather_desc* w, int q)
{
    return q == 0 ? &w->reference_temperature : (q == 1 ? &w->reference_wind_speed : (q == 2 ? &w->reference_humidity : &w->cloud_cover));
}
const double* const* series_of(const mc_weather_in_soa* in, int q)
{
    return q == 0 ? in->reference_temperature_values : (q == 1 ? in->reference_wind_speed_values : (q == 2 ? in->reference_humidity_values : in->cloud_cover_values));
}
double* wx_out_of(const mc_weather_out_soa* o, int q) { return q == 0 ? o->reference_temperature : (q == 1 ? o->reference_wind_speed : (q == 2 ? o->reference_humidity : o->cloud_cover)); }

/* nominal cyclic position of an anchor for the structural coverage check (diel.jl:87-91) */
double nominal_hour(const mc_time_of_day& t)
{
    const double base = t.kind == MC_TOD_SUNRISE ? 6.0 : (t.kind == MC_TOD_MIDDAY ? 12.0 : (t.kind == MC_TOD_SUNSET ? 18.0 : 0.0));
    double x = std::fmod(base + t.value, 24.0);
    if (x < 0.0) x += 24.0;
    return x;
}

int validate_curve(const mc_diel_curve& c, const char* name, std::string& err)
{
    if (c.n_shapes < 1 || c.n_shapes > MC_DIEL_MAX_SHAPES) { err = std::string(name) + ": n_shapes must be in 1.." + std::to_string(MC_DIEL_MAX_SHAPES); return -1; }
    if (c.n_inputs < 0 || c.n_inputs > MC_DIEL_MAX_INPUTS) { err = std::string(name) + ": n_inputs must be in 0.." + std::to_string(MC_DIEL_MAX_INPUTS); return -1; }
    int nfill = 0;
    std::vector<std::pair<double, double>> arcs;                   /* (start, length) of every bounded shape on the nominal event order */
    for (int i = 0; i < c.n_shapes; i++) {
        const mc_diel_shape& s = c.shapes[i];
        if (s.kind < MC_SHAPE_SINE || s.kind > MC_SHAPE_LINEAR) { err = std::string(name) + ": unknown shape kind"; return -1; }
        for (const mc_time_of_day* t : {&s.a, &s.b}) if (t->kind < MC_TOD_CLOCK || t->kind > MC_TOD_MIDNIGHT) { err = std::string(name) + ": unknown time-of-day kind"; return -1; }
        if (s.kind == MC_SHAPE_SINE) { nfill++; continue; }
        const double a = nominal_hour(s.a), b = nominal_hour(s.b);
        double len = std::fmod(b - a, 24.0);
        if (len < 0.0) len += 24.0;
        if (len == 0.0) { err = std::string(name) + ": a shape spans zero length"; return -1; }                      /* diel.jl:196 */
        arcs.push_back({a, len});
    }
    if (nfill > 1) { err = std::string(name) + ": " + std::to_string(nfill) + " space-filling shapes; coverage is ambiguous (allow at most one)"; return -1; }   /* :188 */
    if (!arcs.empty()) {
        std::sort(arcs.begin(), arcs.end());
        const size_t n = arcs.size();
        for (size_t i = 0; i < n; i++) {
            const double cover_end = arcs[i].first + arcs[i].second;
            const double next_start = i + 1 < n ? arcs[i + 1].first : arcs[0].first + 24.0;
            if (cover_end > next_start + 1e-9) { err = std::string(name) + ": shapes overlap"; return -1; }           /* :206 */
            if (nfill == 0 && cover_end < next_start - 1e-9) { err = std::string(name) + ": day not covered - add a shape or a space-filling one"; return -1; }   /* :208 */
        }
    }
    for (int k = 0; k < c.n_inputs; k++) {                          /* every input must name an endpoint of some shape, diel.jl:216-222 */
        bool found = false;
        for (int i = 0; i < c.n_shapes && !found; i++)
            for (const mc_time_of_day* t : {&c.shapes[i].a, &c.shapes[i].b}) if (t->kind == c.inputs[k].kind && t->value == c.inputs[k].value) found = true;
        if (!found) { err = std::string(name) + ": an input is not an endpoint of any shape"; return -1; }
    }
    return 0;
}

int validate_weather(const mc_weather_desc* w, std::string& err)
{
    if (!w) { err = "weather descriptor is NULL"; return -1; }
    static const char* names[4] = {"reference_temperature", "reference_wind_speed", "reference_humidity", "cloud_cover"};
    for (int q = 0; q < 4; q++) if (int rc = validate_curve(*curve_of(w, q), names[q], err)) return rc;
    if (w->diffuse_model != MC_DIFFUSE_ERBS) { err = "diffuse_model: only MC_DIFFUSE_ERBS exists"; return -1; }
    if (w->humidity_source != MC_HUMIDITY_CURVE && w->humidity_source != MC_HUMIDITY_FROM_VAPOUR_PRESSURE) { err = "humidity_source: unknown value"; return -1; }
    if (!(w->days_per_year > 0.0)) { err = "days_per_year must be positive"; return -1; }
    return 0;
}

/* checks shared by the host-buffer and device-pointer entry points; fills the kernel arguments from (device) pointers */
int make_args(mc_handle* h, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites, const mc_weather_in_soa* in,
              const mc_weather_out_soa* out, const int32_t* doy_dev, mcw::WArgs& a)
{
    std::string err;
    if (validate_weather(w, err)) return mc_i_fail(h, err.c_str(), -1);
    if (!in || !out || ndays <= 0 || ndays > 65535) return mc_i_fail(h, "weather in/out must be non-NULL and ndays in 1..65535", -1);
    std::memset(&a, 0, sizeof(a));
    a.S = mc_i_sites(h); a.ndays = ndays; a.doy = doy_dev;
    const bool need_geometry = !in->hour_solar_noon || !in->hour_angle_sunrise || out->zenith_angle || out->zenith_slope_angle;
    if (need_geometry && (!sites || !sites->latitude)) return mc_i_fail(h, "the built-in solar geometry needs mc_solar_sites_soa.latitude", -1);
    if (sites) { a.lat = sites->latitude; a.slope = sites->slope; a.aspect = sites->aspect; a.noon_off = sites->solar_noon_offset_h; }
    const bool want_sw = out->global_radiation || out->diffuse_fraction;
    if (want_sw && (!in->direct_horizontal || !in->diffuse_horizontal)) return mc_i_fail(h, "global_radiation / diffuse_fraction need the clear-sky direct_horizontal and diffuse_horizontal", -1);
    for (int q = 0; q < 4; q++) {
        const bool wanted = wx_out_of(out, q) || (q == 3 && want_sw) || (q == 0 && out->reference_humidity && w->humidity_source == MC_HUMIDITY_FROM_VAPOUR_PRESSURE);
        for (int i = 0; i < curve_of(w, q)->n_inputs; i++) {
            if (wanted && !series_of(in, q)[i]) return mc_i_fail(h, "a value series of a wanted quantity is NULL (values[i] belongs to curve.inputs[i])", -1);
            a.vals[q][i] = series_of(in, q)[i];
        }
    }
    if (out->pressure && !in->pressure_hourly && !in->pressure_site) return mc_i_fail(h, "pressure wanted but neither pressure_hourly nor pressure_site given", -1);
    a.noon_in = in->hour_solar_noon; a.ha_in = in->hour_angle_sunrise; a.direct = in->direct_horizontal; a.diffuse = in->diffuse_horizontal;
    a.gr_hourly = in->global_radiation_hourly; a.p_hourly = in->pressure_hourly; a.p_site = in->pressure_site;
    a.o_P = out->pressure; a.o_T = out->reference_temperature; a.o_wind = out->reference_wind_speed; a.o_rh = out->reference_humidity;
    a.o_cloud = out->cloud_cover; a.o_gr = out->global_radiation; a.o_df = out->diffuse_fraction; a.o_zen = out->zenith_angle; a.o_szen = out->zenith_slope_angle;
    a.o_noon = out->hour_solar_noon; a.o_ha = out->hour_angle_sunrise;
    return 0;
}

#ifdef MCW_HOST_EMULATION
bool mcw_emulation_check(const mcw::WArgs& a);         /* the harness knows which addresses its stand-in cudaMalloc handed out */
#endif

int launch(mc_handle* h, const mc_weather_desc* w, const mcw::WArgs& a)
{
    const dim3 grid((unsigned)((a.S + WX_BLOCK - 1) / WX_BLOCK), (unsigned)a.ndays);
#ifndef MCW_HOST_EMULATION
    weather_kernel<<<grid, WX_BLOCK, 0, mc_i_stream(h)>>>(*w, a);
#else
    if (!mcw_emulation_check(a)) return mc_i_fail(h, "emulation: a pointer that is not device memory reached weather_kernel", -9);
    for (blockIdx.y = 0; blockIdx.y < grid.y; blockIdx.y++) for (blockIdx.x = 0; blockIdx.x < grid.x; blockIdx.x++)
        for (threadIdx.x = 0; threadIdx.x < (unsigned)WX_BLOCK; threadIdx.x++) weather_kernel(*w, a);
#endif
    WX_CUDA(h, cudaGetLastError());
    return 0;
}

/* a planned device copy of one [rows][S] plane of a host array */
struct Plane { const double* host_in; double* host_out; double* dev; size_t rows; };

}  // namespace

void mc_i_weather_release(mc_handle* h)
{
    std::lock_guard<std::mutex> g(g_stage_mutex);
    auto it = g_stage.find(h);
    if (it == g_stage.end()) return;
    cudaFree(it->second.buf); cudaFree(it->second.doy);
    g_stage.erase(it);
}

extern "C" {

void mc_default_weather(mc_weather_desc* w)
{
    if (!w) return;
    std::memset(w, 0, sizeof(*w));
    auto tod = [](int kind, double v) { mc_time_of_day t; t.kind = kind; t._pad = 0; t.value = v; return t; };
    auto shape = [&](int kind, mc_time_of_day a, mc_time_of_day b) { mc_diel_shape s; s.kind = kind; s._pad = 0; s.a = a; s.b = b; return s; };
    /* minmax_forcing_model() with its default offsets, src/inputs/diel.jl:456-495 */
    {   /* temperature: sine from sunrise to one hour past midday, overnight decay from sunset; inputs (min, max) */
        mc_diel_curve& c = w->reference_temperature;
        c.n_shapes = 2; c.n_inputs = 2;
        c.shapes[0] = shape(MC_SHAPE_SINE, tod(MC_TOD_SUNRISE, 0.0), tod(MC_TOD_MIDDAY, 1.0));
        c.shapes[1] = shape(MC_SHAPE_DECAY, tod(MC_TOD_SUNSET, 0.0), tod(MC_TOD_SUNRISE, 0.0));
        c.inputs[0] = tod(MC_TOD_SUNRISE, 0.0); c.inputs[1] = tod(MC_TOD_MIDDAY, 1.0);
    }
    auto three_lines = [&](mc_diel_curve& c, double sunrise_offset, double midday_offset) {   /* inputs (midnight, dawn side, midday side) */
        c.n_shapes = 3; c.n_inputs = 3;
        c.shapes[0] = shape(MC_SHAPE_LINEAR, tod(MC_TOD_MIDNIGHT, 0.0), tod(MC_TOD_SUNRISE, sunrise_offset));
        c.shapes[1] = shape(MC_SHAPE_LINEAR, tod(MC_TOD_SUNRISE, sunrise_offset), tod(MC_TOD_MIDDAY, midday_offset));
        c.shapes[2] = shape(MC_SHAPE_LINEAR, tod(MC_TOD_MIDDAY, midday_offset), tod(MC_TOD_MIDNIGHT, 0.0));
        c.inputs[0] = tod(MC_TOD_MIDNIGHT, 0.0); c.inputs[1] = tod(MC_TOD_SUNRISE, sunrise_offset); c.inputs[2] = tod(MC_TOD_MIDDAY, midday_offset);
    };
    three_lines(w->reference_wind_speed, 0.0, 1.0);               /* (mean, min, max) */
    three_lines(w->reference_humidity, 1.0, 0.0);                 /* (mean, max, min) */
    three_lines(w->cloud_cover, 1.0, 0.0);                        /* (mean, max, min) */
    w->angstrom_a = 0.36; w->angstrom_b = 0.64; w->angstrom_gamma = 1.0;      /* sunshine_fraction/angstrom.jl:8-12 */
    w->solar_constant = 1367.0;                                   /* angstrom_maxwell.jl:27 */
    w->diffuse_model = MC_DIFFUSE_ERBS;
    w->refraction = 1;
    w->humidity_source = MC_HUMIDITY_CURVE;
    w->orbital_eccentricity = 0.0167238; w->sin_obliquity = 0.39784993; w->days_per_year = 365.0; w->equinox_day = 80.0;
}

int mc_validate_weather(const mc_weather_desc* w, char* err, int err_len)
{
    std::string e;
    const int rc = validate_weather(w, e);
    if (err && err_len > 0) { std::strncpy(err, e.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
    return rc;
}

int64_t mc_abi_sizeof(int which)
{
    switch (which) {
    case 0: return (int64_t)sizeof(mc_model_desc);
    case 1: return (int64_t)sizeof(mc_sites_soa);
    case 2: return (int64_t)sizeof(mc_initial_soa);
    case 3: return (int64_t)sizeof(mc_state_soa);
    case 4: return (int64_t)sizeof(mc_forcing_soa);
    case 5: return (int64_t)sizeof(mc_output_soa);
    case 6: return (int64_t)sizeof(mc_counters);
    case 7: return (int64_t)sizeof(mc_weather_desc);
    case 8: return (int64_t)sizeof(mc_solar_sites_soa);
    case 9: return (int64_t)sizeof(mc_weather_in_soa);
    case 10: return (int64_t)sizeof(mc_weather_out_soa);
    default: return -1;
    }
}

int mc_populate_weather_device(mc_handle* h, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites_dev,
                               const mc_weather_in_soa* in_dev, mc_weather_out_soa* out_dev)
{
    if (!h) return -1;
    if (!in_dev || !in_dev->day_of_year) return mc_i_fail(h, "mc_weather_in_soa.day_of_year (host pointer) is required", -1);
    WX_CUDA(h, cudaSetDevice(mc_i_device(h)));
    WxStage* st = stage_of(h);
    if (int rc = ensure(h, st, 0, (size_t)(ndays > 0 ? ndays : 1))) return rc;
    mcw::WArgs a;
    if (int rc = make_args(h, ndays, w, sites_dev, in_dev, out_dev, st->doy, a)) return rc;
    WX_CUDA(h, cudaMemcpyAsync(st->doy, in_dev->day_of_year, (size_t)ndays * sizeof(int32_t), cudaMemcpyHostToDevice, mc_i_stream(h)));
    return launch(h, w, a);
}

/* host arrays in, host arrays out: stage, expand, copy back (synchronous) */
int mc_populate_weather(mc_handle* h, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                        const mc_weather_in_soa* in, mc_weather_out_soa* out)
{
    if (!h) return -1;
    if (!in || !out || ndays <= 0) return mc_i_fail(h, "weather in/out must be non-NULL and ndays positive", -1);
    if (!in->day_of_year) return mc_i_fail(h, "mc_weather_in_soa.day_of_year is required", -1);
    WX_CUDA(h, cudaSetDevice(mc_i_device(h)));
    const size_t S = (size_t)mc_i_sites(h), D = (size_t)ndays, H = 24 * D;
    std::vector<Plane> planes;
    mc_solar_sites_soa sd; std::memset(&sd, 0, sizeof(sd));
    mc_weather_in_soa id; std::memset(&id, 0, sizeof(id));
    mc_weather_out_soa od; std::memset(&od, 0, sizeof(od));
    size_t need = 0;
    auto plan_in = [&](const double* src, const double** dst, size_t rows) { if (src) { planes.push_back({src, nullptr, nullptr, rows}); *dst = (const double*)(uintptr_t)(need + 1); need += rows * S; } };
    auto plan_out = [&](double* src, double** dst, size_t rows) { if (src) { planes.push_back({nullptr, src, nullptr, rows}); *dst = (double*)(uintptr_t)(need + 1); need += rows * S; } };
    if (sites) { plan_in(sites->latitude, &sd.latitude, 1); plan_in(sites->slope, &sd.slope, 1); plan_in(sites->aspect, &sd.aspect, 1); plan_in(sites->solar_noon_offset_h, &sd.solar_noon_offset_h, 1); }
    for (int i = 0; i < MC_DIEL_MAX_INPUTS; i++) {
        plan_in(in->reference_temperature_values[i], &id.reference_temperature_values[i], D);
        plan_in(in->reference_wind_speed_values[i], &id.reference_wind_speed_values[i], D);
        plan_in(in->reference_humidity_values[i], &id.reference_humidity_values[i], D);
        plan_in(in->cloud_cover_values[i], &id.cloud_cover_values[i], D);
    }
    plan_in(in->hour_solar_noon, &id.hour_solar_noon, D); plan_in(in->hour_angle_sunrise, &id.hour_angle_sunrise, D);
    plan_in(in->direct_horizontal, &id.direct_horizontal, H); plan_in(in->diffuse_horizontal, &id.diffuse_horizontal, H);
    plan_in(in->global_radiation_hourly, &id.global_radiation_hourly, H); plan_in(in->pressure_hourly, &id.pressure_hourly, H);
    plan_in(in->pressure_site, &id.pressure_site, 1);
    plan_out(out->pressure, &od.pressure, H); plan_out(out->reference_temperature, &od.reference_temperature, H);
    plan_out(out->reference_wind_speed, &od.reference_wind_speed, H); plan_out(out->reference_humidity, &od.reference_humidity, H);
    plan_out(out->cloud_cover, &od.cloud_cover, H); plan_out(out->global_radiation, &od.global_radiation, H);
    plan_out(out->diffuse_fraction, &od.diffuse_fraction, H); plan_out(out->zenith_angle, &od.zenith_angle, H);
    plan_out(out->zenith_slope_angle, &od.zenith_slope_angle, H);
    plan_out(out->hour_solar_noon, &od.hour_solar_noon, D); plan_out(out->hour_angle_sunrise, &od.hour_angle_sunrise, D);
    WxStage* st = stage_of(h);
    if (int rc = ensure(h, st, need ? need : 1, D)) return rc;
    /* the planned "pointers" are offsets + 1 into the staging buffer: rebase them */
    auto rebase = [&](const double** p) { if (*p) *p = st->buf + ((size_t)(uintptr_t)*p - 1); };
    auto rebase_out = [&](double** p) { if (*p) *p = st->buf + ((size_t)(uintptr_t)*p - 1); };
    rebase(&sd.latitude); rebase(&sd.slope); rebase(&sd.aspect); rebase(&sd.solar_noon_offset_h);
    for (int i = 0; i < MC_DIEL_MAX_INPUTS; i++) { rebase(&id.reference_temperature_values[i]); rebase(&id.reference_wind_speed_values[i]); rebase(&id.reference_humidity_values[i]); rebase(&id.cloud_cover_values[i]); }
    rebase(&id.hour_solar_noon); rebase(&id.hour_angle_sunrise); rebase(&id.direct_horizontal); rebase(&id.diffuse_horizontal);
    rebase(&id.global_radiation_hourly); rebase(&id.pressure_hourly); rebase(&id.pressure_site);
    rebase_out(&od.pressure); rebase_out(&od.reference_temperature); rebase_out(&od.reference_wind_speed); rebase_out(&od.reference_humidity);
    rebase_out(&od.cloud_cover); rebase_out(&od.global_radiation); rebase_out(&od.diffuse_fraction); rebase_out(&od.zenith_angle);
    rebase_out(&od.zenith_slope_angle); rebase_out(&od.hour_solar_noon); rebase_out(&od.hour_angle_sunrise);
    id.day_of_year = in->day_of_year;
    mcw::WArgs a;
    if (int rc = make_args(h, ndays, w, sites ? &sd : nullptr, &id, &od, st->doy, a)) return rc;
    cudaStream_t s = mc_i_stream(h);
    WX_CUDA(h, cudaMemcpyAsync(st->doy, in->day_of_year, D * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    double* p = st->buf;
    for (Plane& pl : planes) { pl.dev = p; p += pl.rows * S; }
    for (const Plane& pl : planes) if (pl.host_in) WX_CUDA(h, cudaMemcpyAsync(pl.dev, pl.host_in, pl.rows * S * sizeof(double), cudaMemcpyHostToDevice, s));
    if (int rc = launch(h, w, a)) return rc;
    for (const Plane& pl : planes) if (pl.host_out) WX_CUDA(h, cudaMemcpyAsync(pl.host_out, pl.dev, pl.rows * S * sizeof(double), cudaMemcpyDeviceToHost, s));
    WX_CUDA(h, cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"

namespace {

/* one [rows][S] plane between a host array that is `pitch` doubles wide (this handle's sites start at column col0) and the packed device copy */
cudaError_t copy_plane(const Plane& pl, bool to_device, size_t S, size_t pitch, size_t col0, cudaStream_t s)
{
    const size_t w = S * sizeof(double);
    if (pitch == S)
        return to_device ? cudaMemcpyAsync(pl.dev, pl.host_in, pl.rows * w, cudaMemcpyHostToDevice, s) : cudaMemcpyAsync(pl.host_out, pl.dev, pl.rows * w, cudaMemcpyDeviceToHost, s);
    return to_device ? cudaMemcpy2DAsync(pl.dev, w, pl.host_in + col0, pitch * sizeof(double), w, pl.rows, cudaMemcpyHostToDevice, s)
                     : cudaMemcpy2DAsync(pl.host_out + col0, pitch * sizeof(double), pl.dev, w, w, pl.rows, cudaMemcpyDeviceToHost, s);
}

/* everything of mc_step_days_minmax up to and including the enqueued device-to-host copies; host arrays may be wider than the handle (a multi-device job) */
int minmax_enqueue(mc_handle* h, int64_t day0, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites, const mc_weather_in_soa* in,
                   const mc_forcing_soa* rest, mc_output_soa* out, mc_weather_out_soa* weather_host, size_t pitch, size_t col0)
{
    if (!in || !rest || !out || ndays <= 0) return mc_i_fail(h, "weather in / rest / out must be non-NULL and ndays positive", -1);
    if (!in->day_of_year && !rest->day_of_year) return mc_i_fail(h, "day_of_year is required (mc_weather_in_soa or mc_forcing_soa)", -1);
    const int32_t* doy_host = in->day_of_year ? in->day_of_year : rest->day_of_year;
    if (int rc = mc_wait(h); rc < 0) return rc;                    /* chunks of the asynchronous pipeline still in flight use the same compute stream */
    WX_CUDA(h, cudaSetDevice(mc_i_device(h)));
    const mc_model_desc* m = mc_i_model(h);
    const size_t S = (size_t)mc_i_sites(h), D = (size_t)ndays, H = 24 * D, N = (size_t)m->n_soil, NSn = (size_t)m->n_snow, NH = (size_t)m->n_heights;
    const bool own_zenith = !rest->zenith_angle || !rest->zenith_slope_angle;

    std::vector<Plane> planes;
    size_t need = 0;
    auto add = [&](const double* hin, double* hout, size_t rows) -> size_t { const size_t off = need; planes.push_back({hin, hout, nullptr, rows}); need += rows * S; return off + 1; };
    /* weather inputs */
    size_t o_site[4] = {0, 0, 0, 0}, o_vals[4][MC_DIEL_MAX_INPUTS] = {}, o_noon = 0, o_ha = 0, o_dir = 0, o_dif = 0, o_grh = 0, o_ph = 0, o_ps = 0;
    if (sites) {
        const double* sp[4] = {sites->latitude, sites->slope, sites->aspect, sites->solar_noon_offset_h};
        for (int k = 0; k < 4; k++) if (sp[k]) o_site[k] = add(sp[k], nullptr, 1);
    }
    for (int q = 0; q < 4; q++) for (int i = 0; i < MC_DIEL_MAX_INPUTS; i++) if (series_of(in, q)[i]) o_vals[q][i] = add(series_of(in, q)[i], nullptr, D);
    if (in->hour_solar_noon) o_noon = add(in->hour_solar_noon, nullptr, D);
    if (in->hour_angle_sunrise) o_ha = add(in->hour_angle_sunrise, nullptr, D);
    if (in->direct_horizontal) o_dir = add(in->direct_horizontal, nullptr, H);
    if (in->diffuse_horizontal) o_dif = add(in->diffuse_horizontal, nullptr, H);
    if (in->global_radiation_hourly) o_grh = add(in->global_radiation_hourly, nullptr, H);
    if (in->pressure_hourly) o_ph = add(in->pressure_hourly, nullptr, H);
    if (in->pressure_site) o_ps = add(in->pressure_site, nullptr, 1);
    /* expanded weather = the hourly members of the soil kernel's forcing (copied back only where weather_host asks) */
    double* wh[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (weather_host) { double* t[9] = {weather_host->pressure, weather_host->reference_temperature, weather_host->reference_wind_speed, weather_host->reference_humidity,
                                        weather_host->cloud_cover, weather_host->global_radiation, weather_host->diffuse_fraction, weather_host->zenith_angle, weather_host->zenith_slope_angle};
                        for (int k = 0; k < 9; k++) wh[k] = t[k]; }
    size_t o_wx[9];
    for (int k = 0; k < 9; k++) {
        if (k >= 7 && !own_zenith) o_wx[k] = add(k == 7 ? rest->zenith_angle : rest->zenith_slope_angle, nullptr, H);    /* SolarRadiation.jl's own angles */
        else o_wx[k] = add(nullptr, wh[k], H);
    }
    size_t o_wnoon = 0, o_wha = 0;
    if (weather_host && weather_host->hour_solar_noon) o_wnoon = add(nullptr, weather_host->hour_solar_noon, D);
    if (weather_host && weather_host->hour_angle_sunrise) o_wha = add(nullptr, weather_host->hour_angle_sunrise, D);
    /* what populate_weather! does not produce */
    size_t o_rainh = 0, o_lw = 0, o_daily[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o_presc = 0;
    if (rest->rainfall_hourly) o_rainh = add(rest->rainfall_hourly, nullptr, H);
    if (rest->longwave_radiation) o_lw = add(rest->longwave_radiation, nullptr, H);
    const double* daily_src[8] = {rest->leaf_area_index, rest->shade, rest->surface_emissivity, rest->cloud_emissivity, rest->soil_wetness, rest->deep_soil_temperature,
                                  rest->rainfall_daily, rest->albedo_daily};
    for (int k = 0; k < 8; k++) if (daily_src[k]) o_daily[k] = add(daily_src[k], nullptr, D);
    if (rest->prescribed_soil_moisture) o_presc = add(rest->prescribed_soil_moisture, nullptr, D * N);
    /* soil outputs */
    double* node_host[7] = {out->soil_temperature, out->soil_moisture, out->soil_water_potential, out->soil_humidity, out->soil_thermal_conductivity, out->soil_heat_capacity, out->soil_bulk_density};
    double* scal_host[5] = {out->surface_water, out->sky_temperature, NSn ? out->snow_fall : nullptr, NSn ? out->snow_depth : nullptr, NSn ? out->snow_density : nullptr};
    double* prof_host[3] = {out->air_temperature, out->wind_speed, out->relative_humidity};
    double* profs_host[2] = {out->convective_heat_flux, out->friction_velocity};
    size_t o_node[7] = {}, o_scal[5] = {}, o_snowT = 0, o_prof[3] = {}, o_profs[2] = {};
    const size_t out_begin = need;
    for (int k = 0; k < 7; k++) if (node_host[k]) o_node[k] = add(nullptr, node_host[k], H * N);
    for (int k = 0; k < 5; k++) if (scal_host[k]) o_scal[k] = add(nullptr, scal_host[k], H);
    if (NSn && out->snow_temperature) o_snowT = add(nullptr, out->snow_temperature, H * NSn);
    for (int k = 0; k < 3; k++) if (prof_host[k]) o_prof[k] = add(nullptr, prof_host[k], H * NH);
    for (int k = 0; k < 2; k++) if (profs_host[k]) o_profs[k] = add(nullptr, profs_host[k], H);
    const size_t out_end = need;

    WxStage* st = stage_of(h);
    if (int rc = ensure(h, st, need ? need : 1, D)) return rc;
    double* base = st->buf;
    auto at = [&](size_t off1) -> double* { return off1 ? base + (off1 - 1) : nullptr; };
    { double* p = base; for (Plane& pl : planes) { pl.dev = p; p += pl.rows * S; } }

    mc_solar_sites_soa sd; std::memset(&sd, 0, sizeof(sd));
    sd.latitude = at(o_site[0]); sd.slope = at(o_site[1]); sd.aspect = at(o_site[2]); sd.solar_noon_offset_h = at(o_site[3]);
    mc_weather_in_soa id; std::memset(&id, 0, sizeof(id));
    for (int i = 0; i < MC_DIEL_MAX_INPUTS; i++) {
        id.reference_temperature_values[i] = at(o_vals[0][i]); id.reference_wind_speed_values[i] = at(o_vals[1][i]);
        id.reference_humidity_values[i] = at(o_vals[2][i]); id.cloud_cover_values[i] = at(o_vals[3][i]);
    }
    id.hour_solar_noon = at(o_noon); id.hour_angle_sunrise = at(o_ha); id.direct_horizontal = at(o_dir); id.diffuse_horizontal = at(o_dif);
    id.global_radiation_hourly = at(o_grh); id.pressure_hourly = at(o_ph); id.pressure_site = at(o_ps); id.day_of_year = doy_host;
    mc_weather_out_soa od; std::memset(&od, 0, sizeof(od));
    od.pressure = at(o_wx[0]); od.reference_temperature = at(o_wx[1]); od.reference_wind_speed = at(o_wx[2]); od.reference_humidity = at(o_wx[3]);
    od.cloud_cover = at(o_wx[4]); od.global_radiation = at(o_wx[5]); od.diffuse_fraction = at(o_wx[6]);
    if (own_zenith) { od.zenith_angle = at(o_wx[7]); od.zenith_slope_angle = at(o_wx[8]); }
    od.hour_solar_noon = at(o_wnoon); od.hour_angle_sunrise = at(o_wha);
    mcw::WArgs a;
    if (int rc = make_args(h, ndays, w, sites ? &sd : nullptr, &id, &od, st->doy, a)) return rc;

    cudaStream_t s = mc_i_stream(h);
    WX_CUDA(h, cudaMemcpyAsync(st->doy, doy_host, D * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    for (const Plane& pl : planes) if (pl.host_in) WX_CUDA(h, copy_plane(pl, true, S, pitch, col0, s));
    if (out_end > out_begin) WX_CUDA(h, cudaMemsetAsync(base + out_begin, 0, (out_end - out_begin) * sizeof(double), s));   /* cells the path never writes read back as zero */
    if (int rc = launch(h, w, a)) return rc;

    mc_forcing_soa fd; std::memset(&fd, 0, sizeof(fd));
    fd.pressure = at(o_wx[0]); fd.reference_temperature = at(o_wx[1]); fd.reference_wind_speed = at(o_wx[2]); fd.reference_humidity = at(o_wx[3]);
    fd.cloud_cover = at(o_wx[4]); fd.global_radiation = at(o_wx[5]); fd.diffuse_fraction = at(o_wx[6]); fd.zenith_angle = at(o_wx[7]); fd.zenith_slope_angle = at(o_wx[8]);
    fd.rainfall_hourly = at(o_rainh); fd.longwave_radiation = at(o_lw);
    fd.leaf_area_index = at(o_daily[0]); fd.shade = at(o_daily[1]); fd.surface_emissivity = at(o_daily[2]); fd.cloud_emissivity = at(o_daily[3]);
    fd.soil_wetness = at(o_daily[4]); fd.deep_soil_temperature = at(o_daily[5]); fd.rainfall_daily = at(o_daily[6]); fd.albedo_daily = at(o_daily[7]);
    fd.prescribed_soil_moisture = at(o_presc); fd.day_of_year = doy_host;
    mc_output_soa odv; std::memset(&odv, 0, sizeof(odv));
    odv.soil_temperature = at(o_node[0]); odv.soil_moisture = at(o_node[1]); odv.soil_water_potential = at(o_node[2]); odv.soil_humidity = at(o_node[3]);
    odv.soil_thermal_conductivity = at(o_node[4]); odv.soil_heat_capacity = at(o_node[5]); odv.soil_bulk_density = at(o_node[6]);
    odv.surface_water = at(o_scal[0]); odv.sky_temperature = at(o_scal[1]); odv.snow_fall = at(o_scal[2]); odv.snow_depth = at(o_scal[3]); odv.snow_density = at(o_scal[4]);
    odv.snow_temperature = at(o_snowT);
    odv.air_temperature = at(o_prof[0]); odv.wind_speed = at(o_prof[1]); odv.relative_humidity = at(o_prof[2]);
    odv.convective_heat_flux = at(o_profs[0]); odv.friction_velocity = at(o_profs[1]);
    /* same stream: the soil kernel starts when the weather kernel has written its forcing */
    if (int rc = mc_step_days_device(h, day0, ndays, &fd, &odv)) return rc;
    for (const Plane& pl : planes) if (pl.host_out) WX_CUDA(h, copy_plane(pl, false, S, pitch, col0, s));
    return 0;
}

/* wait for the chunk, fetch the status words (into out->status at this handle's columns when given); returns the flagged-site count */
int minmax_finish(mc_handle* h, mc_output_soa* out, size_t col0)
{
    if (int rc = mc_synchronize(h)) return rc;
    const size_t S = (size_t)mc_i_sites(h);
    std::vector<int32_t> own;
    int32_t* status = out->status ? out->status + col0 : nullptr;
    if (!status) { own.resize(S); status = own.data(); }
    if (int rc = mc_get_status(h, status)) return rc;
    int nbad = 0;
    for (size_t i = 0; i < S; i++) if (status[i]) nbad++;
    return nbad;
}

}  // namespace

extern "C" {

/* solve!(cache) of a min/max-driven problem: expand the weather on the device into the soil kernel's forcing arrays, integrate, copy the outputs back */
int mc_step_days_minmax(mc_handle* h, int64_t day0, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                        const mc_weather_in_soa* in, const mc_forcing_soa* rest, mc_output_soa* out, mc_weather_out_soa* weather_host)
{
    if (!h) return -1;
    const size_t S = (size_t)mc_i_sites(h);
    if (int rc = minmax_enqueue(h, day0, ndays, w, sites, in, rest, out, weather_host, S, 0)) { if (rc < 0) cudaStreamSynchronize(mc_i_stream(h)); return rc; }
    return minmax_finish(h, out, 0);
}

/* the same over the devices of a multi handle: every device expands and integrates its contiguous column range of ONE set of host arrays; all devices are
 * enqueued before the first one is waited for, so their kernels run concurrently */
int mc_multi_step_days_minmax(mc_multi* m, int64_t day0, int32_t ndays, const mc_weather_desc* w, const mc_solar_sites_soa* sites,
                              const mc_weather_in_soa* in, const mc_forcing_soa* rest, mc_output_soa* out, mc_weather_out_soa* weather_host)
{
    if (!m) return -1;
    const int G = mc_multi_device_count(m);
    int64_t S_total = 0;
    for (int g = 0; g < G; g++) { int64_t lo, hi; mc_multi_range(m, g, &lo, &hi); if (hi > S_total) S_total = hi; }
    int rc_all = 0, enq = 0;
    for (; enq < G; enq++) {
        int64_t lo, hi; mc_multi_range(m, enq, &lo, &hi);
        mc_handle* h = mc_multi_handle(m, enq);
        if (hi == lo) continue;
        const int rc = h ? minmax_enqueue(h, day0, ndays, w, sites, in, rest, out, weather_host, (size_t)S_total, (size_t)lo) : -1;
        if (rc < 0) { rc_all = mc_i_multi_fail(m, enq, h ? mc_last_error(h) : "no handle", rc); break; }
    }
    int total = 0;
    for (int g = 0; g < (rc_all < 0 ? enq : G); g++) {                 /* on an error: still drain what was enqueued, nothing may keep touching the caller's arrays */
        int64_t lo, hi; mc_multi_range(m, g, &lo, &hi);
        if (hi == lo) continue;
        mc_handle* h = mc_multi_handle(m, g);
        const int rc = minmax_finish(h, out, (size_t)lo);
        if (rc < 0 && rc_all == 0) rc_all = mc_i_multi_fail(m, g, mc_last_error(h), rc);
        else if (rc > 0) total += rc;
    }
    return rc_all < 0 ? rc_all : total;
}

}  // extern "C"
