/* TEST HARNESS ONLY (CPU tier).  Compiles the product's weather translation unit — microclimate.jl_b200/csrc/mc_weather.cu: the kernel, the
 * argument checks, the staging plans of mc_populate_weather and mc_step_days_minmax — with g++ over a stand-in CUDA runtime in which device
 * memory is host memory, streams are synchronous and a launch is a loop over the thread grid.  The soil side of mc_step_days_minmax
 * (mc_step_days_device etc.) is served by the soil host emulation (hostemu.cpp, linked into the same library).  Nothing here ships. */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define MCW_HOST_EMULATION 1
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
static dim3 blockIdx, threadIdx;
#define __global__
#define __launch_bounds__(n)
#define __grid_constant__
static long g_fake_allocs = 0;
/* every live "device" allocation, so the harness can tell device memory from host memory the way a real GPU would (by faulting) */
struct FakeAlloc { const char* base; size_t bytes; };
static std::vector<FakeAlloc> g_fake_live;
static bool g_allow_foreign = false;                     /* the *_device entry points are handed caller-owned "device" memory */
static bool fake_is_device(const void* p)
{
    if (!p || g_allow_foreign) return true;
    for (const FakeAlloc& a : g_fake_live) if ((const char*)p >= a.base && (const char*)p < a.base + a.bytes) return true;
    return false;
}
static inline const char* cudaGetErrorString(cudaError_t) { return "fake"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n)
{
    *p = (T*)std::malloc(n ? n : 1);
    if (!*p) return 2;
    std::memset(*p, 0xA5, n);
    g_fake_allocs++; g_fake_live.push_back({(const char*)*p, n ? n : 1});
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p)
{
    if (!p) return cudaSuccess;
    g_fake_allocs--;
    for (size_t i = 0; i < g_fake_live.size(); i++) if (g_fake_live[i].base == (const char*)p) { g_fake_live.erase(g_fake_live.begin() + (long)i); break; }
    std::free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < height; r++) std::memcpy((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}

#include "../../include/microclimate_b200.h"

/* the soil host emulation (hostemu.cpp) */
extern "C" {
void* he_create(const mc_model_desc* m, int64_t S);
void he_destroy(void* h);
int he_set_sites(void* h, const mc_sites_soa* s, const mc_initial_soa* init);
int he_step_days(void* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o);
}

struct mc_handle { void* emu; int64_t S; mc_model_desc model; std::string err; std::vector<int32_t> status; };
/* a multi handle of the emulation: contiguous column ranges over several emulated "devices" (mc_api.cu's mc_multi, reduced to what mc_weather.cu uses) */
struct mc_multi { std::vector<mc_handle*> h; std::vector<int64_t> lo; int64_t S = 0; std::string err; };

#include "../../microclimate.jl_b200/csrc/mc_weather.cu"

namespace {
bool mcw_emulation_check(const mcw::WArgs& a)
{
    const void* ptrs[] = {a.doy, a.lat, a.slope, a.aspect, a.noon_off, a.noon_in, a.ha_in, a.direct, a.diffuse, a.gr_hourly, a.p_hourly, a.p_site,
                          a.o_P, a.o_T, a.o_wind, a.o_rh, a.o_cloud, a.o_gr, a.o_df, a.o_zen, a.o_szen, a.o_noon, a.o_ha};
    for (const void* p : ptrs) if (!fake_is_device(p)) return false;
    for (int q = 0; q < 4; q++) for (int i = 0; i < MC_DIEL_MAX_INPUTS; i++) if (!fake_is_device(a.vals[q][i])) return false;
    return true;
}
}

int mc_i_device(const mc_handle*) { return 0; }
cudaStream_t mc_i_stream(const mc_handle*) { return nullptr; }
int64_t mc_i_sites(const mc_handle* h) { return h->S; }
const mc_model_desc* mc_i_model(const mc_handle* h) { return &h->model; }
int mc_i_fail(mc_handle* h, const char* msg, int rc) { h->err = msg; return rc; }
int mc_i_multi_fail(mc_multi* m, int g, const char* msg, int rc) { m->err = "device " + std::to_string(g) + ": " + (msg ? msg : ""); return rc; }

extern "C" {
/* the entry points of mc_api.cu that mc_step_days_minmax calls, on the emulated soil engine ("device" pointers are host pointers here) */
int mc_wait(mc_handle*) { return 0; }
int mc_synchronize(mc_handle*) { return 0; }
int mc_step_days_device(mc_handle* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o)
{
    /* the real entry point hands these pointers to a kernel: every one must be device memory (day_of_year is the documented host pointer) */
    const void* fp[] = {f->pressure, f->reference_temperature, f->reference_wind_speed, f->reference_humidity, f->cloud_cover, f->global_radiation, f->diffuse_fraction,
                        f->zenith_angle, f->zenith_slope_angle, f->rainfall_hourly, f->longwave_radiation, f->leaf_area_index, f->shade, f->surface_emissivity,
                        f->cloud_emissivity, f->soil_wetness, f->deep_soil_temperature, f->rainfall_daily, f->albedo_daily, f->prescribed_soil_moisture,
                        o->soil_temperature, o->soil_moisture, o->soil_water_potential, o->soil_humidity, o->soil_thermal_conductivity, o->soil_heat_capacity,
                        o->soil_bulk_density, o->surface_water, o->sky_temperature, o->snow_fall, o->snow_depth, o->snow_density, o->snow_temperature,
                        o->air_temperature, o->wind_speed, o->relative_humidity, o->convective_heat_flux, o->friction_velocity};
    for (const void* p : fp) if (!fake_is_device(p)) { h->err = "emulation: mc_step_days_device was handed a pointer that is not device memory"; return -9; }
    if (fake_is_device(f->day_of_year) && !g_allow_foreign && f->day_of_year) { h->err = "emulation: day_of_year must stay a host pointer"; return -9; }
    mc_output_soa oo = *o;
    oo.status = h->status.data();
    const int rc = he_step_days(h->emu, day0, ndays, f, &oo);
    return rc < 0 ? rc : 0;
}
int mc_get_status(mc_handle* h, int32_t* out) { std::memcpy(out, h->status.data(), (size_t)h->S * sizeof(int32_t)); return 0; }

const char* mc_last_error(const mc_handle* h) { return h ? h->err.c_str() : ""; }
int mc_multi_device_count(const mc_multi* m) { return (int)m->h.size(); }
mc_handle* mc_multi_handle(mc_multi* m, int g) { return m->h[(size_t)g]; }
int mc_multi_range(const mc_multi* m, int g, int64_t* lo, int64_t* hi) { *lo = m->lo[(size_t)g]; *hi = m->lo[(size_t)g + 1]; return 0; }

mc_handle* wa_create(const mc_model_desc* m, int64_t S);
void wa_destroy(mc_handle* h);
mc_multi* wa_multi_create(const mc_model_desc* desc, int64_t S, int n_devices)
{
    mc_multi* m = new mc_multi(); m->S = S;
    const int64_t base = S / n_devices, rem = S % n_devices;
    int64_t lo = 0;
    for (int g = 0; g < n_devices; g++) { const int64_t n = base + (g < rem ? 1 : 0); m->lo.push_back(lo); m->h.push_back(n > 0 ? wa_create(desc, n) : nullptr); lo += n; }
    m->lo.push_back(S);
    return m;
}
void wa_multi_destroy(mc_multi* m) { if (!m) return; for (mc_handle* h : m->h) wa_destroy(h); delete m; }
const char* wa_multi_last_error(const mc_multi* m) { return m->err.c_str(); }

mc_handle* wa_create(const mc_model_desc* m, int64_t S)
{
    mc_handle* h = new mc_handle();
    h->emu = he_create(m, S); h->S = S; h->model = *m; h->status.assign((size_t)S, 0);
    if (!h->emu) { delete h; return nullptr; }
    return h;
}
void wa_destroy(mc_handle* h) { if (!h) return; mc_i_weather_release(h); he_destroy(h->emu); delete h; }
int wa_set_sites(mc_handle* h, const mc_sites_soa* s, const mc_initial_soa* i) { return he_set_sites(h->emu, s, i); }
const char* wa_last_error(const mc_handle* h) { return h->err.c_str(); }
long wa_live_allocations(void) { return g_fake_allocs; }
void wa_allow_foreign_pointers(int on) { g_allow_foreign = on != 0; }
}
