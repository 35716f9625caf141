/*
 * mc_api.cu — C ABI of the B200 soil engine (include/microclimate_b200.h) + the CUDA kernels.
 *
 * One handle drives one GPU.  Site parameters and the carried per-site state live in HBM as
 * site-fastest SoA arrays for the lifetime of the handle; mc_step_days streams a chunk of days of
 * forcing in, launches ONE persistent thread-per-site kernel for the whole chunk, and streams the
 * selected output rows back.  There is no CPU compute path in this library: without a CUDA
 * device every entry point fails with an error string.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>

#include "mc_device.cuh"

namespace {

thread_local std::string g_create_error;

#define MC_CUDA(h, call)                                                                                       \
    do {                                                                                                       \
        cudaError_t _e = (call);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                                      \
            return -100 - (int)_e;                                                                             \
        }                                                                                                      \
    } while (0)

/* 64-thread blocks, 16 resident per SM = 32 warps/SM at 64 registers/thread: the measured optimum on B200 (occupancy sweep in
 * profiles/README.md: 16 / 20 / 24 / 32 / 40 / 48 / 64 warps per SM -> 92 / 90 / 94 / 105 / 103 / 99 / 93 k site-days/s) */
#ifndef MC_BLOCK
#define MC_BLOCK 64
#endif
#ifndef MC_MIN_BLOCKS
#define MC_MIN_BLOCKS 16
#endif
constexpr int BLOCK = MC_BLOCK;

/* The flattened model lives in ONE __constant__ bank per device (mcd::g_model_c).  Handles with different models may share a device, and
 * mc_step_days_device is asynchronous, so an upload must never land while another handle's kernel is still reading the bank: the resident
 * model of every device is remembered here, uploads happen only when it differs from the launching handle's, after the device has drained,
 * and the check + upload + launch of a chunk are one critical section per device (model_guard).  Same-model handles never upload or wait. */
constexpr int MAX_DEVICES = 64;
std::mutex g_model_mutex[MAX_DEVICES];
mcd::DevModel g_model_resident[MAX_DEVICES];
bool g_model_valid[MAX_DEVICES];

}  // namespace

/* ---------------------------------------------------------------- kernels ---- */
template <int NA, int NS>
__global__ void __launch_bounds__(BLOCK, MC_MIN_BLOCKS) soil_chunk_kernel(const mcd::KArgs a)
{
    mcd::Counters cn = {};
    /* grid-stride over sites: the grid may be sized to the machine (blocks per SM x SM count) instead of to the problem */
    for (int64_t s = a.s_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < a.s_end; s += (int64_t)gridDim.x * blockDim.x)
        mcd::solve_site<NA, NS>(a, s, cn);
    if (a.counters) {
        atomicAdd(a.counters + 0, (unsigned long long)cn.n_rhs);
        atomicAdd(a.counters + 1, (unsigned long long)cn.n_accept);
        atomicAdd(a.counters + 2, (unsigned long long)cn.n_reject);
        atomicAdd(a.counters + 3, (unsigned long long)cn.n_obukhov);
        atomicAdd(a.counters + 4, (unsigned long long)cn.n_newton);
        atomicAdd(a.counters + 5, (unsigned long long)cn.n_infil);
        atomicAdd(a.counters + 6, (unsigned long long)cn.n_day_passes);
    }
}

/* Block-cooperative variant (NoSnow, fixed day-pass counts): a block owns blockDim.x sites; the conductivity of every RHS evaluation is
 * and every other node loop of the ODE hour are served by the whole block over (requesting site, node) items (mc_coop_ode.cuh,
 * "Block-cooperative ODE phase").  Threads past the end of
 * the range mirror the last site (same values to the same addresses) so that every thread of a block reaches every barrier. */
#ifndef MC_COOP_MIN_BLOCKS
#define MC_COOP_MIN_BLOCKS 1
#endif
#define MC_COOP_ROWS(NA) mcd::Coop2Rows<NA>::value
template <int NA>
__global__ void __launch_bounds__(mcd::COOP_TB, MC_COOP_MIN_BLOCKS) soil_chunk_coop_kernel(const mcd::KArgs a)
{
    __shared__ mcd::Coop2Shared sh;
    __shared__ mcd::MoistShared ms;
    if (threadIdx.x == 0) { for (int k = 0; k < 3; k++) { sh.count[k] = 0; ms.countU[k] = 0; ms.countF[k] = 0; ms.alive[k] = 0; } }
    int rot_ode = 0, rot_moist = 0;        /* rotation of the list counters: block-uniform, carried across site batches */
    __syncthreads();
    mcd::Counters cn = {};
    double* scr = a.coop_scr + (size_t)blockIdx.x * (MC_COOP_ROWS(NA) * blockDim.x);
    /* contiguous site batches (a warp-interleaved assignment that gives every block the whole range's mix of sites was measured 12 % slower: blocks
     * of neighbouring, climatically similar sites need fewer rounds per hour, profiles/README.md) */
    for (int64_t base = a.s_begin + (int64_t)blockIdx.x * blockDim.x; base < a.s_end; base += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = base + threadIdx.x;
        const bool valid = s < a.s_end;
        if (!valid) s = a.s_end - 1;
        mcd::Counters c1 = {};
        mcd::solve_site_coop2<NA>(a, s, c1, sh, ms, scr, rot_ode, rot_moist);
        if (valid) {
            cn.n_rhs += c1.n_rhs; cn.n_accept += c1.n_accept; cn.n_reject += c1.n_reject; cn.n_obukhov += c1.n_obukhov;
            cn.n_newton += c1.n_newton; cn.n_infil += c1.n_infil; cn.n_day_passes += c1.n_day_passes;
        }
    }
    if (a.counters) {
        atomicAdd(a.counters + 0, (unsigned long long)cn.n_rhs);
        atomicAdd(a.counters + 1, (unsigned long long)cn.n_accept);
        atomicAdd(a.counters + 2, (unsigned long long)cn.n_reject);
        atomicAdd(a.counters + 3, (unsigned long long)cn.n_obukhov);
        atomicAdd(a.counters + 4, (unsigned long long)cn.n_newton);
        atomicAdd(a.counters + 5, (unsigned long long)cn.n_infil);
        atomicAdd(a.counters + 6, (unsigned long long)cn.n_day_passes);
    }
}

/* Asynchronous kernel (mc_async.cuh): one persistent block per SM keeps SLOTS sites resident in shared memory; its warps are generic workers that
 * claim node items / single-lane tasks of whatever sites offer them.  Serves every NoSnow model with the water-potential infiltration form. */
#ifndef MC_ASYNC_THREADS
#define MC_ASYNC_THREADS 1024
#endif
template <int NA, int SLOTS>
__global__ void __launch_bounds__(MC_ASYNC_THREADS, 1) soil_async_kernel(const __grid_constant__ mcd::KArgs a)
{
    extern __shared__ __align__(16) unsigned char mc_async_smem[];
    mcd::AsyncBlock<NA, SLOTS>& B = *reinterpret_cast<mcd::AsyncBlock<NA, SLOTS>*>(mc_async_smem);
    mcd::async_block_init<NA, SLOTS>(B, (int)threadIdx.x, (int)blockDim.x, a.async_slots, a.async_ser_min, a.async_affinity);
    __syncthreads();
    double* ga = a.async_scr + (size_t)blockIdx.x * ((size_t)mcd::G_ARRAYS * NA * SLOTS);
    mcd::async_worker<NA, SLOTS>(B, a, ga, (int)(threadIdx.x >> 5));
}

/* solve_air! for the rows of the chunk the soil kernel just wrote (reference src/simulation.jl:822-851) */
__global__ void __launch_bounds__(128) air_profile_kernel(const mcd::KArgs a)
{
    mcd::Counters cn = {};
    for (int64_t s = a.s_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < a.s_end; s += (int64_t)gridDim.x * blockDim.x)
        mcd::solve_air_site(a, s, cn);
    if (a.counters && cn.n_obukhov) atomicAdd(a.counters + 3, (unsigned long long)cn.n_obukhov);
}

/* fresh state == what init() + the top of solve_soil! establish (reference src/simulation.jl:141-243,383-452) */
__global__ void init_state_kernel(mcd::KArgs a, int N, int n_snow, mc_snow_desc sm)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.S) return;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = 0; i < N; i++) {
        const int64_t k = (int64_t)i * a.S + s;
        a.T[k] = a.has_init_T ? a.init_T[k] : qnan;
        a.theta[k] = a.init_theta[k];
        a.sum_phase[k] = 0.0;
    }
    for (int i = 0; i < N + n_snow; i++) a.fsal[(int64_t)i * a.S + s] = qnan;
    if (n_snow > 0) {              /* init_snow_temperatures + initial_snow_state(model, depth, density) + reset_snow_scratch! */
        for (int i = 0; i < n_snow; i++) { a.snow_T[(int64_t)i * a.S + s] = sm.initial_snow_temperature_K; a.snow_nd[(int64_t)i * a.S + s] = 0.0; }
        const double dflt = sm.density_function[0] > 0.0 ? sm.density_function[1] : sm.snow_density;
        const double d = sm.initial_snow_density > 0.0 ? sm.initial_snow_density : dflt;
        const double v[MC_SNOW_STATE_FIELDS] = {sm.initial_snow_depth_cm, 0.0, 0.3, d, d, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < MC_SNOW_STATE_FIELDS; k++) a.snow_state[(int64_t)k * a.S + s] = v[k];
        a.q_freeze[s] = 0.0;
    }
    a.pool[s] = 0.0; a.wetness[s] = 0.0; a.L_ode[s] = -0.3; a.L_prof[s] = -0.3; a.L_air[s] = -0.3; a.tdeep_prev[s] = qnan; a.flags[s] = 0;
}

/* DFMA throughput micro-benchmark: 8 independent FMA chains per thread */
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

/* Headroom probe (not on the product path): the hottest function of the soil right-hand side, soil_conductivity, evaluated over a flat
 * array of independent (site, node) items — full lanes, no divergence, a tiny instruction window.  Its rate against the rate the same
 * function achieves inside soil_chunk_kernel quantifies what a block-cooperative, node-parallel formulation could gain. */
__global__ void __launch_bounds__(256) conductivity_items_kernel(const double* __restrict__ T, const double* __restrict__ theta, double* __restrict__ out, int64_t n, int reps)
{
    const mcd::DevModel& m = mcd::g_model_c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double th = theta[i], phim = 0.5, phig = mcd::jmax(0.0, 1.0 - th - phim), fw = 0.5, lam_m = 2.5;
        double Ti = T[i], acc = 0.0;
        for (int r = 0; r < reps; r++) {
            acc += mcd::soil_conductivity(Ti, 101325.0, 1.0, 1.0, th, phim, phig, fw, lam_m);
            Ti += 1e-3;
        }
        out[i] = acc + m.ga * 0.0;
    }
}

/* ---------------------------------------------------------------- handle ---- */
struct mc_handle {
    mc_model_desc model;
    mcd::DevModel dm;
    int64_t S = 0;
    int device = 0;
    int N = 0;
    cudaStream_t stream = nullptr;
    /* host-buffer pipeline (mc_step_days_async): copy-in, compute and copy-out streams over two staging sets */
    static constexpr int DEPTH = 2;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_h2d[DEPTH] = {nullptr, nullptr}, ev_k[DEPTH] = {nullptr, nullptr}, ev_d2h[DEPTH] = {nullptr, nullptr};
    cudaEvent_t ev_t0[DEPTH] = {nullptr, nullptr}, ev_t1[DEPTH] = {nullptr, nullptr};
    bool busy[DEPTH] = {false, false};                 /* a chunk enqueued on staging set b has not been harvested yet */
    int32_t* status_dst[DEPTH] = {nullptr, nullptr};   /* where that chunk's status words land on the host */
    std::vector<int32_t> status_own[DEPTH];
    uint64_t chunk_seq = 0;
    int last_flagged = 0; double pending_ms = 0.0; bool timed_by_ev01 = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    /* device arrays (all one arena) */
    double* arena = nullptr; size_t arena_doubles = 0;
    double *z0, *slope, *svf, *albedo, *bd, *md, *lam_m, *c_m, *psi_e, *ks, *bb, *rootd, *init_T, *init_theta;
    double *T, *theta, *sum_phase, *pool, *wetness, *L_ode, *L_prof, *tdeep_prev, *fsal;
    double *snow_T = nullptr, *snow_state = nullptr, *snow_nd = nullptr, *q_freeze = nullptr, *L_air = nullptr;
    int32_t* doy_dev[2] = {nullptr, nullptr}; size_t doy_cap[2] = {0, 0};      /* day-of-year vector of the chunk, one per staging set */
    int32_t* flags = nullptr;
    unsigned long long* counters = nullptr;
    bool has_init_T = false, sites_set = false;
    int64_t days_done = 0;
    /* chunk staging (grown on demand) */
    double* stage_in[DEPTH] = {nullptr, nullptr}; size_t stage_in_doubles[DEPTH] = {0, 0};
    double* stage_out[DEPTH] = {nullptr, nullptr}; size_t stage_out_doubles[DEPTH] = {0, 0};
    double* coop_scr = nullptr; size_t coop_scr_doubles = 0;      /* per-block scratch of the block-cooperative kernel */
    double* async_scr = nullptr; size_t async_scr_doubles = 0;    /* per-block L2-resident scratch of the asynchronous kernel */
    double last_ms = 0.0; unsigned long long last_launches = 0;
};

static int pick_na(int N) { return N <= 10 ? 10 : (N <= 19 ? 19 : MC_MAX_SOIL_NODES); }

static int validate_model(const mc_model_desc* m, std::string& err)
{
    if (!m) { err = "model is NULL"; return -1; }
    if (m->n_soil < 3 || m->n_soil > MC_MAX_SOIL_NODES) { err = "n_soil must be in 3.." + std::to_string(MC_MAX_SOIL_NODES); return -1; }
    if (m->n_snow != 0 && m->n_snow != MC_SNOW_NODES) { err = "n_snow must be 0 (NoSnow) or 8 (SnowModel)"; return -3; }
    if (!(m->moisture_timestep_s > 0.0) || m->moisture_timestep_s > 3600.0) { err = "moisture_timestep_s must be in (0, 3600]"; return -1; }
    if (m->n_heights < 0 || m->n_heights > MC_MAX_HEIGHTS) { err = "n_heights must be in 0.." + std::to_string(MC_MAX_HEIGHTS); return -1; }
    for (int i = 1; i < m->n_soil; i++) if (!(m->depths_m[i] > m->depths_m[i - 1])) { err = "depths must be strictly increasing"; return -1; }
    return 0;
}

template <int NA, int NS>
static cudaError_t launch_chunk(const mcd::KArgs& a, cudaStream_t st)
{
    int grid = (int)((a.s_end - a.s_begin + BLOCK - 1) / BLOCK);
    static int per_sm = -1, n_sm = 0;    /* development knob: MC_BLOCKS_PER_SM caps resident blocks by sizing the grid (persistent, grid-stride) */
    if (per_sm < 0) {
        const char* e = getenv("MC_BLOCKS_PER_SM");
        per_sm = e ? atoi(e) : 0;
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    if (per_sm > 0 && grid > per_sm * n_sm) grid = per_sm * n_sm;
    static int pad = -1;                 /* development knob: dynamic shared-memory padding limits resident blocks per SM */
    if (pad < 0) {
        const char* e = getenv("MC_SMEM_PAD");
        pad = e ? atoi(e) : 0;
        if (pad > 48 * 1024) cudaFuncSetAttribute(soil_chunk_kernel<NA, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
    }
    soil_chunk_kernel<NA, NS><<<grid, BLOCK, pad, st>>>(a);
    return cudaGetLastError();
}

/* the block-cooperative kernel serves NoSnow models whose day-pass count is the same for every site (FixedIterations) */
static bool coop_enabled(const mc_handle* h)
{
    const char* e = getenv("MC_KERNEL");                           /* MC_KERNEL=legacy: per-site kernel for every model (A/B and test knob) */
    const bool legacy = e && std::strcmp(e, "legacy") == 0;
    return !legacy && h->model.n_snow == 0 && h->dm.conv_kind == MC_CONV_FIXED;
}

template <int NA>
static int launch_chunk_coop(mc_handle* h, mcd::KArgs& a, cudaStream_t st)
{
    static int per_sm = -1, n_sm = 0;
    if (per_sm < 0) {
        const char* e = getenv("MC_BLOCKS_PER_SM");
        per_sm = e ? atoi(e) : 0;
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    /* block size: COOP_TB threads (one block per SM) when the range fills the machine, else the range spread over all SMs */
    const int64_t n = a.s_end - a.s_begin;
    int TB = mcd::COOP_TB;
    if (n < (int64_t)n_sm * TB) {
        const int64_t per = (n + n_sm - 1) / n_sm;
        TB = (int)std::min<int64_t>(mcd::COOP_TB, std::max<int64_t>(128, (per + 31) / 32 * 32));
    }
    if (const char* e = getenv("MC_COOP_BLOCK")) TB = std::min(mcd::COOP_TB, std::max(32, atoi(e) / 32 * 32));      /* test knob: force the block size */
    int grid = (int)((n + TB - 1) / TB);
    /* more blocks than 8 per SM run as a grid-stride loop over batches of TB sites: bounds the scratch (2.8 MB per block at N = 19) */
    const int cap = (per_sm > 0 ? per_sm : 8) * n_sm;
    if (grid > cap) grid = cap;
    if (const char* e = getenv("MC_COOP_GRID")) grid = std::max(1, std::min(grid, atoi(e)));                          /* test knob: force the grid-stride loop */
    const size_t need = (size_t)grid * MC_COOP_ROWS(NA) * TB;
    if (h->coop_scr_doubles < need) {
        MC_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->coop_scr) { MC_CUDA(h, cudaFree(h->coop_scr)); h->coop_scr = nullptr; h->coop_scr_doubles = 0; }
        MC_CUDA(h, cudaMalloc(&h->coop_scr, need * sizeof(double)));
        h->coop_scr_doubles = need;
    }
    a.coop_scr = h->coop_scr;
    soil_chunk_coop_kernel<NA><<<grid, TB, 0, st>>>(a);
    MC_CUDA(h, cudaGetLastError());
    return 0;
}

/* The asynchronous shared-memory-resident kernel (mc_async.cuh) can serve every NoSnow model whose infiltration step is the water-potential form.
 * It is OPT-IN (MC_KERNEL=async): measured on B200 it keeps the working set on chip (DRAM traffic ~ algorithmic) but runs at 25-29 k site-days/s
 * against the block-synchronous cooperative kernel's ~200 k, because 64 resident sites per SM cannot hide the serial per-site chains and warps that
 * run different code out of phase starve on instruction fetch (profiles/README.md, round 2). */
static bool async_enabled(const mc_handle* h)
{
    const char* e = getenv("MC_KERNEL");
    if (!(e && std::strcmp(e, "async") == 0)) return false;
    return h->model.n_snow == 0 && !h->dm.flux_alg;
}

template <int NA, int SLOTS>
static int launch_chunk_async(mc_handle* h, mcd::KArgs& a, cudaStream_t st, int slice)
{
    static int n_sm = 0; static bool attr_set = false;
    constexpr size_t smem = sizeof(mcd::AsyncBlock<NA, SLOTS>);
    if (!attr_set) {
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        MC_CUDA(h, cudaFuncSetAttribute(soil_async_kernel<NA, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        /* shared memory exactly as large as the tile needs: whatever is left of the 256 KB stays L1 (the stack frames of the math calls live there) */
        MC_CUDA(h, cudaFuncSetAttribute(soil_async_kernel<NA, SLOTS>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<size_t>(100, (smem + 1024) * 100 / (228 * 1024) + 1)));
        attr_set = true;
    }
    const int64_t n = a.s_end - a.s_begin;
    int grid = (int)std::min<int64_t>(n_sm, n);
    if (const char* e = getenv("MC_ASYNC_GRID")) grid = std::max(1, std::min(grid, atoi(e)));                           /* test knob */
    int slots = (int)std::min<int64_t>(SLOTS, (n + grid - 1) / grid);                                                   /* small ranges: spread the sites over all SMs */
    if (const char* e = getenv("MC_ASYNC_SLOTS")) slots = std::max(1, std::min((int)SLOTS, atoi(e)));                   /* test / tuning knob */
    int threads = MC_ASYNC_THREADS;
    if (const char* e = getenv("MC_ASYNC_THREADS")) threads = std::max(32, std::min(MC_ASYNC_THREADS, atoi(e) / 32 * 32));
    const size_t need = (size_t)grid * mcd::G_ARRAYS * NA * SLOTS;
    if (h->async_scr_doubles < need) {
        MC_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->async_scr) { MC_CUDA(h, cudaFree(h->async_scr)); h->async_scr = nullptr; h->async_scr_doubles = 0; }
        MC_CUDA(h, cudaMalloc(&h->async_scr, need * sizeof(double)));
        h->async_scr_doubles = need;
    }
    a.async_scr = h->async_scr; a.async_slots = slots;
    { const char* e = getenv("MC_ASYNC_SERMIN"); a.async_ser_min = e ? atoi(e) : 0; }                                  /* tuning knob */
    { const char* e = getenv("MC_ASYNC_IDLE_NS"); a.async_idle_ns = e ? atoi(e) : 0; }
    { const char* e = getenv("MC_ASYNC_AFFINITY"); a.async_affinity = e ? atoi(e) : 0; }
    a.async_stats = getenv("MC_ASYNC_STATS") ? h->counters + 16 : nullptr;                                             /* work census (development) */
    a.next_site = h->counters + 8 + slice;
    MC_CUDA(h, cudaMemsetAsync(a.next_site, 0, sizeof(unsigned long long), st));
    soil_async_kernel<NA, SLOTS><<<grid, threads, smem, st>>>(a);
    MC_CUDA(h, cudaGetLastError());
    return 0;
}

static void fill_site_state_args(mc_handle* h, mcd::KArgs& a)
{
    std::memset(&a, 0, sizeof(a));
    a.S = h->S; a.s_begin = 0; a.s_end = h->S; a.has_init_T = h->has_init_T;
    a.z0 = h->z0; a.slope = h->slope; a.svf = h->svf; a.albedo = h->albedo; a.bd = h->bd; a.md = h->md; a.lam_m = h->lam_m; a.c_m = h->c_m;
    a.psi_e = h->psi_e; a.ks = h->ks; a.bb = h->bb; a.rootd = h->rootd; a.init_T = h->init_T; a.init_theta = h->init_theta;
    a.T = h->T; a.theta = h->theta; a.sum_phase = h->sum_phase; a.pool = h->pool; a.wetness = h->wetness; a.L_ode = h->L_ode;
    a.L_prof = h->L_prof; a.tdeep_prev = h->tdeep_prev; a.fsal = h->fsal; a.flags = h->flags; a.counters = h->counters;
    a.snow_T = h->snow_T; a.snow_state = h->snow_state; a.snow_nd = h->snow_nd; a.q_freeze = h->q_freeze;
    a.L_air = h->L_air; a.n_snow = h->model.n_snow;
}

/* ---------------------------------------------------------------- C ABI ---- */
extern "C" {

int mc_abi_version(void) { return MC_ABI_VERSION; }
const char* mc_backend_name(void) { return "cuda-sm_100a"; }

void mc_default_fluid(mc_fluid_constants* f)
{
    f->molar_mass_water = 0.018016; f->molar_mass_air = 0.028966; f->ideal_gas_departure = 1.0053; f->cp_dry_air = 1004.84;
    f->cp_water_vapour = 1846.40; f->gas_constant = 8.314462618; f->stefan_boltzmann = 5.670374419e-8; f->standard_gravity = 9.80665;
    f->standard_pressure = 101325.0; f->latent_heat_a = 2.5012e6; f->latent_heat_b = 2.3787e3; f->molar_latent_heat_a = 45144.0;
    f->molar_latent_heat_b = 48.0;
}

/* MicroModel()/MicroConfig() defaults (reference src/types.jl:71-120, src/constants.jl:8) */
void mc_default_model(mc_model_desc* m)
{
    std::memset(m, 0, sizeof(*m));
    static const double dep_cm[19] = {0.0, 1.25, 2.5, 3.75, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0, 25.0, 30.0, 40.0, 50.0, 75.0, 100.0, 150.0, 200.0};
    m->n_soil = 19; m->n_snow = 0;
    for (int i = 0; i < 19; i++) m->depths_m[i] = dep_cm[i] / 100.0;
    m->reference_height_m = 2.0; m->local_height_m = 0.01;
    m->reltol = 1e-6; m->abstol = 1e-3; m->maximum_surface_temperature_K = 273.15 + 85.0;
    m->de_vries_shape_factor = 0.1; m->recirculation_power = 4.0; m->return_flow_threshold = 0.162;
    m->root_resistance = 2.5e10; m->stomatal_closure_potential = -1500.0; m->leaf_resistance = 2.0e6;
    m->stomatal_stability_parameter = 10.0; m->root_radius = 0.001;
    m->infiltration_algorithm = MC_INFIL_MATRIC_POTENTIAL;
    m->karman_constant = 0.4; m->dyer_constant = 16.0;
    m->time_mode = MC_TIME_CONSECUTIVE; m->iterations_per_day = 1; m->spinup_first_day = 0;
    m->convergence_kind = MC_CONV_FIXED; m->convergence_iterations = 3; m->convergence_tolerance_K = 0.0;
    m->rainfall_schedule = MC_RAIN_DAILY; m->moisture_strategy = MC_MOIST_PRESCRIBED;
    m->moisture_tolerance = 1e-6; m->moisture_max_iterations = 500; m->moisture_timestep_s = 360.0;
    m->max_surface_pool = 10000.0; m->atmospheric_radiation = MC_ATMRAD_CAMPBELL_NORMAN;
    mc_default_fluid(&m->fluid);
}

const char* mc_last_error(const mc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

}  // extern "C"
/* accessors for the library's other translation unit (mc_weather.cu: the widening rows); hidden, not part of the ABI */
#include "mc_internal.h"
int mc_i_device(const mc_handle* h) { return h->device; }
cudaStream_t mc_i_stream(const mc_handle* h) { return h->stream; }
int64_t mc_i_sites(const mc_handle* h) { return h->S; }
const mc_model_desc* mc_i_model(const mc_handle* h) { return &h->model; }
int mc_i_fail(mc_handle* h, const char* msg, int rc) { h->err = msg; return rc; }
extern "C" {

int mc_create(const mc_model_desc* model, int64_t n_sites, int device, mc_handle** out)
{
    if (out) *out = nullptr;
    int rc = validate_model(model, g_create_error);
    if (rc) return rc;
    if (n_sites <= 0 || !out) { g_create_error = "n_sites must be positive and out non-NULL"; return -1; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU path";
        return -2;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return -1; }
    mc_handle* h = new mc_handle();
    h->model = *model; h->S = n_sites; h->device = device; h->N = model->n_soil;
    mcd::build_dev_model(*model, h->dm);
    auto fail = [&](const char* what, cudaError_t ce) { g_create_error = std::string(what) + ": " + cudaGetErrorString(ce); mc_destroy(h); return -100 - (int)ce; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    for (int k = 0; k < mc_handle::DEPTH; k++) {
        if ((e = cudaEventCreateWithFlags(&h->ev_h2d[k], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&h->ev_k[k], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&h->ev_d2h[k], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaEventCreate(&h->ev_t0[k])) != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaEventCreate(&h->ev_t1[k])) != cudaSuccess) return fail("cudaEventCreate", e);
    }
    const size_t S = (size_t)n_sites, N = (size_t)h->N;
    const size_t NSn = (size_t)model->n_snow;
    const size_t per_site = 4 + 5 + 1, per_node = 8 + 2 + 4;
    h->arena_doubles = S * (per_site + per_node * N + (NSn ? NSn * 3 + MC_SNOW_STATE_FIELDS + 1 : 0));
    if ((e = cudaMalloc(&h->arena, h->arena_doubles * sizeof(double))) != cudaSuccess) return fail("cudaMalloc(state arena)", e);
    double* p = h->arena;
    auto take = [&](size_t n) { double* q = p; p += n; return q; };
    h->z0 = take(S); h->slope = take(S); h->svf = take(S); h->albedo = take(S);
    h->bd = take(N * S); h->md = take(N * S); h->lam_m = take(N * S); h->c_m = take(N * S);
    h->psi_e = take(N * S); h->ks = take(N * S); h->bb = take(N * S); h->rootd = take(N * S);
    h->init_T = take(N * S); h->init_theta = take(N * S);
    h->T = take(N * S); h->theta = take(N * S); h->sum_phase = take(N * S); h->fsal = take((N + NSn) * S);
    if (NSn) { h->snow_T = take(NSn * S); h->snow_nd = take(NSn * S); h->snow_state = take(MC_SNOW_STATE_FIELDS * S); h->q_freeze = take(S); }
    h->pool = take(S); h->wetness = take(S); h->L_ode = take(S); h->L_prof = take(S); h->tdeep_prev = take(S); h->L_air = take(S);
    if ((e = cudaMalloc(&h->flags, S * sizeof(int32_t))) != cudaSuccess) return fail("cudaMalloc(flags)", e);
    if ((e = cudaMalloc(&h->counters, (16 + mcd::MCA_STATS) * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc(counters)", e);
    if ((e = cudaMemset(h->counters, 0, (16 + mcd::MCA_STATS) * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMemset", e);
    *out = h;
    return 0;
}

void mc_destroy(mc_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    mc_i_weather_release(h);                           /* staging of mc_step_days_minmax / mc_populate_weather (mc_weather.cu) */
    if (h->s_in) cudaStreamSynchronize(h->s_in);
    if (h->s_out) cudaStreamSynchronize(h->s_out);
    cudaFree(h->arena); cudaFree(h->doy_dev[0]); cudaFree(h->doy_dev[1]); cudaFree(h->flags); cudaFree(h->counters); cudaFree(h->coop_scr); cudaFree(h->async_scr);
    for (int k = 0; k < mc_handle::DEPTH; k++) { cudaFree(h->stage_in[k]); cudaFree(h->stage_out[k]); }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int k = 0; k < mc_handle::DEPTH; k++) {
        cudaEvent_t evs[5] = {h->ev_h2d[k], h->ev_k[k], h->ev_d2h[k], h->ev_t0[k], h->ev_t1[k]};
        for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

/* [rows] x width bytes between a pitched and a packed array.  Contiguous on both sides (the single-handle case: the host arrays are exactly as wide as the handle)
 * is ONE linear copy: round 1's linear copies of these buffers moved ~56 GB/s over PCIe, round 2's 2-D copies of the same contiguous buffers ~18 GB/s (the exposed
 * copy time in the e2e figures of profiles/r1_bench_1gpu_v10.json vs r2_bench_1gpu_a.json); the 2-D form is kept for genuinely strided multi-device slices */
static cudaError_t copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaMemcpyKind kind, cudaStream_t st)
{
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, st);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
}

/* sites of this handle = columns [col0, col0 + S) of host arrays that are `host_pitch` doubles wide (a multi-device job shares one set of arrays) */
static int set_sites_pitched(mc_handle* h, const mc_sites_soa* s, const mc_initial_soa* init, int64_t host_pitch, int64_t col0)
{
    if (!s || !init || !init->initial_soil_moisture) { h->err = "sites/init must be non-NULL and initial_soil_moisture given"; return -1; }
    MC_CUDA(h, cudaSetDevice(h->device));
    if (mc_wait(h) < 0) return -1;
    const size_t S = (size_t)h->S, N = (size_t)h->N, bS = S * sizeof(double), hp = (size_t)host_pitch * sizeof(double);
    const double* site_src[4] = {s->roughness_height, s->slope, s->sky_view_fraction, s->albedo};
    double* site_dst[4] = {h->z0, h->slope, h->svf, h->albedo};
    const double* node_src[8] = {s->bulk_density, s->mineral_density, s->mineral_conductivity, s->mineral_heat_capacity,
                                 s->air_entry_water_potential, s->saturated_hydraulic_conductivity, s->campbell_b_parameter, s->root_density};
    double* node_dst[8] = {h->bd, h->md, h->lam_m, h->c_m, h->psi_e, h->ks, h->bb, h->rootd};
    for (int k = 0; k < 4; k++) { if (!site_src[k]) { h->err = "mc_sites_soa has a NULL site field"; return -1; } MC_CUDA(h, cudaMemcpyAsync(site_dst[k], site_src[k] + col0, bS, cudaMemcpyHostToDevice, h->stream)); }
    for (int k = 0; k < 8; k++) { if (!node_src[k]) { h->err = "mc_sites_soa has a NULL node field"; return -1; } MC_CUDA(h, copy_rows(node_dst[k], bS, node_src[k] + col0, hp, bS, N, cudaMemcpyHostToDevice, h->stream)); }
    h->has_init_T = init->initial_soil_temperature != nullptr;
    if (h->has_init_T) MC_CUDA(h, copy_rows(h->init_T, bS, init->initial_soil_temperature + col0, hp, bS, N, cudaMemcpyHostToDevice, h->stream));
    MC_CUDA(h, copy_rows(h->init_theta, bS, init->initial_soil_moisture + col0, hp, bS, N, cudaMemcpyHostToDevice, h->stream));
    mcd::KArgs a; fill_site_state_args(h, a);
    init_state_kernel<<<(unsigned)((S + 255) / 256), 256, 0, h->stream>>>(a, h->N, h->model.n_snow, h->model.snow);
    MC_CUDA(h, cudaGetLastError());
    MC_CUDA(h, cudaMemsetAsync(h->counters, 0, 8 * sizeof(unsigned long long), h->stream));
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    h->days_done = 0; h->sites_set = true;
    return 0;
}

int mc_set_sites(mc_handle* h, const mc_sites_soa* s, const mc_initial_soa* init)
{
    if (!h) return -1;
    return set_sites_pitched(h, s, init, h->S, 0);
}

static int copy_state(mc_handle* h, const mc_state_soa* st, bool to_host)
{
    MC_CUDA(h, cudaSetDevice(h->device));
    const size_t S = (size_t)h->S, N = (size_t)h->N, NSn = (size_t)h->model.n_snow, bS = S * sizeof(double), bNS = N * bS;
    struct Item { double* host; double* dev; size_t bytes; };
    const Item items[] = {{st->soil_temperature, h->T, bNS}, {st->soil_moisture, h->theta, bNS}, {st->sum_phase, h->sum_phase, bNS}, {st->fsal, h->fsal, (N + NSn) * bS},
                          {st->pool, h->pool, bS}, {st->soil_wetness, h->wetness, bS}, {st->obukhov_ode, h->L_ode, bS}, {st->obukhov_profile, h->L_prof, bS},
                          {st->deep_temperature_prev, h->tdeep_prev, bS},
                          {st->snow_temperature, h->snow_T, NSn * bS}, {st->snow_state, h->snow_state, MC_SNOW_STATE_FIELDS * bS},
                          {st->snow_node_depths, h->snow_nd, NSn * bS}, {st->q_freeze, h->q_freeze, bS}, {st->obukhov_air, h->L_air, bS}};
    for (const Item& it : items) {
        if (!it.host || !it.dev) continue;
        if (to_host) MC_CUDA(h, cudaMemcpyAsync(it.host, it.dev, it.bytes, cudaMemcpyDeviceToHost, h->stream));
        else MC_CUDA(h, cudaMemcpyAsync(it.dev, it.host, it.bytes, cudaMemcpyHostToDevice, h->stream));
    }
    if (st->status) {                            /* the accumulated MC_STATUS_* bits travel with the checkpoint */
        if (to_host) MC_CUDA(h, cudaMemcpyAsync(st->status, h->flags, S * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        else MC_CUDA(h, cudaMemcpyAsync(h->flags, st->status, S * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    }
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (st->days_done) { if (to_host) st->days_done[0] = h->days_done; else h->days_done = st->days_done[0]; }
    return 0;
}
int mc_get_state(mc_handle* h, mc_state_soa* out) { if (!h || !out) return -1; return copy_state(h, out, true); }
int mc_set_state(mc_handle* h, const mc_state_soa* in) { if (!h || !in) return -1; return copy_state(h, in, false); }

static int check_forcing(mc_handle* h, const mc_forcing_soa* f)
{
    const void* req[] = {f->pressure, f->reference_temperature, f->reference_wind_speed, f->reference_humidity, f->cloud_cover, f->global_radiation,
                         f->diffuse_fraction, f->zenith_angle, f->zenith_slope_angle, f->leaf_area_index, f->shade, f->surface_emissivity,
                         f->cloud_emissivity, f->soil_wetness, f->deep_soil_temperature, f->rainfall_daily};
    for (const void* p : req) if (!p) { h->err = "mc_forcing_soa: a required field is NULL"; return -1; }
    if (h->dm.rain_hourly && !f->rainfall_hourly) { h->err = "HourlyRainfall needs rainfall_hourly"; return -1; }
    return 0;
}

/* argument validation + KArgs for a chunk whose forcing/output pointers are DEVICE pointers; uploads the model constants and the
 * day-of-year vector on the handle's main stream */
static int prepare_chunk(mc_handle* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o, mcd::KArgs& a, int b = 0)
{
    if (!h->sites_set) { h->err = "mc_set_sites has not been called"; return -1; }
    if (!f || !o || ndays <= 0) { h->err = "forcing/out must be non-NULL and ndays positive"; return -1; }
    if (day0 != h->days_done) { h->err = "day0 must equal the number of days already integrated"; return -2; }
    if (check_forcing(h, f)) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    fill_site_state_args(h, a);
    a.day0 = day0; a.ndays = ndays;
    a.has_albedo_daily = f->albedo_daily != nullptr; a.has_lw = f->longwave_radiation != nullptr; a.has_presc = f->prescribed_soil_moisture != nullptr;
    a.f_P = f->pressure; a.f_T = f->reference_temperature; a.f_wind = f->reference_wind_speed; a.f_rh = f->reference_humidity;
    a.f_cloud = f->cloud_cover; a.f_solar = f->global_radiation; a.f_diffuse = f->diffuse_fraction; a.f_zen = f->zenith_angle;
    a.f_szen = f->zenith_slope_angle; a.f_rain = f->rainfall_hourly; a.f_lw = f->longwave_radiation;
    a.d_lai = f->leaf_area_index; a.d_shade = f->shade; a.d_eps = f->surface_emissivity; a.d_epsc = f->cloud_emissivity; a.d_wet = f->soil_wetness;
    a.d_tdeep = f->deep_soil_temperature; a.d_rain = f->rainfall_daily; a.d_albedo = f->albedo_daily; a.presc = f->prescribed_soil_moisture;
    a.o_T = o->soil_temperature; a.o_theta = o->soil_moisture; a.o_psi = o->soil_water_potential; a.o_h = o->soil_humidity;
    a.o_lam = o->soil_thermal_conductivity; a.o_cp = o->soil_heat_capacity; a.o_rho = o->soil_bulk_density; a.o_pool = o->surface_water; a.o_sky = o->sky_temperature;
    a.o_snow_fall = o->snow_fall; a.o_snow_depth = o->snow_depth; a.o_snow_dens = o->snow_density; a.o_snow_T = o->snow_temperature;
    a.o_air_T = o->air_temperature; a.o_wind = o->wind_speed; a.o_rh = o->relative_humidity; a.o_qc = o->convective_heat_flux; a.o_ustar = o->friction_velocity;
    if (a.o_air_T || a.o_wind || a.o_rh || a.o_qc || a.o_ustar) {
        if (h->model.n_heights < 1) { h->err = "solve_air! outputs requested but the model has no heights"; return -1; }
        if (!a.o_T || (h->model.n_snow > 0 && (!a.o_snow_depth || !a.o_snow_T))) { h->err = "solve_air! outputs need soil_temperature (and snow_depth + snow_temperature with SnowModel) selected"; return -1; }
    }
    if (f->day_of_year) {                /* MicroProblem.days: always a HOST pointer, staged here (used by the snow melt-factor / Q_freeze guards) */
        if (h->doy_cap[b] < (size_t)ndays) {
            MC_CUDA(h, cudaStreamSynchronize(h->stream));
            if (h->doy_dev[b]) MC_CUDA(h, cudaFree(h->doy_dev[b]));
            MC_CUDA(h, cudaMalloc(&h->doy_dev[b], (size_t)ndays * sizeof(int32_t))); h->doy_cap[b] = (size_t)ndays;
        }
        MC_CUDA(h, cudaMemcpyAsync(h->doy_dev[b], f->day_of_year, (size_t)ndays * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        a.doy = h->doy_dev[b];
    }
    return 0;
}

/* must be called with g_model_mutex[h->device] held, immediately before the launches that read the constant bank */
static int ensure_model_resident(mc_handle* h)
{
    const int d = h->device;
    if (g_model_valid[d] && std::memcmp(&g_model_resident[d], &h->dm, sizeof(mcd::DevModel)) == 0) return 0;
    MC_CUDA(h, cudaDeviceSynchronize());                                  /* kernels of other handles may still read the old model */
    MC_CUDA(h, cudaMemcpyToSymbol(mcd::g_model_c, &h->dm, sizeof(mcd::DevModel), 0, cudaMemcpyHostToDevice));
    g_model_resident[d] = h->dm; g_model_valid[d] = true;
    return 0;
}

static int launch_range(mc_handle* h, mcd::KArgs a, int64_t lo, int64_t hi, cudaStream_t st, int slice = 0)
{
    a.s_begin = lo; a.s_end = hi;
    const int na = pick_na(h->N);
    cudaError_t e = cudaSuccess;
    if (async_enabled(h)) {
        /* slices of one async scratch would overlap: concurrent slices share h->async_scr only because slice kernels are serialised (mc_step_days) */
        const char* se = getenv("MC_ASYNC_SLOTS");
        const bool half = se && atoi(se) > 0 && atoi(se) <= 32;      /* tuning knob: the 32-slot build (half the shared memory, more L1) */
        int rc;
        if (na == 10) rc = half ? launch_chunk_async<10, 32>(h, a, st, slice) : launch_chunk_async<10, 64>(h, a, st, slice);
        else if (na == 19) rc = half ? launch_chunk_async<19, 32>(h, a, st, slice) : launch_chunk_async<19, 64>(h, a, st, slice);
        else rc = half ? launch_chunk_async<MC_MAX_SOIL_NODES, 32>(h, a, st, slice) : launch_chunk_async<MC_MAX_SOIL_NODES, 48>(h, a, st, slice);
        if (rc) return rc;
    }
    else if (coop_enabled(h)) {
        if (int rc = na == 10 ? launch_chunk_coop<10>(h, a, st) : (na == 19 ? launch_chunk_coop<19>(h, a, st) : launch_chunk_coop<MC_MAX_SOIL_NODES>(h, a, st))) return rc;
    }
    else if (h->model.n_snow == 0) e = na == 10 ? launch_chunk<10, 0>(a, st) : (na == 19 ? launch_chunk<19, 0>(a, st) : launch_chunk<MC_MAX_SOIL_NODES, 0>(a, st));
    else e = na == 10 ? launch_chunk<10, MC_SNOW_NODES>(a, st) : (na == 19 ? launch_chunk<19, MC_SNOW_NODES>(a, st) : launch_chunk<MC_MAX_SOIL_NODES, MC_SNOW_NODES>(a, st));
    MC_CUDA(h, e);
    if (a.o_air_T || a.o_wind || a.o_rh || a.o_qc || a.o_ustar) {
        const int grid = (int)((hi - lo + 127) / 128);
        air_profile_kernel<<<grid, 128, 0, st>>>(a);
        MC_CUDA(h, cudaGetLastError());
    }
    return 0;
}

int mc_step_days_device(mc_handle* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o)
{
    if (!h) return -1;
    mcd::KArgs a;
    if (int rc = prepare_chunk(h, day0, ndays, f, o, a, (int)(h->chunk_seq++ & 1))) return rc;
    std::lock_guard<std::mutex> model_guard(g_model_mutex[h->device % MAX_DEVICES]);
    if (int rc = ensure_model_resident(h)) return rc;
    MC_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (int rc = launch_range(h, a, 0, h->S, h->stream)) return rc;
    MC_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    h->last_launches = 1; h->timed_by_ev01 = true;
    h->days_done += ndays;
    return 0;
}

/* the asynchronous kernel gives up (instead of hanging the device) when a warp sees no work for ~70 s; counters[7] counts such blocks */
static int check_watchdog(mc_handle* h)
{
    unsigned long long v = 0;
    MC_CUDA(h, cudaMemcpy(&v, h->counters + 7, sizeof(v), cudaMemcpyDeviceToHost));
    if (v) { h->err = "soil_async_kernel watchdog: " + std::to_string(v) + " block(s) stopped after ~70 s without work; results of this chunk are incomplete"; return -5; }
    return 0;
}

/* a chunk on staging set b is complete on the device: collect its kernel time and flagged-site count */
static int harvest(mc_handle* h, int b)
{
    if (!h->busy[b]) return 0;
    MC_CUDA(h, cudaEventSynchronize(h->ev_d2h[b]));
    if (int rc = check_watchdog(h)) { h->busy[b] = false; return rc; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_t0[b], h->ev_t1[b]) == cudaSuccess) h->pending_ms += ms;
    int nbad = 0;
    const int32_t* st = h->status_dst[b];
    for (int64_t i = 0; i < h->S; i++) if (st[i]) nbad++;
    h->last_flagged = nbad;
    h->busy[b] = false;
    return 0;
}

int mc_synchronize(mc_handle* h)
{
    if (!h) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (h->timed_by_ev01 && h->last_launches && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_ms = ms;
    return check_watchdog(h);
}

static int ensure_stage(mc_handle* h, double** buf, size_t* have, size_t need)
{
    if (*have >= need) return 0;
    /* growing a staging set: nothing may still be using the old one */
    MC_CUDA(h, cudaStreamSynchronize(h->s_in)); MC_CUDA(h, cudaStreamSynchronize(h->stream)); MC_CUDA(h, cudaStreamSynchronize(h->s_out));
    if (*buf) { MC_CUDA(h, cudaFree(*buf)); *buf = nullptr; *have = 0; }
    MC_CUDA(h, cudaMalloc(buf, need * sizeof(double)));
    *have = need;
    return 0;
}

/* Host-buffer chunk, asynchronous.  Three streams over two staging sets:
 *     copy-in :  H2D(forcing q)                      waits for kernel q-2 (same staging set)
 *     compute :  zero outputs, kernel q              waits for H2D q and for D2H q-2
 *     copy-out:  D2H(outputs q, status)              waits for kernel q
 * so the H2D of chunk q+1 and the D2H of chunk q-1 overlap the kernel of chunk q.  Host arrays may be wider than the handle (a multi-device
 * job: `host_pitch` doubles per row, this handle's sites start at column `col0`); site-fastest SoA makes that a strided 2-D copy per plane. */
static int step_days_enqueue(mc_handle* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o, int64_t host_pitch, int64_t col0)
{
    if (!f || !o || ndays <= 0) { h->err = "forcing/out must be non-NULL and ndays positive"; return -1; }
    if (check_forcing(h, f)) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    const int b = (int)(h->chunk_seq & 1);
    if (int rc = harvest(h, b)) return rc;                         /* at most DEPTH chunks in flight: reusing a staging set waits for its previous chunk */
    const size_t S = (size_t)h->S, N = (size_t)h->N, H = 24 * (size_t)ndays, D = (size_t)ndays;
    const size_t NSn = (size_t)h->model.n_snow;
    struct Plane { const double* host_in; double* host_out; double* dev; size_t rows; };   /* [rows][S] planes */
    std::vector<Plane> in_planes, out_planes;
    const double* hourly_src[11] = {f->pressure, f->reference_temperature, f->reference_wind_speed, f->reference_humidity, f->cloud_cover,
                                    f->global_radiation, f->diffuse_fraction, f->zenith_angle, f->zenith_slope_angle, f->rainfall_hourly, f->longwave_radiation};
    const double* daily_src[8] = {f->leaf_area_index, f->shade, f->surface_emissivity, f->cloud_emissivity, f->soil_wetness, f->deep_soil_temperature,
                                  f->rainfall_daily, f->albedo_daily};
    size_t need_in = 0;
    for (const double* p : hourly_src) if (p) need_in += H * S;
    for (const double* p : daily_src) if (p) need_in += D * S;
    if (f->prescribed_soil_moisture) need_in += D * N * S;
    if (int rc = ensure_stage(h, &h->stage_in[b], &h->stage_in_doubles[b], need_in)) return rc;
    mc_forcing_soa fd; std::memset(&fd, 0, sizeof(fd));
    const double** hourly_dst[11] = {&fd.pressure, &fd.reference_temperature, &fd.reference_wind_speed, &fd.reference_humidity, &fd.cloud_cover,
                                     &fd.global_radiation, &fd.diffuse_fraction, &fd.zenith_angle, &fd.zenith_slope_angle, &fd.rainfall_hourly, &fd.longwave_radiation};
    const double** daily_dst[8] = {&fd.leaf_area_index, &fd.shade, &fd.surface_emissivity, &fd.cloud_emissivity, &fd.soil_wetness, &fd.deep_soil_temperature,
                                   &fd.rainfall_daily, &fd.albedo_daily};
    double* p = h->stage_in[b];
    for (int k = 0; k < 11; k++) if (hourly_src[k]) { in_planes.push_back({hourly_src[k], nullptr, p, H}); *hourly_dst[k] = p; p += H * S; }
    for (int k = 0; k < 8; k++) if (daily_src[k]) { in_planes.push_back({daily_src[k], nullptr, p, D}); *daily_dst[k] = p; p += D * S; }
    if (f->prescribed_soil_moisture) { in_planes.push_back({f->prescribed_soil_moisture, nullptr, p, D * N}); fd.prescribed_soil_moisture = p; p += D * N * S; }
    fd.day_of_year = f->day_of_year;
    double* node_host[7] = {o->soil_temperature, o->soil_moisture, o->soil_water_potential, o->soil_humidity, o->soil_thermal_conductivity, o->soil_heat_capacity, o->soil_bulk_density};
    double* scal_host[5] = {o->surface_water, o->sky_temperature, NSn ? o->snow_fall : nullptr, NSn ? o->snow_depth : nullptr, NSn ? o->snow_density : nullptr};
    double* snowT_host = NSn ? o->snow_temperature : nullptr;
    size_t need_out = 0;
    for (double* q : node_host) if (q) need_out += H * N * S;
    for (double* q : scal_host) if (q) need_out += H * S;
    if (snowT_host) need_out += H * NSn * S;
    const size_t NH = (size_t)h->model.n_heights;
    double* prof_host[3] = {o->air_temperature, o->wind_speed, o->relative_humidity};
    double* profs_host[2] = {o->convective_heat_flux, o->friction_velocity};
    for (double* q2 : prof_host) if (q2) need_out += H * NH * S;
    for (double* q2 : profs_host) if (q2) need_out += H * S;
    if (int rc = ensure_stage(h, &h->stage_out[b], &h->stage_out_doubles[b], need_out ? need_out : 1)) return rc;
    mc_output_soa od; std::memset(&od, 0, sizeof(od));
    double** node_dev[7] = {&od.soil_temperature, &od.soil_moisture, &od.soil_water_potential, &od.soil_humidity, &od.soil_thermal_conductivity, &od.soil_heat_capacity, &od.soil_bulk_density};
    double** scal_dev[5] = {&od.surface_water, &od.sky_temperature, &od.snow_fall, &od.snow_depth, &od.snow_density};
    double** prof_dev[3] = {&od.air_temperature, &od.wind_speed, &od.relative_humidity};
    double** profs_dev[2] = {&od.convective_heat_flux, &od.friction_velocity};
    double* q = h->stage_out[b];
    for (int k = 0; k < 7; k++) if (node_host[k]) { out_planes.push_back({nullptr, node_host[k], q, H * N}); *node_dev[k] = q; q += H * N * S; }
    for (int k = 0; k < 5; k++) if (scal_host[k]) { out_planes.push_back({nullptr, scal_host[k], q, H}); *scal_dev[k] = q; q += H * S; }
    if (snowT_host) { out_planes.push_back({nullptr, snowT_host, q, H * NSn}); od.snow_temperature = q; q += H * NSn * S; }
    for (int k = 0; k < 3; k++) if (prof_host[k]) { out_planes.push_back({nullptr, prof_host[k], q, H * NH}); *prof_dev[k] = q; q += H * NH * S; }
    for (int k = 0; k < 2; k++) if (profs_host[k]) { out_planes.push_back({nullptr, profs_host[k], q, H}); *profs_dev[k] = q; q += H * S; }

    mcd::KArgs a;
    if (int rc = prepare_chunk(h, day0, ndays, &fd, &od, a, b)) return rc;
    const size_t w = S * sizeof(double), hp = (size_t)host_pitch * sizeof(double);
    /* copy-in */
    MC_CUDA(h, cudaStreamWaitEvent(h->s_in, h->ev_k[b], 0));      /* kernel q-2 has finished reading this staging set (no-op for a never-recorded event) */
    for (const Plane& pl : in_planes) MC_CUDA(h, copy_rows(pl.dev, w, pl.host_in + col0, hp, w, pl.rows, cudaMemcpyHostToDevice, h->s_in));
    MC_CUDA(h, cudaEventRecord(h->ev_h2d[b], h->s_in));
    /* compute */
    MC_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_h2d[b], 0));
    MC_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_d2h[b], 0));  /* D2H q-2 has drained this staging set's outputs */
    /* cells the path never writes (e.g. water potential rows under PrescribedSoilMoisture) read back as zero */
    if (need_out) MC_CUDA(h, cudaMemsetAsync(h->stage_out[b], 0, need_out * sizeof(double), h->stream));
    {
        std::lock_guard<std::mutex> model_guard(g_model_mutex[h->device % MAX_DEVICES]);
        if (int rc = ensure_model_resident(h)) return rc;
        MC_CUDA(h, cudaEventRecord(h->ev_t0[b], h->stream));
        if (int rc = launch_range(h, a, 0, h->S, h->stream)) return rc;
        MC_CUDA(h, cudaEventRecord(h->ev_t1[b], h->stream));
    }
    MC_CUDA(h, cudaEventRecord(h->ev_k[b], h->stream));
    /* copy-out */
    int32_t* status_host = o->status ? o->status + col0 : nullptr;
    if (!status_host) { h->status_own[b].resize(S); status_host = h->status_own[b].data(); }
    MC_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_k[b], 0));
    for (const Plane& pl : out_planes) MC_CUDA(h, copy_rows(pl.host_out + col0, hp, pl.dev, w, w, pl.rows, cudaMemcpyDeviceToHost, h->s_out));
    MC_CUDA(h, cudaMemcpyAsync(status_host, h->flags, S * sizeof(int32_t), cudaMemcpyDeviceToHost, h->s_out));
    MC_CUDA(h, cudaEventRecord(h->ev_d2h[b], h->s_out));
    h->status_dst[b] = status_host; h->busy[b] = true;
    h->chunk_seq++;
    h->last_launches = 1; h->timed_by_ev01 = false;
    h->days_done += ndays;
    return 0;
}

/* an enqueue that failed half-way may have copies into / out of caller buffers in flight: nothing may still touch them when the error is returned */
static int drain_after_error(mc_handle* h, int rc)
{
    if (rc < 0) {
        if (h->s_in) cudaStreamSynchronize(h->s_in);
        if (h->stream) cudaStreamSynchronize(h->stream);
        if (h->s_out) cudaStreamSynchronize(h->s_out);
        for (int k = 0; k < mc_handle::DEPTH; k++) h->busy[k] = false;
    }
    return rc;
}

int mc_step_days_async(mc_handle* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o)
{
    if (!h) return -1;
    return drain_after_error(h, step_days_enqueue(h, day0, ndays, f, o, h->S, 0));
}

/* wait for every chunk enqueued with mc_step_days_async; returns the flagged-site count of the most recent one */
int mc_wait(mc_handle* h)
{
    if (!h) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    const int first = (int)(h->chunk_seq & 1);                     /* older chunk first */
    for (int k = 0; k < mc_handle::DEPTH; k++) if (int rc = harvest(h, (first + k) & 1)) return rc;
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    h->last_ms = h->pending_ms; h->pending_ms = 0.0;
    return h->last_flagged;
}

/* Host-buffer entry point: one chunk, synchronous (enqueue + wait). */
int mc_step_days(mc_handle* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o)
{
    if (!h) return -1;
    if (int rc = mc_wait(h)) { if (rc < 0) return drain_after_error(h, rc); }
    if (int rc = step_days_enqueue(h, day0, ndays, f, o, h->S, 0)) return drain_after_error(h, rc);
    return drain_after_error(h, mc_wait(h));
}

int mc_get_counters(mc_handle* h, mc_counters* c)
{
    if (!h || !c) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    unsigned long long v[8];
    MC_CUDA(h, cudaMemcpyAsync(v, h->counters, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    std::memset(c, 0, sizeof(*c));
    c->n_rhs = v[0]; c->n_accept = v[1]; c->n_reject = v[2]; c->n_obukhov_iter = v[3]; c->n_newton_iter = v[4]; c->n_infil_steps = v[5]; c->n_day_passes = v[6];
    c->n_site_days = (uint64_t)(h->days_done * h->S);
    c->gpu_ms = h->last_ms; c->n_kernel_launches = h->last_launches;
    return 0;
}

/* development probe: the asynchronous kernel's work census since the last call (MC_ASYNC_STATS=1), see mcd::MCA_STATS */
int mc_get_async_stats(mc_handle* h, unsigned long long* out, int n)
{
    if (!h || !out || n < mcd::MCA_STATS) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    MC_CUDA(h, cudaMemcpyAsync(out, h->counters + 16, mcd::MCA_STATS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    MC_CUDA(h, cudaMemsetAsync(h->counters + 16, 0, mcd::MCA_STATS * sizeof(unsigned long long), h->stream));
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return mcd::MCA_STATS;
}

int mc_reset_counters(mc_handle* h)
{
    if (!h) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    MC_CUDA(h, cudaMemsetAsync(h->counters, 0, 8 * sizeof(unsigned long long), h->stream));
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mc_get_status(mc_handle* h, int32_t* status_host)
{
    if (!h || !status_host) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    MC_CUDA(h, cudaMemcpyAsync(status_host, h->flags, (size_t)h->S * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

/* development probe: evaluations per second of soil_conductivity over independent items (see conductivity_items_kernel) */
int mc_bench_conductivity(mc_handle* h, int64_t n_items, int reps, double* evals_per_s)
{
    if (!h || !evals_per_s || n_items <= 0 || reps <= 0) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    std::vector<double> T((size_t)n_items), th((size_t)n_items);
    for (int64_t i = 0; i < n_items; i++) { T[(size_t)i] = 275.0 + (double)(i % 5000) * 0.01; th[(size_t)i] = 0.05 + (double)(i % 3000) * 1e-4; }
    double *dT = nullptr, *dth = nullptr, *dout = nullptr;
    MC_CUDA(h, cudaMalloc(&dT, (size_t)n_items * 8)); MC_CUDA(h, cudaMalloc(&dth, (size_t)n_items * 8)); MC_CUDA(h, cudaMalloc(&dout, (size_t)n_items * 8));
    MC_CUDA(h, cudaMemcpyAsync(dT, T.data(), (size_t)n_items * 8, cudaMemcpyHostToDevice, h->stream));
    MC_CUDA(h, cudaMemcpyAsync(dth, th.data(), (size_t)n_items * 8, cudaMemcpyHostToDevice, h->stream));
    std::lock_guard<std::mutex> model_guard(g_model_mutex[h->device % MAX_DEVICES]);
    if (int rc = ensure_model_resident(h)) return rc;
    cudaDeviceProp prop; MC_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    const int blocks = prop.multiProcessorCount * 8;
    conductivity_items_kernel<<<blocks, 256, 0, h->stream>>>(dT, dth, dout, n_items, 1);
    MC_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    conductivity_items_kernel<<<blocks, 256, 0, h->stream>>>(dT, dth, dout, n_items, reps);
    MC_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    MC_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f; MC_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *evals_per_s = (double)n_items * reps / (ms * 1e-3);
    cudaFree(dT); cudaFree(dth); cudaFree(dout);
    return 0;
}

int mc_measure_fp64_peak(mc_handle* h, double* tflops)
{
    if (!h || !tflops) return -1;
    MC_CUDA(h, cudaSetDevice(h->device));
    cudaDeviceProp prop; MC_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double* buf = nullptr;
    MC_CUDA(h, cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
    dfma_peak_kernel<<<blocks, threads, 0, h->stream>>>(buf, 1024);                 /* warm-up */
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        MC_CUDA(h, cudaEventRecord(h->ev0, h->stream));
        dfma_peak_kernel<<<blocks, threads, 0, h->stream>>>(buf, iters);
        MC_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        MC_CUDA(h, cudaStreamSynchronize(h->stream));
        float ms = 0.f; MC_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    MC_CUDA(h, cudaFree(buf));
    *tflops = best;
    return 0;
}

/* ---------------------------------------------------------------- one host process, several GPUs ---- */
/* The reference idiom is one process sweeping sites (reinit!(cache, inputs); solve!(cache), src/simulation.jl:266-277).  A multi handle shards the
 * sites of ONE set of host arrays contiguously over several devices (replicated model, no collective); every device runs the asynchronous
 * pipeline above on its column range, so all GPUs copy and integrate concurrently from a single host thread. */
struct mc_multi { std::vector<mc_handle*> h; std::vector<int64_t> lo; int64_t S = 0; std::string err; };
}  // extern "C"
int mc_i_multi_fail(mc_multi* m, int device_index, const char* msg, int rc) { m->err = "device " + std::to_string(device_index) + ": " + (msg ? msg : ""); return rc; }
extern "C" {

int mc_multi_create(const mc_model_desc* model, int64_t n_sites, const int* device_ids, int n_devices, mc_multi** out)
{
    if (out) *out = nullptr;
    if (!out || !device_ids || n_devices <= 0 || n_sites < n_devices) { g_create_error = "mc_multi_create: need out, device_ids, n_devices >= 1 and at least one site per device"; return -1; }
    mc_multi* m = new mc_multi(); m->S = n_sites;
    for (int g = 0; g < n_devices; g++) {
        const int64_t lo = n_sites * g / n_devices, hi = n_sites * (g + 1) / n_devices;
        mc_handle* hg = nullptr;
        const int rc = mc_create(model, hi - lo, device_ids[g], &hg);
        if (rc) { mc_multi_destroy(m); return rc; }
        m->h.push_back(hg); m->lo.push_back(lo);
    }
    m->lo.push_back(n_sites);
    *out = m;
    return 0;
}
void mc_multi_destroy(mc_multi* m) { if (!m) return; for (mc_handle* hg : m->h) mc_destroy(hg); delete m; }
const char* mc_multi_last_error(const mc_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }
int mc_multi_device_count(const mc_multi* m) { return m ? (int)m->h.size() : -1; }
mc_handle* mc_multi_handle(mc_multi* m, int g) { return (m && g >= 0 && g < (int)m->h.size()) ? m->h[(size_t)g] : nullptr; }
int mc_multi_range(const mc_multi* m, int g, int64_t* lo, int64_t* hi)
{
    if (!m || g < 0 || g >= (int)m->h.size()) return -1;
    if (lo) *lo = m->lo[(size_t)g];
    if (hi) *hi = m->lo[(size_t)g + 1];
    return 0;
}
int mc_multi_set_sites(mc_multi* m, const mc_sites_soa* s, const mc_initial_soa* init)
{
    if (!m) return -1;
    for (size_t g = 0; g < m->h.size(); g++)
        if (int rc = set_sites_pitched(m->h[g], s, init, m->S, m->lo[g])) { m->err = "device " + std::to_string(g) + ": " + m->h[g]->err; return rc; }
    return 0;
}
int mc_multi_step_days_async(mc_multi* m, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o)
{
    if (!m) return -1;
    for (size_t g = 0; g < m->h.size(); g++)
        if (int rc = step_days_enqueue(m->h[g], day0, ndays, f, o, m->S, m->lo[g])) {
            m->err = "device " + std::to_string(g) + ": " + m->h[g]->err;
            for (size_t k = 0; k < m->h.size(); k++) drain_after_error(m->h[k], rc < 0 ? rc : -1);      /* the other devices' copies touch the same caller arrays */
            return rc;
        }
    return 0;
}
int mc_multi_wait(mc_multi* m)
{
    if (!m) return -1;
    int total = 0;
    for (size_t g = 0; g < m->h.size(); g++) {
        const int rc = mc_wait(m->h[g]);
        if (rc < 0) { m->err = "device " + std::to_string(g) + ": " + m->h[g]->err; return rc; }
        total += rc;
    }
    return total;
}
int mc_multi_step_days(mc_multi* m, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o)
{
    if (!m) return -1;
    if (int rc = mc_multi_wait(m)) { if (rc < 0) return rc; }
    if (int rc = mc_multi_step_days_async(m, day0, ndays, f, o)) return rc;
    return mc_multi_wait(m);
}

}  /* extern "C" */
