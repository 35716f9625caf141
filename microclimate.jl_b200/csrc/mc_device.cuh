/*
 * mc_device.cuh — device code of the batched-sites soil heat + moisture integration (sm_100a).
 *
 * B200-native restatement of Microclimate.jl's solve_soil! hot path (reference src/simulation.jl:355-819 and everything
 * below it).  This is NOT a translation of the Julia control flow:
 *   - theta is constant during an ODE hour, so every theta-only term of the Campbell-de Vries soil properties (areal heat
 *     capacity, volume fractions, recirculation factor) is hoisted out of the right-hand side into a per-hour table;
 *   - soil temperature is constant during the moisture hour, so the infiltration setup block (reference campbell.jl:180-248,
 *     recomputed 10x per hour upstream) and the per-node vapour densities inside the Newton loop are computed once per hour;
 *   - forcing is read straight from the site-fastest SoA arrays (a warp reads 32 consecutive doubles per knot: one 256 B
 *     coalesced request), never copied into per-site buffers;
 *   - in the block-cooperative kernel the node loops of the ODE hour are item passes of the whole block over (site, node)
 *     pairs and the integrator is a resumable state machine (mc_coop_ode.cuh).
 * Only results of whole sub-expressions are hoisted, never re-associated: every floating-point operation of the reference
 * algorithm is executed in the reference's order with the shared bit-reproducible math kernel (include/mc_math.h), so the
 * outputs are bit-identical to the CPU oracle (tests/).
 *
 * Layout of the device code (all in namespace mcd, included below in this order):
 *   mc_physics.cuh   psychrometrics, longwave, Monin-Obukhov, soil properties, no-snow RHS
 *   mc_snow.cuh      SnowModel{8} and the snow-aware RHS
 *   mc_tsit5.cuh     adaptive Tsit5 + PI controller, per-site form
 *   mc_coop_ode.cuh  block-cooperative ODE phase
 *   mc_moisture.cuh  freeze/thaw, Campbell (1985) soil water balance
 *   mc_site.cuh      solve_soil! / solve_air! per site, host-side model flattening
 *   mc_async.cuh     asynchronous shared-memory-resident solve_soil! (opt-in, MC_KERNEL=async: bit-identical, measured slower)
 *
 * The functions are written against a tiny platform shim so the very same code can be compiled by g++ for tests/hostemu (a
 * development harness that runs the kernel body on the CPU where no GPU exists).  The shipped library (mc_api.cu) only ever
 * launches the CUDA kernels.
 */
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/microclimate_b200.h"
#include "../../include/mc_math.h"

#ifndef MC_COOP_MOIST
#define MC_COOP_MOIST 0            /* 1: block-cooperative soil water balance (experimental: bit-identical but ~15 % slower than the per-site form, profiles/README.md) */
#endif
#if defined(__CUDACC__)
#define MC_DEV __device__ __forceinline__
#define MC_MEMBER __device__ __forceinline__
#define MC_DEV_NOINLINE __device__ __noinline__
#define MC_LDG(p) __ldg(p)
#define MC_ROLLED _Pragma("unroll 1")
#ifndef MC_TSIT5_ROLLED
#define MC_TSIT5_ROLLED 1       /* rolled stage loops: 2.2 k -> 0.9 k instructions, +5 % at 32 warps/SM (instruction-fetch bound) */
#endif
#define MC_PRAGMA_STR(x) _Pragma(#x)
#define MC_PRAGMA(x) MC_PRAGMA_STR(x)
#ifndef MC_THOMAS_UNROLL_N
#define MC_THOMAS_UNROLL_N 3       /* owner-side Thomas sweeps: loads of the next rows issue under the divisions of the current one */
#endif
#define MC_THOMAS_UNROLL MC_PRAGMA(unroll MC_THOMAS_UNROLL_N)
#if MC_TSIT5_ROLLED
#define MC_UNROLL4 _Pragma("unroll 1")
#else
#define MC_UNROLL4 _Pragma("unroll 4")
#endif
#define MC_CONST_MODEL g_model_c
#define MC_TID ((int)threadIdx.x)
#define MC_TB ((int)blockDim.x)
#define MC_BLOCK_SYNC() __syncthreads()
#else
#include <cmath>
#include <thread>
#define MC_DEV static inline
#define MC_MEMBER inline
#define MC_DEV_NOINLINE static __attribute__((noinline))
#define MC_LDG(p) (*(p))
#define MC_ROLLED
#define MC_UNROLL4
#define MC_THOMAS_UNROLL
#define MC_CONST_MODEL g_model_h
#ifdef MC_HOST_THREADS          /* tests/hostemu threaded mode: a "block" of OS threads with a real barrier, to exercise the block protocol on the CPU */
#include <barrier>
namespace mcd_host { inline thread_local int tl_tid = 0; inline int g_tb = 1; inline std::barrier<>* g_barrier = nullptr; }
#define MC_TID (mcd_host::tl_tid)
#define MC_TB (mcd_host::g_tb)
#define MC_BLOCK_SYNC() mcd_host::g_barrier->arrive_and_wait()
#else
#define MC_TID 0
#define MC_TB 1
#define MC_BLOCK_SYNC() ((void)0)
#endif
#endif

#ifndef MC_USE_GG3
#define MC_USE_GG3 0
#endif
#ifndef MC_LERP_LOOP
#define MC_LERP_LOOP 0
#endif
#ifndef MC_WETAIR_NOINLINE
#define MC_WETAIR_NOINLINE 0
#endif

namespace mcd {

constexpr int MAXN = MC_MAX_SOIL_NODES;
constexpr double T0C = 273.15;
constexpr double T_LO = 273.15 - 81.0;              /* heat_transport_1d.jl:159 */
constexpr double WATER_CP = 4184.0;                 /* src/constants.jl */
constexpr double WATER_RHO = 1000.0;
constexpr double LATENT_FUSION = 333550.0;
constexpr double MW_INFIL = 0.01801528;             /* campbell.jl:176 (differs from FluidProperties' 0.018016 on purpose) */
constexpr double DV_INFIL = 2.4e-5;                 /* campbell.jl:178 */
constexpr double RHOCPTKG = 6.003e-8 * 4.184 * 3600.0 * 1.0e8;   /* monin_obukhov.jl:135 */
constexpr double PI_D = 3.14159265358979323846;

/* ---- model constants, flattened once on the host (mc_api.cu) into constant memory ---- */
struct DevModel {
    int N, dynamic, nonconsec, conv_kind, conv_iters, iters_per_day, spinup, rain_hourly, flux_alg, swinbank;
    int moist_max_it, niter_moist;
    double depths[MAXN];
    double dzc[MAXN];       /* ld[i+1]-ld[i-1] (i>=1), ld[1] for i=0, with the zero sentinel at ld[N]   (heat_transport_1d.jl:175-210) */
    double dzg[MAXN];       /* ld[i+1]-ld[i] */
    double mdz[MAXN + 2], mmass[MAXN + 2];   /* infiltration grid: depth[i+1]-depth[i] and layer water mass (campbell.jl:228-230), index i = 1..L */
    double zref, zloc, reltol, abstol, T_hi;
    double ga, gc, recirc_pow, theta0;
    double root_res, PC, RL, SP, root_radius;
    double kappa, gamma, conv_tol, moist_tol, moist_dt, max_pool;
    /* FluidProperties constants */
    double Mw, eps, dep, cpa, cpv, R, sigma, g, P0, Lv_a, Lv_b, Hv_a, Hv_b, Ma_over_R;
    double schmidt_factor;  /* (0.71/0.60)^0.666, monin_obukhov.jl:406-407 */
    mc_snow_desc snow;      /* SnowModel (n_snow = 8) */
    int n_heights; double heights[MC_MAX_HEIGHTS];      /* MicroModel.heights (solve_air!) */
};

#if defined(__CUDACC__)
__constant__ DevModel g_model_c;
#else
static DevModel g_model_h;
#endif

/* ---- kernel arguments: every array is site-fastest SoA (x[node][site], f[hour][site]) ---- */
struct KArgs {
    int64_t S;        /* sites of the handle = stride of every SoA array */
    int64_t s_begin, s_end;   /* site range [s_begin, s_end) this launch integrates (pipelined site slices) */
    int64_t day0;     /* global index of the first day of the chunk */
    int32_t ndays;    /* days in the chunk */
    int32_t has_init_T, has_albedo_daily, has_lw, has_presc;
    /* sites */
    const double *z0, *slope, *svf, *albedo;
    const double *bd, *md, *lam_m, *c_m, *psi_e, *ks, *bb, *rootd;
    const double *init_T, *init_theta;
    /* carried state */
    double *T, *theta, *sum_phase, *pool, *wetness, *L_ode, *L_prof, *tdeep_prev, *fsal;
    int32_t* flags;
    double *snow_T, *snow_state, *snow_nd, *q_freeze;     /* SnowModel: [8][S], [MC_SNOW_STATE_FIELDS][S], [8][S], [S] */
    const int32_t* doy;                                   /* [ndays] day of year (device) */
    /* forcing (chunk-relative rows) */
    const double *f_P, *f_T, *f_wind, *f_rh, *f_cloud, *f_solar, *f_diffuse, *f_zen, *f_szen, *f_rain, *f_lw;
    const double *d_lai, *d_shade, *d_eps, *d_epsc, *d_wet, *d_tdeep, *d_rain, *d_albedo, *presc;
    /* outputs (NULL = deselected) */
    double *o_T, *o_theta, *o_psi, *o_h, *o_lam, *o_cp, *o_rho, *o_pool, *o_sky;
    double *o_snow_fall, *o_snow_depth, *o_snow_dens, *o_snow_T;
    double *o_air_T, *o_wind, *o_rh, *o_qc, *o_ustar;     /* solve_air! profiles [H][n_heights][S] and scalars [H][S] */
    double *L_air;                                        /* [S] Obukhov warm start of the air-profile buffer */
    int32_t n_snow;
    unsigned long long* counters;   /* [8]: rhs, accept, reject, obukhov, newton, infil, day_passes, spare */
    double* coop_scr;               /* block-cooperative kernel: per-block scratch [block][SCR_ARRAYS][NA][block threads] */
    double* async_scr;              /* asynchronous kernel: per-block L2-resident scratch [block][G_ARRAYS][NA][SLOTS] (site constants, hourly / daily copies) */
    unsigned long long* next_site;  /* asynchronous kernel: the launch's site cursor (slots take their next site from it) */
    int32_t async_slots;            /* asynchronous kernel: slots a block activates (fewer than its capacity when the range is small, so that every SM gets sites) */
    int32_t async_ser_min;          /* asynchronous kernel: single-lane tasks of one type that must be waiting before a warp prefers them to node items */
    int32_t async_affinity;         /* asynchronous kernel: 1 = warps prefer one task class per scheduler (instruction-cache locality) */
    int32_t async_idle_ns;          /* asynchronous kernel: sleep of a warp that found nothing to claim */
    unsigned long long* async_stats;/* asynchronous kernel: [MCA_STATS] work census (claims, items, cycles per task type; idle polls), or NULL */
};

struct Counters { unsigned int n_rhs, n_accept, n_reject, n_obukhov, n_newton, n_infil, n_day_passes; };

MC_DEV double jmax(double a, double b) { return (a != a || b != b) ? (a + b) : (a > b ? a : b); }   /* Julia max/min propagate NaN */
MC_DEV double jmin(double a, double b) { return (a != a || b != b) ? (a + b) : (a < b ? a : b); }
MC_DEV double jclamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

#include "mc_physics.cuh"
#include "mc_snow.cuh"
#include "mc_tsit5.cuh"
#include "mc_coop_ode.cuh"
#include "mc_moisture.cuh"
#include "mc_site.cuh"
#include "mc_async.cuh"

} // namespace mcd
