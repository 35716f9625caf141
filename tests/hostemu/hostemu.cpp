/*
 * hostemu.cpp — DEVELOPMENT/TEST HARNESS, not product code and not a fallback.
 *
 * Compiles the product's per-site kernel body (microclimate.jl_b200/csrc/mc_device.cuh) with g++ and
 * runs it site by site on the CPU, so the logic the CUDA kernel executes can be checked against the
 * oracle in the CPU-only test tier (`pytest -m "not gpu"`), where no GPU exists.  Nothing in the
 * product loads this library: libmicroclimate_b200.so has no CPU path and fails without a CUDA device.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <limits>
#ifdef MC_HOST_THREADS
#include <mutex>
#include <thread>
#endif
#include "../../microclimate.jl_b200/csrc/mc_device.cuh"

namespace {
#ifdef MC_HOST_THREADS
/* Threaded mode: one "block" = HE_TB OS threads sharing Coop2Shared / MoistShared / the block scratch, with a real barrier behind
 * MC_BLOCK_SYNC() — the block protocol of the cooperative kernel (request lists, counter parity, role swaps seen by other threads,
 * mirror threads of a ragged last block) exercised on the CPU. */
template <int NA>
void run_blocks_threaded(const mcd::KArgs& a, int TB, int grid, unsigned long long* counters)
{
    const int64_t n = a.s_end - a.s_begin;
    const int nbatch = (int)((n + TB - 1) / TB);
    if (grid <= 0 || grid > nbatch) grid = nbatch;                 /* grid < nbatch: grid-stride loop over batches, as with MC_COOP_GRID */
    std::vector<double> scr(mcd::Coop2Rows<NA>::value * (size_t)TB);
    std::mutex mu;
    for (int b = 0; b < grid; b++) {                               /* blocks one after the other (one barrier object) */
        static mcd::Coop2Shared sh; static mcd::MoistShared ms;
        std::memset(&sh, 0, sizeof(sh)); std::memset(&ms, 0, sizeof(ms));
        std::barrier<> bar(TB);
        mcd_host::g_tb = TB; mcd_host::g_barrier = &bar;
        std::vector<std::thread> th;
        for (int t = 0; t < TB; t++) th.emplace_back([&, t]() {
            mcd_host::tl_tid = t;
            int rot_ode = 0, rot_moist = 0;
            mcd::Counters cn = {};
            for (int64_t base = a.s_begin + (int64_t)b * TB; base < a.s_end; base += (int64_t)grid * TB) {
                int64_t s = base + t; const bool valid = s < a.s_end; if (!valid) s = a.s_end - 1;      /* mirror threads, as in the kernel */
                mcd::Counters c1 = {};
                mcd::solve_site_coop2<NA>(a, s, c1, sh, ms, scr.data(), rot_ode, rot_moist);
                if (valid) { cn.n_rhs += c1.n_rhs; cn.n_accept += c1.n_accept; cn.n_reject += c1.n_reject; cn.n_obukhov += c1.n_obukhov;
                             cn.n_newton += c1.n_newton; cn.n_infil += c1.n_infil; cn.n_day_passes += c1.n_day_passes; }
            }
            std::lock_guard<std::mutex> lk(mu);
            counters[0] += cn.n_rhs; counters[1] += cn.n_accept; counters[2] += cn.n_reject; counters[3] += cn.n_obukhov;
            counters[4] += cn.n_newton; counters[5] += cn.n_infil; counters[6] += cn.n_day_passes;
        });
        for (auto& x : th) x.join();
    }
}
#endif
#ifndef HE_SLOTS
#define HE_SLOTS 6
#endif
/* The asynchronous kernel body (mc_async.cuh): `grid` blocks run concurrently, each with `nwarps` worker "warps".  A warp is one OS thread (with
 * -DMC_HOST_THREADS; otherwise a single warp on the calling thread) that executes its MCA_LANES lanes one after the other; the request masks, cursor
 * words and retire counters are real atomics, so the claim / retire protocol runs under arbitrary interleavings. */
static long long g_async_runs = 0;
static unsigned long long g_async_stats[64];
template <int NA>
void run_async(mcd::KArgs a, int nwarps, int grid, unsigned long long* counters)
{
    g_async_runs++;
    using Blk = mcd::AsyncBlock<NA, HE_SLOTS>;
    unsigned long long cursor = 0, cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    a.next_site = &cursor; a.counters = cnt;
    a.async_stats = getenv("MC_ASYNC_STATS") ? g_async_stats : nullptr;
    { const char* e = getenv("MC_ASYNC_SERMIN"); a.async_ser_min = e ? atoi(e) : 0; }
    { const char* e = getenv("MC_ASYNC_AFFINITY"); a.async_affinity = e ? atoi(e) : 0; }
    std::vector<Blk*> blocks; std::vector<std::vector<double>> scr;
    for (int b = 0; b < grid; b++) {
        Blk* B = new Blk(); std::memset(B, 0, sizeof(Blk));
        for (int t = 0; t < HE_SLOTS; t++) mcd::async_block_init<NA, HE_SLOTS>(*B, t, HE_SLOTS, a.async_slots, a.async_ser_min, a.async_affinity);
        blocks.push_back(B); scr.emplace_back((size_t)mcd::G_ARRAYS * NA * HE_SLOTS, 0.0);
    }
#ifdef MC_HOST_THREADS
    std::vector<std::thread> th;
    for (int b = 0; b < grid; b++) for (int w = 0; w < nwarps; w++)
        th.emplace_back([&, b, w]() { mcd::async_worker<NA, HE_SLOTS>(*blocks[b], a, scr[b].data(), w); });
    for (auto& x : th) x.join();
#else
    (void)nwarps;
    for (int b = 0; b < grid; b++) mcd::async_worker<NA, HE_SLOTS>(*blocks[b], a, scr[b].data(), 0);
#endif
    for (auto* B : blocks) delete B;
    for (int k = 0; k < 7; k++) counters[k] += cnt[k];
}
struct Emu {
    mc_model_desc model; int64_t S; int N; int64_t days_done;
    std::vector<double> site[4], node[8], init_T, init_theta, T, theta, sum_phase, fsal, pool, wet, L_ode, L_prof, tdeep;
    std::vector<double> snow_T, snow_state, snow_nd, q_freeze, L_air;
    std::vector<int32_t> flags; bool has_init_T;
    unsigned long long counters[8];
};
}

extern "C" {

void* he_create(const mc_model_desc* m, int64_t S)
{
    if (!m || (m->n_snow != 0 && m->n_snow != MC_SNOW_NODES)) return nullptr;
    Emu* e = new Emu(); e->model = *m; e->S = S; e->N = m->n_soil; e->days_done = 0; std::memset(e->counters, 0, sizeof(e->counters));
    return e;
}
void he_destroy(void* h) { delete (Emu*)h; }

int he_set_sites(void* h, const mc_sites_soa* s, const mc_initial_soa* init)
{
    Emu* e = (Emu*)h; const size_t S = e->S, N = e->N;
    const double* ss[4] = {s->roughness_height, s->slope, s->sky_view_fraction, s->albedo};
    const double* ns[8] = {s->bulk_density, s->mineral_density, s->mineral_conductivity, s->mineral_heat_capacity, s->air_entry_water_potential,
                           s->saturated_hydraulic_conductivity, s->campbell_b_parameter, s->root_density};
    for (int k = 0; k < 4; k++) e->site[k].assign(ss[k], ss[k] + S);
    for (int k = 0; k < 8; k++) e->node[k].assign(ns[k], ns[k] + N * S);
    e->has_init_T = init->initial_soil_temperature != nullptr;
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    e->init_T.assign(N * S, qnan);
    if (e->has_init_T) e->init_T.assign(init->initial_soil_temperature, init->initial_soil_temperature + N * S);
    e->init_theta.assign(init->initial_soil_moisture, init->initial_soil_moisture + N * S);
    e->T = e->init_T; e->theta = e->init_theta; e->sum_phase.assign(N * S, 0.0); e->fsal.assign((N + e->model.n_snow) * S, qnan);
    if (e->model.n_snow) {
        const mc_snow_desc& sm = e->model.snow; const size_t NS = e->model.n_snow;
        e->snow_T.assign(NS * S, sm.initial_snow_temperature_K); e->snow_nd.assign(NS * S, 0.0); e->q_freeze.assign(S, 0.0);
        const double dflt = sm.density_function[0] > 0.0 ? sm.density_function[1] : sm.snow_density;
        const double d = sm.initial_snow_density > 0.0 ? sm.initial_snow_density : dflt;
        const double v[MC_SNOW_STATE_FIELDS] = {sm.initial_snow_depth_cm, 0.0, 0.3, d, d, 0.0, 0.0, 0.0, 0.0, 0.0};
        e->snow_state.assign(MC_SNOW_STATE_FIELDS * S, 0.0);
        for (int k = 0; k < MC_SNOW_STATE_FIELDS; k++) for (size_t q = 0; q < S; q++) e->snow_state[k * S + q] = v[k];
    }
    e->pool.assign(S, 0.0); e->wet.assign(S, 0.0); e->L_ode.assign(S, -0.3); e->L_prof.assign(S, -0.3); e->tdeep.assign(S, qnan);
    e->L_air.assign(S, -0.3);
    e->flags.assign(S, 0); e->days_done = 0; std::memset(e->counters, 0, sizeof(e->counters));
    return 0;
}

int he_step_days(void* h, int64_t day0, int32_t ndays, const mc_forcing_soa* f, mc_output_soa* o)
{
    Emu* e = (Emu*)h;
    if (day0 != e->days_done) return -2;
    mcd::build_dev_model(e->model, mcd::g_model_h);
    mcd::KArgs a; std::memset(&a, 0, sizeof(a));
    a.S = e->S; a.day0 = day0; a.ndays = ndays; a.has_init_T = e->has_init_T;
    a.has_albedo_daily = f->albedo_daily != nullptr; a.has_lw = f->longwave_radiation != nullptr; a.has_presc = f->prescribed_soil_moisture != nullptr;
    a.z0 = e->site[0].data(); a.slope = e->site[1].data(); a.svf = e->site[2].data(); a.albedo = e->site[3].data();
    a.bd = e->node[0].data(); a.md = e->node[1].data(); a.lam_m = e->node[2].data(); a.c_m = e->node[3].data(); a.psi_e = e->node[4].data();
    a.ks = e->node[5].data(); a.bb = e->node[6].data(); a.rootd = e->node[7].data(); a.init_T = e->init_T.data(); a.init_theta = e->init_theta.data();
    a.T = e->T.data(); a.theta = e->theta.data(); a.sum_phase = e->sum_phase.data(); a.pool = e->pool.data(); a.wetness = e->wet.data();
    a.L_ode = e->L_ode.data(); a.L_prof = e->L_prof.data(); a.tdeep_prev = e->tdeep.data(); a.fsal = e->fsal.data(); a.flags = e->flags.data();
    if (e->model.n_snow) { a.snow_T = e->snow_T.data(); a.snow_state = e->snow_state.data(); a.snow_nd = e->snow_nd.data(); a.q_freeze = e->q_freeze.data(); }
    a.doy = f->day_of_year;
    a.f_P = f->pressure; a.f_T = f->reference_temperature; a.f_wind = f->reference_wind_speed; a.f_rh = f->reference_humidity;
    a.f_cloud = f->cloud_cover; a.f_solar = f->global_radiation; a.f_diffuse = f->diffuse_fraction; a.f_zen = f->zenith_angle;
    a.f_szen = f->zenith_slope_angle; a.f_rain = f->rainfall_hourly; a.f_lw = f->longwave_radiation;
    a.d_lai = f->leaf_area_index; a.d_shade = f->shade; a.d_eps = f->surface_emissivity; a.d_epsc = f->cloud_emissivity; a.d_wet = f->soil_wetness;
    a.d_tdeep = f->deep_soil_temperature; a.d_rain = f->rainfall_daily; a.d_albedo = f->albedo_daily; a.presc = f->prescribed_soil_moisture;
    a.o_T = o->soil_temperature; a.o_theta = o->soil_moisture; a.o_psi = o->soil_water_potential; a.o_h = o->soil_humidity;
    a.o_lam = o->soil_thermal_conductivity; a.o_cp = o->soil_heat_capacity; a.o_rho = o->soil_bulk_density; a.o_pool = o->surface_water; a.o_sky = o->sky_temperature;
    a.o_snow_fall = o->snow_fall; a.o_snow_depth = o->snow_depth; a.o_snow_dens = o->snow_density; a.o_snow_T = o->snow_temperature;
    a.o_air_T = o->air_temperature; a.o_wind = o->wind_speed; a.o_rh = o->relative_humidity; a.o_qc = o->convective_heat_flux; a.o_ustar = o->friction_velocity;
    a.L_air = e->L_air.data(); a.n_snow = e->model.n_snow; a.s_begin = 0; a.s_end = e->S;
    const bool want_air = a.o_air_T || a.o_wind || a.o_rh || a.o_qc || a.o_ustar;
    int nbad = 0;
    /* same kernel selection as launch_range(): the block-cooperative path (here a "block" of one thread) for NoSnow + FixedIterations */
    const char* kenv = getenv("MC_KERNEL");
    const bool legacy = kenv && std::strcmp(kenv, "legacy") == 0, want_async = kenv && std::strcmp(kenv, "async") == 0;
    const bool coop = !legacy && e->model.n_snow == 0 && mcd::g_model_h.conv_kind == MC_CONV_FIXED;
    /* same kernel selection as launch_range(): the asynchronous kernel is opt-in (MC_KERNEL=async), for NoSnow models with the water-potential infiltration form */
    if (want_async && e->model.n_snow == 0 && !mcd::g_model_h.flux_alg) {
        const char* we = getenv("HE_WARPS"); const int nwarps = we ? atoi(we) : 3;
        const char* ge = getenv("HE_GRID"); const int grid = ge && atoi(ge) > 0 ? atoi(ge) : 2;
        if (e->N <= 10) run_async<10>(a, nwarps, grid, e->counters); else if (e->N <= 19) run_async<19>(a, nwarps, grid, e->counters); else run_async<MC_MAX_SOIL_NODES>(a, nwarps, grid, e->counters);
        for (int64_t s = 0; s < e->S; s++) {
            if (want_air) { mcd::Counters cn = {}; mcd::solve_air_site(a, s, cn); e->counters[3] += cn.n_obukhov; }
            if (o->status) o->status[s] = e->flags[s];
            if (e->flags[s]) nbad++;
        }
        e->days_done += ndays;
        return nbad;
    }
    std::vector<double> scr(mcd::Coop2Rows<MC_MAX_SOIL_NODES>::value);
    static mcd::MoistShared ms; std::memset(&ms, 0, sizeof(ms));
    static mcd::Coop2Shared sh2; std::memset(&sh2, 0, sizeof(sh2)); int rot_ode = 0, rot_moist = 0;
#ifdef MC_HOST_THREADS
    if (coop) {
        const char* tbe = getenv("HE_TB"); const int TB = tbe ? atoi(tbe) : 8;
        const char* ge = getenv("HE_GRID"); const int grid = ge ? atoi(ge) : 0;
        if (e->N <= 10) run_blocks_threaded<10>(a, TB, grid, e->counters); else if (e->N <= 19) run_blocks_threaded<19>(a, TB, grid, e->counters); else run_blocks_threaded<MC_MAX_SOIL_NODES>(a, TB, grid, e->counters);
        for (int64_t s = 0; s < e->S; s++) {
            if (want_air) { mcd::Counters cn = {}; mcd::solve_air_site(a, s, cn); e->counters[3] += cn.n_obukhov; }
            if (o->status) o->status[s] = e->flags[s];
            if (e->flags[s]) nbad++;
        }
        e->days_done += ndays;
        return nbad;
    }
#endif
    for (int64_t s = 0; s < e->S; s++) {
        mcd::Counters cn = {};
        if (coop) { if (e->N <= 10) mcd::solve_site_coop2<10>(a, s, cn, sh2, ms, scr.data(), rot_ode, rot_moist); else if (e->N <= 19) mcd::solve_site_coop2<19>(a, s, cn, sh2, ms, scr.data(), rot_ode, rot_moist); else mcd::solve_site_coop2<MC_MAX_SOIL_NODES>(a, s, cn, sh2, ms, scr.data(), rot_ode, rot_moist); }
        else if (e->model.n_snow == 0) { if (e->N <= 10) mcd::solve_site<10, 0>(a, s, cn); else if (e->N <= 19) mcd::solve_site<19, 0>(a, s, cn); else mcd::solve_site<MC_MAX_SOIL_NODES, 0>(a, s, cn); }
        else { if (e->N <= 10) mcd::solve_site<10, MC_SNOW_NODES>(a, s, cn); else if (e->N <= 19) mcd::solve_site<19, MC_SNOW_NODES>(a, s, cn); else mcd::solve_site<MC_MAX_SOIL_NODES, MC_SNOW_NODES>(a, s, cn); }
        if (want_air) mcd::solve_air_site(a, s, cn);
        e->counters[0] += cn.n_rhs; e->counters[1] += cn.n_accept; e->counters[2] += cn.n_reject; e->counters[3] += cn.n_obukhov;
        e->counters[4] += cn.n_newton; e->counters[5] += cn.n_infil; e->counters[6] += cn.n_day_passes;
        if (o->status) o->status[s] = e->flags[s];
        if (e->flags[s]) nbad++;
    }
    e->days_done += ndays;
    return nbad;
}

int he_get_counters(void* h, mc_counters* c)
{
    Emu* e = (Emu*)h; std::memset(c, 0, sizeof(*c));
    c->n_rhs = e->counters[0]; c->n_accept = e->counters[1]; c->n_reject = e->counters[2]; c->n_obukhov_iter = e->counters[3];
    c->n_newton_iter = e->counters[4]; c->n_infil_steps = e->counters[5]; c->n_day_passes = e->counters[6]; c->n_site_days = (uint64_t)(e->days_done * e->S);
    return 0;
}

int he_get_state(void* h, mc_state_soa* st)
{
    Emu* e = (Emu*)h; const size_t S = e->S, N = e->N;
    auto cp = [](double* dst, const std::vector<double>& v, size_t n) { if (dst) std::memcpy(dst, v.data(), n * sizeof(double)); };
    cp(st->soil_temperature, e->T, N * S); cp(st->soil_moisture, e->theta, N * S); cp(st->sum_phase, e->sum_phase, N * S); cp(st->fsal, e->fsal, N * S);
    cp(st->pool, e->pool, S); cp(st->soil_wetness, e->wet, S); cp(st->obukhov_ode, e->L_ode, S); cp(st->obukhov_profile, e->L_prof, S);
    cp(st->deep_temperature_prev, e->tdeep, S); cp(st->obukhov_air, e->L_air, S);
    if (st->days_done) st->days_done[0] = e->days_done;
    if (st->status) std::memcpy(st->status, e->flags.data(), S * sizeof(int32_t));
    return 0;
}

}

/* leaf functions for white-box comparison with the oracle */
extern "C" void he_soil_properties(const mc_model_desc* m, double P, double T, double theta, double bd, double md, double lam_m, double c_m, double* out3)
{
    mcd::build_dev_model(*m, mcd::g_model_h);
    mcd::soil_properties_full(T, P, theta, bd, md, lam_m, c_m, out3[0], out3[1], out3[2]);
}
extern "C" double he_goff_gratch(double T) { return mcd::goff_gratch(T); }
extern "C" double he_fastpow(double x, double y) { return mcd::fastpow(x, y); }

/* stomatal Newton (campbell.jl:270-284): the kernel's cycle-resolving form against the plain loop the reference runs */
extern "C" double he_stomatal_newton(const mc_model_desc* m, double psi_bar, double R_bar, double TP)
{
    mcd::build_dev_model(*m, mcd::g_model_h);
    return mcd::stomatal_newton(0.0, psi_bar, R_bar, TP);
}
extern "C" double he_stomatal_newton_plain(const mc_model_desc* m, double psi_bar, double R_bar, double TP, int* iterations)
{
    mcd::build_dev_model(*m, mcd::g_model_h);
    const mcd::DevModel& d = mcd::g_model_h;
    const double RL = d.RL, PC = d.PC, SP = d.SP;
    double psi_leaf = 0.0; int counter = 0;
    while (counter < d.moist_max_it) {
        if (psi_leaf > psi_bar) psi_leaf = psi_bar - TP * (R_bar + RL);
        const double xp = mcm::pow(psi_leaf / PC, SP);
        const double slope = TP * (R_bar + RL) * SP * xp / (psi_leaf * (1.0 + xp) * (1.0 + xp)) - 1.0;
        const double resid = psi_bar - psi_leaf - TP * (R_bar + RL) / (1.0 + xp);
        psi_leaf = psi_leaf - resid / slope;
        counter++;
        if (std::fabs(resid) <= 10.0) break;
    }
    if (iterations) *iterations = counter;
    return psi_leaf;
}

extern "C" void he_async_stats(unsigned long long* out) { for (int k = 0; k < 64; k++) { out[k] = g_async_stats[k]; g_async_stats[k] = 0; } }
extern "C" long long he_async_runs(void) { return g_async_runs; }      /* how many chunks went through the asynchronous kernel body */

extern "C" int he_set_state(void* h, const mc_state_soa* st)
{
    Emu* e = (Emu*)h; const size_t S = e->S, N = e->N;
    auto cp = [](std::vector<double>& v, const double* src, size_t n) { if (src) std::memcpy(v.data(), src, n * sizeof(double)); };
    cp(e->T, st->soil_temperature, N * S); cp(e->theta, st->soil_moisture, N * S); cp(e->sum_phase, st->sum_phase, N * S); cp(e->fsal, st->fsal, N * S);
    cp(e->pool, st->pool, S); cp(e->wet, st->soil_wetness, S); cp(e->L_ode, st->obukhov_ode, S); cp(e->L_prof, st->obukhov_profile, S);
    cp(e->tdeep, st->deep_temperature_prev, S); cp(e->L_air, st->obukhov_air, S);
    if (st->days_done) e->days_done = st->days_done[0];
    if (st->status) std::memcpy(e->flags.data(), st->status, S * sizeof(int32_t));
    return 0;
}
