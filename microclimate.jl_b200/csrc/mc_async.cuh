/*
 * mc_async.cuh - Asynchronous, shared-memory-resident form of solve_soil! (NoSnow, water-potential infiltration or prescribed moisture).
 * Part of mc_device.cuh (included from there, inside namespace mcd, after mc_site.cuh); not a stand-alone header.
 *
 * One persistent block per SM keeps SLOTS sites RESIDENT IN SHARED MEMORY: the integrator vectors k1..k7 / u / uprev, the clamped
 * stage temperatures and conductances, the hour's theta-only soil terms and (aliased onto the same rows, a site is never in both
 * phases) the Newton state of the implicit soil-water step live in a [array][node][slot] tile of 16 x NA x SLOTS doubles; the scalar
 * state of a site's state machine lives in a per-slot record next to it.  Nothing of the hot path is thread-local or in global scratch,
 * so a site's working set never leaves the SM between its load (once per launch) and its store.
 *
 * Sites advance INDEPENDENTLY (no hour lock-step, no block barrier): every site is a state machine whose node loops are published
 * as item passes and whose scalar control runs as single-lane tasks.  The 32 warps of the block are generic workers: a warp claims
 * up to 32 items of ONE task type from whatever slots offer them, executes them, and the lane that retires the last item of a site's
 * pass publishes the site's next request.  Task types:
 *
 *   channel 0, node items (claimed through a per-slot cursor word with compare-and-swap):
 *     COND   stage vector -> clamp -> Campbell-de Vries conductivity -> conductance        (N-1 items per RHS evaluation)
 *     P2     conduction du_i (node 0: surface balance total), error-estimate term          (N items)
 *     HOIST  theta-only soil terms of the coming ODE hour                                   (N items per hour)
 *     MU     soil-water Newton: water content / potential / humidity / conductivity of a row (N-1 items per iteration)
 *     MF     soil-water Newton: residual and tridiagonal row                                (N-1 items per iteration)
 *   channel 1, one lane per site (the mask bit IS the request):
 *     SURF   forcing at the stage time + surface energy balance of node 0 (runs beside COND)
 *     CTRL   Tsit5 controller: resumes the integrator's state machine, publishes the next stage
 *     THOMAS residual sum, Thomas elimination, back-substitution of one Newton iteration
 *     SUB    soil-water sub-step bookkeeping + stomatal Newton
 *     SITE   everything else: site load/store, day / pass / hour boundaries, outputs
 *
 * Arithmetic and operation order are those of solve_site_coop2 / coop2_* / soil_water_hour_coop (bit-identical to the oracle); only the
 * schedule differs.  Sites are handed to slots from a global counter, so a slow site (a Newton solve that runs to its iteration cap) delays
 * nobody but itself.
 */

#if defined(__CUDACC__)
#define MCA_LANES 32
MC_DEV int mca_cas(int* p, int cmp, int val) { return atomicCAS(p, cmp, val); }
MC_DEV int mca_add(int* p, int v) { return atomicAdd(p, v); }
MC_DEV unsigned mca_and32(unsigned* p, unsigned v) { return atomicAnd(p, v); }
MC_DEV unsigned mca_or32(unsigned* p, unsigned v) { return atomicOr(p, v); }
MC_DEV int mca_ld(const int* p) { return *(const volatile int*)p; }
MC_DEV unsigned mca_ldu(const unsigned* p) { return *(const volatile unsigned*)p; }
MC_DEV void mca_fence() { __threadfence_block(); }
MC_DEV void mca_idle(int ns) { __nanosleep(ns > 0 ? (unsigned)ns : 100u); }
MC_DEV unsigned long long mca_grab(unsigned long long* p) { return atomicAdd(p, 1ull); }
MC_DEV void mca_gadd(unsigned long long* p, unsigned long long v) { if (v) atomicAdd(p, v); }
MC_DEV int mca_popc(unsigned v) { return __popc(v); }
MC_DEV int mca_ctz(unsigned v) { return __ffs((int)v) - 1; }
#else
#ifndef MCA_LANES
#define MCA_LANES 4
#endif
MC_DEV int mca_cas(int* p, int cmp, int val) { __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); return cmp; }
MC_DEV int mca_add(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
MC_DEV unsigned mca_and32(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_ACQ_REL); }
MC_DEV unsigned mca_or32(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_ACQ_REL); }
MC_DEV int mca_ld(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
MC_DEV unsigned mca_ldu(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
MC_DEV void mca_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
MC_DEV void mca_idle(int) { std::this_thread::yield(); }
MC_DEV unsigned long long mca_grab(unsigned long long* p) { return __atomic_fetch_add(p, 1ull, __ATOMIC_ACQ_REL); }
MC_DEV void mca_gadd(unsigned long long* p, unsigned long long v) { if (v) __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
MC_DEV int mca_popc(unsigned v) { return __builtin_popcount(v); }
MC_DEV int mca_ctz(unsigned v) { return __builtin_ctz(v); }
#endif

/* rows of the shared-memory tile (per slot: 16 arrays x NA nodes) */
enum { A_K1A = 0, A_K2, A_K3, A_K4, A_K5, A_K6, A_K1B, A_UA, A_UB, A_T, A_LAM, A_Q, A_TH, A_PG, A_FW, A_CAP, A_ARRAYS };
/* soil-water phase: the Newton state is aliased onto rows the ODE does not need between two hours (k1 / u keep their rows).
 * Row r of these arrays is the reference's profile row i = r + 2 (water content of soil node r, parameters of soil node r + 1). */
enum { W_KVC = A_K2, W_RTS = A_K3, W_PSI = A_K4, W_K = A_K5, W_H = A_K6, W_WCN = A_T, W_SUB = A_LAM, W_DIA = A_Q, W_SUP = A_PG, W_F = A_FW, W_UPT = A_CAP };
/* per-block global scratch (L2-resident): site constants and the once-per-hour / once-per-day copies */
enum { G_PM = 0, G_LM, G_CPM, G_BD, G_PSIE, G_SAT, G_BB, G_KS, G_NEXP, G_ROOTR, G_ROOTZ, G_SUMPH, G_TBEF, G_TDAY, G_SPDAY, G_TPREV, G_ARRAYS };

enum { T0_COND = 0, T0_P2, T0_HOIST, T0_MU, T0_MF, T0_TYPES };
enum { T1_SURF = 0, T1_CTRL, T1_THOMAS, T1_SUB, T1_SITE, T1_TYPES };
enum { PC_LOAD = 0, PC_INIT_A, PC_INIT_B, PC_DAY_START, PC_PASS_START, PC_HOUR_START, PC_ODE_BEGIN, PC_HOUR_POST, PC_HOUR_END, PC_STORE, PC_DEAD };

/* work census of a launch (optional): per task type {claims, items, clock cycles spent executing}, then {idle polls, idle cycles} */
constexpr int MCA_STAT_TYPES = T0_TYPES + T1_TYPES;
constexpr int MCA_STATS = 3 * MCA_STAT_TYPES + 2;
#if defined(__CUDACC__)
MC_DEV long long mca_clock() { return clock64(); }
MC_DEV void mca_stat_add(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
#else
MC_DEV long long mca_clock() { return 0; }
MC_DEV void mca_stat_add(unsigned long long* p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
#endif

struct SlotState {
    /* Tsit5 controller (Own2) */
    double t, dt, dtpropose, qold, q11, tstop, target, w0, w1, w2, st_t, st_dt, req_t;
    /* published with a pass-1 request */
    double P, p0p, pp0, dtpub;
    /* surface-balance partial sums (SURF -> P2 item 0) */
    double Qa, Qc, Qe;
    /* forcing knots of the current hour, staged once per hour: {P, T, wind, zen, solar, cloud, rh, szen, fdiff} x {knot hr, knot hr+1} */
    double fk[9][2];
    /* SoilEnergyInputs scalars (HourPar) */
    double shade, eps_surf, eps_cloud, soil_wetness, albedo, L_prev, z0, lz, slope, svf;
    /* the environment instant the moisture step / day end consume */
    double e_lai, e_Tref, e_P, e_rh, e_wind, e_zen, e_rain_daily, e_Tdeep;
    /* carried scalars */
    double pool, wet_dyn, L_ode, L_prof, tdeep_stale, T_out_bottom, sky_T, site_albedo;
    /* soil-water owner */
    double EP, TP, rh_loc, K2, K3, TR, evaporation, surface_water_flux, sat0, half, EPot, psi1, rts1, psi_leaf, RLTR;
    int32_t s;
    int32_t st, guard, iter, iflags, stage, code, clampflag;
    int32_t flags, pc, jl, pass_iter, niter, hr, midnight_i;
    int32_t k, counter, mflags;
    uint32_t cn[7], cn_s0, cn_s3;
    unsigned char accept, u_modified, need_initdt, u_stale, kswap, uswap, is_last, have_infil;
    MC_MEMBER int k1arr() const { return kswap ? A_K1B : A_K1A; }
    MC_MEMBER int k7arr() const { return kswap ? A_K1A : A_K1B; }
    MC_MEMBER int uarr() const { return uswap ? A_UB : A_UA; }
    MC_MEMBER int uparr() const { return uswap ? A_UA : A_UB; }
};

template <int NA, int SLOTS>
struct AsyncBlock {
    static constexpr int SP = SLOTS | 1;                                   /* odd row pitch: the nodes of one site fall into different banks */
    static constexpr int ZD = (int)((sizeof(SlotState) + 7) / 8) | 1;      /* slot-record pitch in doubles, odd for the same reason */
    double arr[A_ARRAYS * NA * SP];
    double zraw[ZD * SLOTS];
    unsigned mask0[T0_TYPES][2], mask1[T1_TYPES][2];                       /* request bits, slot o = word o >> 5, bit o & 31 */
    int req0[SLOTS];                                                       /* channel-0 request of a slot: type << 16 | items claimed so far */
    int pend[SLOTS];                                                       /* items of the slot's current pass not yet retired */
    int alive;                                                             /* slots that may still receive a site */
    int ser_min;
    int abort;                                                             /* watchdog: a warp saw no work for far longer than any launch lasts */
    int affinity;
    int nreq;                                                              /* requests with unclaimed items: what an idle warp polls (one word) before it scans the masks */
    MC_MEMBER SlotState& z(int o) { return *reinterpret_cast<SlotState*>(zraw + (size_t)ZD * o); }
};

#define MCA_SA(arr_, node_) sa[((arr_) * NA + (node_)) * SP + o]
#define MCA_GA(arr_, node_) ga[((size_t)((arr_) * NA + (node_))) * SLOTS + o]

/* items per site of a channel-0 pass */
MC_DEV int async_items(int type, int N) { return (type == T0_P2 || type == T0_HOIST) ? N : N - 1; }

template <int NA, int SLOTS>
MC_DEV void async_enqueue1(AsyncBlock<NA, SLOTS>& B, int type, int o)
{
    mca_fence();
    mca_or32(&B.mask1[type][o >> 5], 1u << (o & 31));
    mca_add(&B.nreq, 1);
}
template <int NA, int SLOTS>
MC_DEV void async_request0(AsyncBlock<NA, SLOTS>& B, int type, int o, int extra_pending)
{
    const int N = MC_CONST_MODEL.N;
    B.req0[o] = type << 16;
    B.pend[o] = async_items(type, N) + extra_pending;
    mca_fence();
    mca_or32(&B.mask0[type][o >> 5], 1u << (o & 31));
    mca_add(&B.nreq, 1);
}

/* ---- forcing: staged knots of the current hour, global fallback for a stage time that sits exactly on the next knot ---- */
MC_DEV void async_stage_knots(SlotState& Z, const KArgs& a, int jl, int hr)
{
    const int64_t S = a.S;
    const int h1 = hr + 1 > 23 ? 23 : hr + 1;
    const int64_t o0 = ((int64_t)jl * 24 + hr) * S + Z.s, o1 = ((int64_t)jl * 24 + h1) * S + Z.s;
    const double* tab[9] = {a.f_P, a.f_T, a.f_wind, a.f_zen, a.f_solar, a.f_cloud, a.f_rh, a.f_szen, a.f_diffuse};
#if defined(__CUDACC__)
    /* cp.async: the 18 doubles go global -> shared without passing through registers; the wait comes after the issue of all of them */
    _Pragma("unroll") for (int v = 0; v < 9; v++) {
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(&Z.fk[v][0]), d1 = (unsigned)__cvta_generic_to_shared(&Z.fk[v][1]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0), "l"(tab[v] + o0) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d1), "l"(tab[v] + o1) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#else
    for (int v = 0; v < 9; v++) { Z.fk[v][0] = tab[v][o0]; Z.fk[v][1] = tab[v][o1]; }
#endif
}
MC_DEV double async_lerp(const SlotState& Z, const KArgs& a, int v, const double* base, int ik, double dk)
{
    if (ik == Z.hr) return (1.0 - dk) * Z.fk[v][0] + dk * Z.fk[v][1];
    const int i1 = ik + 1 > 23 ? 23 : ik + 1;
    const int64_t S = a.S; const double* b = base + (int64_t)Z.jl * 24 * S + Z.s;
    return (1.0 - dk) * MC_LDG(b + (int64_t)ik * S) + dk * MC_LDG(b + (int64_t)i1 * S);
}

/* stage vector of node i for the published stage: uprev + dt * sum(a_sj k_j) (tsit5_step) */
template <int NA, int SP>
MC_DEV double async_stage_x(const double* sa, int o, int i, int stage, int K1, int UP, double dt)
{
    using namespace tsit5;
    const double up = MCA_SA(UP, i);
    switch (stage) {
    case ST_INIT_F1: return up + dt * MCA_SA(A_K2, i);
    case ST_2: return up + dt * a21 * MCA_SA(K1, i);
    case ST_3: return up + dt * (a31 * MCA_SA(K1, i) + a32 * MCA_SA(A_K2, i));
    case ST_4: return up + dt * (a41 * MCA_SA(K1, i) + a42 * MCA_SA(A_K2, i) + a43 * MCA_SA(A_K3, i));
    case ST_5: return up + dt * (a51 * MCA_SA(K1, i) + a52 * MCA_SA(A_K2, i) + a53 * MCA_SA(A_K3, i) + a54 * MCA_SA(A_K4, i));
    case ST_6: return up + dt * (a61 * MCA_SA(K1, i) + a62 * MCA_SA(A_K2, i) + a63 * MCA_SA(A_K3, i) + a64 * MCA_SA(A_K4, i) + a65 * MCA_SA(A_K5, i));
    case ST_7: return up + dt * (a71 * MCA_SA(K1, i) + a72 * MCA_SA(A_K2, i) + a73 * MCA_SA(A_K3, i) + a74 * MCA_SA(A_K4, i) + a75 * MCA_SA(A_K5, i) + a76 * MCA_SA(A_K6, i));
    default: return up;                            /* ST_INIT_F0, ST_REFRESH: the RHS at the current state (u == uprev there) */
    }
}
MC_DEV int async_out_array(int stage, bool kswap)
{
    switch (stage) {
    case ST_INIT_F0: case ST_2: return A_K2;
    case ST_INIT_F1: case ST_3: return A_K3;
    case ST_4: return A_K4;
    case ST_5: return A_K5;
    case ST_6: return A_K6;
    case ST_7: return kswap ? A_K1A : A_K1B;
    default: return kswap ? A_K1B : A_K1A;         /* ST_REFRESH -> k1 */
    }
}

/* ---- channel-0 items ---- */
/* COND, node i (0..N-2): stage vector, clamp, conductance of the layer below node i; the item of node N-2 also stages the bottom node */
template <int NA, int SLOTS>
MC_DEV void async_item_cond(AsyncBlock<NA, SLOTS>& B, const double* ga, int o, int i)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL;
    double* sa = B.arr; SlotState& Z = B.z(o);
    const int code = Z.code, stage = code & 15;
    const bool kswap = (code >> 4) & 1, uswap = (code >> 5) & 1;
    const int K1 = kswap ? A_K1B : A_K1A, U = uswap ? A_UB : A_UA, UP = uswap ? A_UA : A_UB;
    const double dt = Z.dtpub;
    const int last = (i == m.N - 2) ? m.N - 1 : i;
    MC_ROLLED for (int n = i; n <= last; n++) {
        const double x = async_stage_x<NA, SP>(sa, o, n, stage, K1, UP, dt);
        if (stage == ST_7) MCA_SA(U, n) = x;
        const double Tc = jclamp(x, T_LO, m.T_hi);
        if (stage == ST_7 && Tc != x) mca_add(&Z.clampflag, 1);
        MCA_SA(A_T, n) = Tc;
        if (n < m.N - 1)
            MCA_SA(A_LAM, n) = soil_conductivity(Tc, Z.P, Z.p0p, Z.pp0, MCA_SA(A_TH, n), (MCA_GA(G_PM, n)), MCA_SA(A_PG, n), MCA_SA(A_FW, n),
                                                 (MCA_GA(G_LM, n))) / m.dzg[n];
    }
}

/* P2, node i (0..N-1): conduction (heat_transport_1d.jl:192-210), node 0 closes the surface balance; stage 7 adds the error-estimate term */
template <int NA, int SLOTS>
MC_DEV void async_item_p2(AsyncBlock<NA, SLOTS>& B, int o, int i)
{
    using namespace tsit5;
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL;
    double* sa = B.arr; const SlotState& Z = B.z(o);
    const int code = Z.code, stage = code & 15;
    const bool kswap = (code >> 4) & 1, uswap = (code >> 5) & 1;
    double du = 0.0;
    if (i == 0) {
        const double Q_cond = MCA_SA(A_LAM, 0) * (MCA_SA(A_T, 1) - MCA_SA(A_T, 0));
        du = 60.0 * (Z.Qa + Q_cond + Z.Qc + 0.0 - Z.Qe) / MCA_SA(A_CAP, 0);
    } else if (i < m.N - 1) {
        const double T_up = MCA_SA(A_T, i - 1), T_i = MCA_SA(A_T, i), T_dn = MCA_SA(A_T, i + 1);
        const double G_up = MCA_SA(A_LAM, i - 1), G_i = MCA_SA(A_LAM, i);
        du = 60.0 * (G_up * (T_up - T_i) + G_i * (T_dn - T_i)) / MCA_SA(A_CAP, i);
    }
    MCA_SA(async_out_array(stage, kswap), i) = du;
    if (stage == ST_7) {
        const int K1 = kswap ? A_K1B : A_K1A, U = uswap ? A_UB : A_UA, UP = uswap ? A_UA : A_UB;
        const double dt = Z.dtpub;
        const double ut = dt * (bt1 * MCA_SA(K1, i) + bt2 * MCA_SA(A_K2, i) + bt3 * MCA_SA(A_K3, i) + bt4 * MCA_SA(A_K4, i) + bt5 * MCA_SA(A_K5, i)
                                + bt6 * MCA_SA(A_K6, i) + bt7 * du);
        MCA_SA(A_Q, i) = ut / (m.abstol + fmax(fabs(MCA_SA(UP, i)), fabs(MCA_SA(U, i))) * m.reltol);
    }
}

/* HOIST, node i: theta-only soil property terms of the coming ODE hour (hoist_soil) */
template <int NA, int SLOTS>
MC_DEV void async_item_hoist(AsyncBlock<NA, SLOTS>& B, const double* ga, int o, int i)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL;
    double* sa = B.arr;
    const double th = MCA_SA(A_TH, i), pm = (MCA_GA(G_PM, i));
    const double cp = (MCA_GA(G_CPM, i)) + th * WATER_CP;
    const double rho = th * WATER_RHO + (MCA_GA(G_BD, i));
    MCA_SA(A_CAP, i) = (rho * cp) * m.dzc[i] / 2.0;
    MCA_SA(A_PG, i) = jmax(0.0, 1.0 - th - pm);
    MCA_SA(A_FW, i) = 1.0 / (1.0 + mcm::pow(th / m.theta0, -m.recirc_pow));
}

/* MU, profile row i = r + 2 (2..L): new water content / potential / humidity / conductivity (moist_item_update) */
template <int NA, int SLOTS>
MC_DEV void async_item_mu(AsyncBlock<NA, SLOTS>& B, const double* ga, int o, int r)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL;
    double* sa = B.arr; const SlotState& Z = B.z(o);
    const int fl = Z.mflags, n = r + 1;                                     /* parameters of soil node i - 1 = r + 1 */
    const double psi_e = (MCA_GA(G_PSIE, n)), sat = (MCA_GA(G_SAT, n)), bb = (MCA_GA(G_BB, n));
    if (fl & MF_S0) {                                                       /* first pass of the hour: the hour's temperature-dependent constants of this row */
        const int U = Z.uarr();
        MCA_SA(W_KVC, r) = 0.66 * DV_INFIL * vapour_density_sat(MCA_SA(U, r + 1));      /* B.Ts[i] = T0[i-1] */
        MCA_SA(W_RTS, r) = m.R * MCA_SA(U, r);                                           /* R * B.Ts[i-1] */
    }
    double wcn;
    if (fl & MF_S0) wcn = MCA_SA(A_TH, r);
    else wcn = jmax(sat * mcm::pow(psi_e / MCA_SA(W_PSI, r), 1.0 / bb), 1e-7);
    const double psi = psi_e * mcm::pow(sat / wcn, bb);
    MCA_SA(W_WCN, r) = wcn; MCA_SA(W_PSI, r) = psi;
    MCA_SA(W_H, r) = mcm::exp(MW_INFIL * psi / MCA_SA(W_RTS, r));
    const double K = (MCA_GA(G_KS, n)) * mcm::pow(psi_e / psi, (MCA_GA(G_NEXP, n)));
    MCA_SA(W_K, r) = K;
    if (fl & MF_FINAL) MCA_SA(A_TH, r) = wcn;
    if (fl & (MF_FINAL | MF_S0)) {          /* this state starts the next sub-step: root + soil resistance terms of its stomatal balance (campbell.jl:253-267) */
        const double rt = (MCA_GA(G_ROOTZ, n)) / K + (MCA_GA(G_ROOTR, n));
        MCA_SA(W_SUP, r) = rt; MCA_SA(W_SUB, r) = psi / rt; MCA_SA(W_DIA, r) = 1.0 / rt;       /* RT, Q1, Q2 live in rows the first MF pass overwrites */
    }
}

/* MF, profile row i = r + 2: residual and tridiagonal row (moist_item_assemble); root uptake is formed in a sub-step's first iteration */
template <int NA, int SLOTS>
MC_DEV void async_item_mf(AsyncBlock<NA, SLOTS>& B, const double* ga, int o, int r)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL;
    double* sa = B.arr; const SlotState& Z = B.z(o);
    const int L = m.N, i = r + 2, n = r + 1;
    const double g = m.g, dt = m.moist_dt;
    const bool first = (Z.mflags & MF_FIRST) != 0;
    const double psic = MCA_SA(W_PSI, r), Kc = MCA_SA(W_K, r), wcnc = MCA_SA(W_WCN, r), hc = MCA_SA(W_H, r);
    const double nexp = (MCA_GA(G_NEXP, n)), sat = (MCA_GA(G_SAT, n));
    double psim, Km, jvm, djvm;
    if (i == 2) {
        const double EP = Z.EP, rh_loc = Z.rh_loc;
        psim = first ? psic : Z.psi1; Km = 0.0;
        jvm = EP * (hc - rh_loc) / (1.0 - rh_loc);
        djvm = EP * MW_INFIL * hc / (Z.rts1 * (1.0 - rh_loc));
    } else {
        psim = MCA_SA(W_PSI, r - 1); Km = MCA_SA(W_K, r - 1);
        const double wcnm = MCA_SA(W_WCN, r - 1), hm1 = MCA_SA(W_H, r - 1);
        const double kvm = MCA_SA(W_KVC, r - 1) * ((MCA_GA(G_SAT, n - 1)) - (wcnm + wcnc) / 2.0) / m.mdz[i - 1];
        jvm = kvm * (hc - hm1);
        djvm = MW_INFIL * hm1 * kvm / MCA_SA(W_RTS, r - 1);
    }
    double psip, Kp, wcnp, hp;
    if (i + 1 <= L) { psip = MCA_SA(W_PSI, r + 1); Kp = MCA_SA(W_K, r + 1); wcnp = MCA_SA(W_WCN, r + 1); hp = MCA_SA(W_H, r + 1); }
    else { psip = (MCA_GA(G_PSIE, n)); Kp = (MCA_GA(G_KS, n)); wcnp = sat; hp = first ? 1.0 : hc; }     /* row L+1: psi_e[L], ks[L], sat[L], h = 1 then h[L] */
    const double kv = MCA_SA(W_KVC, r) * (sat - (wcnc + wcnp) / 2.0) / m.mdz[i];
    const double jvi = kv * (hp - hc);
    const double djvi = MW_INFIL * hc * kv / MCA_SA(W_RTS, r);
    const double cap = -1.0 * m.mmass[i] * wcnc / ((MCA_GA(G_BB, n)) * psic * dt);
    const double sub = -1.0 * Km / m.mdz[i - 1] + g * nexp * Km / psim;
    const double sup = -1.0 * Kp / m.mdz[i];
    const double dia = Kc / m.mdz[i - 1] + Kc / m.mdz[i] + cap - g * nexp * Kc / psic + djvm + djvi;
    double uptake;
    if (first) { uptake = (psic - Z.psi_leaf - Z.RLTR) / MCA_SA(W_SUP, r); MCA_SA(W_UPT, r) = uptake; }      /* W_SUP still holds RT of this row */
    else uptake = MCA_SA(W_UPT, r);
    const double F = ((psic * Kc - psim * Km) / m.mdz[i - 1] - (psip * Kp - psic * Kc) / m.mdz[i]) / (1.0 - nexp)
                   + m.mmass[i] * (wcnc - MCA_SA(A_TH, r)) / dt - g * (Km - Kc) + jvm - jvi + uptake;
    MCA_SA(W_SUB, r) = sub; MCA_SA(W_DIA, r) = dia; MCA_SA(W_SUP, r) = sup; MCA_SA(W_F, r) = F;
}

/* ---- channel-1 tasks ---- */
/* SURF: forcing at the stage time (interpolate_forcings) + surface energy balance of node 0 up to the conduction term (coop2_pre + coop2_surface) */
template <int NA, int SLOTS>
MC_DEV_NOINLINE void async_task_surf(AsyncBlock<NA, SLOTS>& B, const KArgs& a, int o)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL;
    const double* sa = B.arr; SlotState& Z = B.z(o);
    const double xk = Z.req_t / 60.0;
    int ik = (int)floor(xk);
    ik = ik < 0 ? 0 : (ik > 23 ? 23 : ik);
    const double dk = xk - (double)ik;
    const double P = async_lerp(Z, a, 0, a.f_P, ik, dk);
    const double T_air = async_lerp(Z, a, 1, a.f_T, ik, dk);
    const double wind = jmax(async_lerp(Z, a, 2, a.f_wind, ik, dk), 0.1);
    const double zen = jmin(90.0, nearbyint(async_lerp(Z, a, 3, a.f_zen, ik, dk) * 1000.0) / 1000.0);
    const double solar = jmax(0.0, async_lerp(Z, a, 4, a.f_solar, ik, dk));
    const double cloud = jclamp(async_lerp(Z, a, 5, a.f_cloud, ik, dk), 0.0, 1.0);
    const double rh = jclamp(async_lerp(Z, a, 6, a.f_rh, ik, dk), 0.0, 1.0);
    const int code = Z.code, stage = code & 15;
    const bool kswap = (code >> 4) & 1, uswap = (code >> 5) & 1;
    const double Ts_node = jclamp(async_stage_x<NA, SP>(sa, o, 0, stage, kswap ? A_K1B : A_K1A, uswap ? A_UA : A_UB, Z.dtpub), T_LO, m.T_hi);
    double Q_solar = (1.0 - Z.albedo) * solar * (1.0 - Z.shade);
    if (Z.slope > 0.0 && zen < 90.0) {
        const double szen = jmin(90.0, async_lerp(Z, a, 7, a.f_szen, ik, dk));
        const double fd = jclamp(async_lerp(Z, a, 8, a.f_diffuse, ik, dk), 0.0, 1.0);
        const double cz = mcm::cos(zen * PI_D / 180.0), csz = mcm::cos(szen * PI_D / 180.0);
        Q_solar = jmax(0.0, (((1.0 - fd) * Q_solar) / cz) * csz) + (fd * Q_solar) * Z.svf;
    }
    const WetAir wa = wet_air(T_air, rh, P);
    const double atm = atmospheric_radiation(wa.vapour_pressure, T_air);
    const double cloud_rad = m.sigma * Z.eps_cloud * mcm::pow4(T_air - 2.0);
    const double hill_rad = m.sigma * Z.eps_cloud * mcm::pow4(T_air);
    const double lw_sky = (atm * (1.0 - cloud) + cloud_rad * cloud) * (1.0 - Z.shade);
    const double incoming = (lw_sky + Z.shade * hill_rad) * Z.svf + hill_rad * (1.0 - Z.svf);
    const double Q_ir = incoming - ((1.0 - Z.shade) * m.sigma * Z.eps_surf) * mcm::pow4(Ts_node) - Z.shade * hill_rad;
    double Ts = Ts_node;
    if (fabs(Ts - T0C) > 100.0) Ts = T_air + 1.0;
    if (Ts == T_air) Ts = Ts + 0.1;
    const double dT = T_air - Ts;
    const double rho_cp = calc_rho_cp((Ts + T_air) / 2.0);
    double Qc;
    if (T_air >= Ts || zen >= 90.0) {
        const double ustar = jmax(m.kappa * wind / Z.lz, 1.0e-6);
        Qc = convective_flux(rho_cp, dT, ustar, 0.64 / Z.lz, sublayer_stanton(Z.z0, ustar));
    } else {
        const Obukhov ob = calc_obukhov(wind, Z.z0, m.zref, Z.lz, dT, rho_cp, Z.L_prev >= 0.0 ? -0.3 : Z.L_prev);
        Z.cn[3] += ob.iters;
        Qc = ob.Q;
    }
    const double hc = dT == 0.0 ? 0.5 : jmax(fabs(Qc / dT), 0.5);
    const double hd = (hc / (wa.specific_heat * wa.density)) * m.schmidt_factor;
    const double Tsc = Ts < T_LO ? T_LO : Ts;
    Z.Qe = (Z.soil_wetness * hd * (vapour_density_sat(Tsc) - wa.vapour_density)) * latent_heat(Tsc);
    Z.Qa = Q_solar + Q_ir; Z.Qc = Qc;
    Z.cn[0]++;
}

#define MCA_YIELD(label, stg, tt) do { I.st = (label); I.stage = (stg); I.req_t = (tt); return true; case (label):; } while (0)

/* Resumable adaptive integration to Z.target (coop2_advance on the slot's rows): true = an RHS evaluation is wanted at (stage, req_t) */
template <int NA, int SLOTS>
MC_DEV_NOINLINE bool async_ode_advance(AsyncBlock<NA, SLOTS>& B, int o)
{
    using namespace tsit5;
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL; const int N = m.N;
    double* sa = B.arr; SlotState& I = B.z(o);
    switch (I.st) {
    case 0:
        if (I.need_initdt) {                                             /* Hairer-Wanner initial step: two RHS evaluations */
            I.need_initdt = 0;
            { double s0 = 0.0; const int UP = I.uparr();
              MC_ROLLED for (int i = 0; i < N; i++) { const double ui = MCA_SA(UP, i); const double sk = m.abstol + fabs(ui) * m.reltol; const double q = ui / sk; s0 += q * q; }
              I.w0 = sqrt(s0 / N); }
            MCA_YIELD(1, ST_INIT_F0, I.t);
            { bool finite = true;
              MC_ROLLED for (int i = 0; i < N; i++) if (!isfinite(MCA_SA(A_K2, i))) finite = false;
              if (!finite) { I.dt = 1e-12; goto initdt_done; } }
            { double s1 = 0.0; const int UP = I.uparr();
              MC_ROLLED for (int i = 0; i < N; i++) { const double sk = m.abstol + fabs(MCA_SA(UP, i)) * m.reltol; const double q = MCA_SA(A_K2, i) / sk; s1 += q * q; }
              I.w1 = sqrt(s1 / N); }
            I.w2 = (I.w0 < 1e-5 || I.w1 < 1e-5) ? 1e-6 : (I.w0 / I.w1) / 100.0;
            I.w2 = fmin(I.w2, 1440.0);
            I.dtpub = I.w2;
            MCA_YIELD(2, ST_INIT_F1, I.t + I.w2);
            { double s2 = 0.0; const int UP = I.uparr();
              MC_ROLLED for (int i = 0; i < N; i++) { const double sk = m.abstol + fabs(MCA_SA(UP, i)) * m.reltol; const double q = (MCA_SA(A_K3, i) - MCA_SA(A_K2, i)) / sk; s2 += q * q; }
              const double d2 = sqrt(s2 / N) / I.w2;
              const double md = fmax(I.w1, d2);
              double dt1;
              if (md <= 1e-15) dt1 = fmax(1e-6, I.w2 * 1e-3);
              else dt1 = mcm::exp10(-(2.0 + mcm::log10(md)) / 5.0);
              I.dt = fmin(fmin(100.0 * I.w2, dt1), 1440.0); }
        initdt_done:;
        }
        while (I.t < I.target - 1e-10) {                                 /* step_to_hour! */
            I.guard = 0;
            for (;;) {                                                   /* one accepted step */
                if (I.iter > 0) {
                    if (!I.accept) {
                        I.dt = I.dt / fmin(5.0, I.q11 / 0.9);
                    } else {
                        I.uswap = !I.uswap; I.u_stale = 1;               /* uprev <- u */
                        I.dt = I.dtpropose;
                        if (I.u_modified) { MCA_YIELD(3, ST_REFRESH, I.t); }
                        else I.kswap = !I.kswap;                          /* k1 <- k7 (FSAL) */
                    }
                } else if (I.u_modified) {
                    const int U = I.uarr(), UP = I.uparr();
                    MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(UP, i) = MCA_SA(U, i);
                }
                I.iter++;
                I.dt = fmin(1440.0, I.dt);
                I.dt = fmin(I.dt, I.tstop - I.t);
                if (!(I.dt > 1e-12) || ++I.guard > 100000) {
                    I.iflags |= (I.dt != I.dt) ? MC_STATUS_NAN : (I.guard > 100000 ? MC_STATUS_STEP_LIMIT : MC_STATUS_DT_UNDERFLOW);
                    I.t = I.tstop; I.accept = 1; I.dtpropose = 1.0;
                    const int K1 = I.k1arr(), K7 = I.k7arr(), U = I.uarr(), UP = I.uparr();
                    MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(K7, i) = MCA_SA(K1, i);
                    if (I.u_stale) { MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(U, i) = MCA_SA(UP, i); I.u_stale = 0; }
                    I.clampflag = 1;                                     /* force the exact check below */
                    break;
                }
                I.st_t = I.t; I.st_dt = I.dt;
                I.dtpub = I.st_dt;
                MCA_YIELD(4, ST_2, I.st_t + c1 * I.st_dt);
                MCA_YIELD(5, ST_3, I.st_t + c2 * I.st_dt);
                MCA_YIELD(6, ST_4, I.st_t + c3 * I.st_dt);
                MCA_YIELD(7, ST_5, I.st_t + c4 * I.st_dt);
                MCA_YIELD(8, ST_6, I.st_t + I.st_dt);
                I.clampflag = 0;
                MCA_YIELD(9, ST_7, I.st_t + I.st_dt);
                {
                    double ss = 0.0;
                    MC_ROLLED for (int i = 0; i < N; i++) { const double q = MCA_SA(A_Q, i); ss += q * q; }      /* in order, node 0 first */
                    const double EEst = sqrt(ss / N);
                    I.u_modified = 0;
                    double q;
                    if (EEst == 0.0) q = 0.1;
                    else {
                        I.q11 = fastpow(EEst, 7.0 / 50.0);
                        q = I.q11 / fastpow(I.qold, 2.0 / 25.0);
                        q = fmax(0.1, fmin(5.0, q / 0.9));
                    }
                    I.accept = EEst <= 1.0;
                    if (EEst != EEst) { I.iflags |= MC_STATUS_NAN; I.accept = 1; q = 1.0; }
                    if (I.accept) {
                        I.cn[1]++;
                        I.qold = fmax(EEst, 1e-4);
                        const double dtnew = I.dt / q;
                        double ttmp = I.t + I.dt;
                        if (fabs(ttmp - I.tstop) < 100.0 * 2.220446049250313e-16 * fmax(I.t, I.tstop)) ttmp = I.tstop;
                        I.t = ttmp;
                        I.dtpropose = fmin(1440.0, dtnew);
                        I.u_stale = 0;
                        break;
                    }
                    I.cn[2]++;
                    I.u_stale = 1;                                        /* u <- uprev: the array of u now holds the rejected proposal */
                }
            }
            if (I.clampflag) {   /* re-clamp the state after every internal step (simulation.jl:891-900); the COND items flag candidates */
                const int U = I.uarr(), UP = I.uparr();
                bool changed = false;
                MC_ROLLED for (int k = 0; k < N; k++) { const double v = MCA_SA(U, k); const double c = jclamp(v, T_LO, m.T_hi); if (c != v) { MCA_SA(U, k) = c; changed = true; } }
                if (changed) { MC_ROLLED for (int k = 0; k < N; k++) MCA_SA(UP, k) = MCA_SA(U, k); I.u_modified = 1; }
            }
        }
        I.st = 0; I.stage = ST_NONE;
        return false;
    case 10:                                                             /* one-shot RHS (prototype evaluation of init()) */
        I.st = 0; I.stage = ST_NONE;
        return false;
    }
    return false;
}

/* publish the RHS evaluation the integrator asked for: pressure terms for the COND items, then COND (N-1 items) + SURF */
template <int NA, int SLOTS>
MC_DEV void async_publish_rhs(AsyncBlock<NA, SLOTS>& B, const KArgs& a, int o)
{
    const DevModel& m = MC_CONST_MODEL;
    SlotState& Z = B.z(o);
    const double xk = Z.req_t / 60.0;
    int ik = (int)floor(xk);
    ik = ik < 0 ? 0 : (ik > 23 ? 23 : ik);
    const double dk = xk - (double)ik;
    const double P = async_lerp(Z, a, 0, a.f_P, ik, dk);
    Z.P = P; Z.p0p = m.P0 / P; Z.pp0 = P / m.P0;
    Z.code = Z.stage | (Z.kswap ? 16 : 0) | (Z.uswap ? 32 : 0);
    async_request0(B, T0_COND, o, 1);
    mca_or32(&B.mask1[T1_SURF][o >> 5], 1u << (o & 31));
    mca_add(&B.nreq, 1);
}

/* CTRL: an RHS evaluation is complete -> resume the integrator */
template <int NA, int SLOTS>
MC_DEV void async_task_ctrl(AsyncBlock<NA, SLOTS>& B, const KArgs& a, int o)
{
    if (async_ode_advance(B, o)) async_publish_rhs(B, a, o);
    else async_enqueue1(B, T1_SITE, o);
}

/* THOMAS: the serial part of one Newton iteration (moist_thomas): residual sum, elimination, back-substitution with the saturation clamps */
template <int NA, int SLOTS>
MC_DEV_NOINLINE void async_task_thomas(AsyncBlock<NA, SLOTS>& B, const double* ga, int o)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL; const int L = m.N;
    double* sa = B.arr; SlotState& W = B.z(o);
    W.counter++; W.cn[4]++;
    W.K2 = MCA_SA(W_K, 0);
    W.K3 = 3 <= L ? MCA_SA(W_K, 1) : (MCA_GA(G_KS, L - 1));
    double err = 0.0, ccm = 0.0, rrm = 0.0, diaL = 0.0, FL = 0.0;
    MC_ROLLED for (int i = 2; i <= L; i++) {
        const int r = i - 2;
        const double sub = MCA_SA(W_SUB, r), sup = MCA_SA(W_SUP, r);
        double dia = MCA_SA(W_DIA, r);
        double F = MCA_SA(W_F, r);
        err += fabs(F);
        if (i >= 3) { dia -= sub * ccm; F -= sub * rrm; }
        if (i <= L - 1) {
            double cc = sup / dia;
            if (fabs(cc) < 1e-8) cc = 0.0;
            const double rr = F / dia;
            MCA_SA(W_SUP, r) = cc; MCA_SA(W_F, r) = rr; ccm = cc; rrm = rr;       /* in place: cc over sup, rr over F */
        } else { diaL = dia; FL = F; }
    }
    double dpsi = FL / diaL;
    MC_ROLLED for (int i = L; i >= 2; i--) {
        const int r = i - 2;
        double psi = MCA_SA(W_PSI, r);
        const double psi_e = (MCA_GA(G_PSIE, r + 1));
        if (i == L) psi = jmin(psi - dpsi, psi_e);
        else {
            dpsi = MCA_SA(W_F, r) - MCA_SA(W_SUP, r) * dpsi;
            psi -= dpsi;
            if (psi > psi_e) psi = (psi + dpsi + psi_e) / 2.0;
        }
        MCA_SA(W_PSI, r) = psi;                                               /* raw potential: the MU pass turns it into the new state */
    }
    const bool converged = err <= m.moist_tol;
    const bool last = converged || W.counter >= m.moist_max_it;
    if (last && !converged) W.flags |= MC_STATUS_NEWTON_MAXITER;
    W.mflags = last ? MF_FINAL : 0;
    async_request0(B, T0_MU, o, 0);
}

/* SUB: end of sub-step k (after its final MU pass) and start of the next one (root/soil resistances, stomatal Newton) */
template <int NA, int SLOTS>
MC_DEV_NOINLINE void async_task_sub(AsyncBlock<NA, SLOTS>& B, const double* ga, int o)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL; const int L = m.N;
    double* sa = B.arr; SlotState& W = B.z(o);
    if (W.counter > 0) {                                              /* infiltration_step!'s outputs + pool */
        const double psi2 = MCA_SA(W_PSI, 0), psi3 = 3 <= L ? MCA_SA(W_PSI, 1) : (MCA_GA(G_PSIE, L - 1));
        W.surface_water_flux = ((psi2 * W.K2 - psi3 * W.K3) / ((1.0 - (MCA_GA(G_NEXP, 1))) * m.mdz[2]) + m.g * W.K2 + W.TR) * m.moist_dt;
        W.evaporation = W.EP * (MCA_SA(W_H, 0) - W.rh_loc) / (1.0 - W.rh_loc) * m.moist_dt;
        W.pool = jclamp(W.pool - jmax(0.0, W.surface_water_flux) - jmax(0.0, W.evaporation), 0.0, m.max_pool);
        if (W.pool > 0.0) {                                           /* wet_surface_node on theta[0] = row 2 */
            double th0 = MCA_SA(A_TH, 0);
            const double before = th0;
            W.pool = wet_surface_node(&th0, W.pool, W.sat0, W.half);
            if (th0 != before) MCA_SA(A_TH, 0) = th0;
            if (th0 != before && W.k + 1 < m.niter_moist) {           /* the next sub-step's start recomputes the state of row 2 from the wetted water content */
                const double psi_e = (MCA_GA(G_PSIE, 1));
                const double psi = psi_e * mcm::pow((MCA_GA(G_SAT, 1)) / th0, (MCA_GA(G_BB, 1)));
                MCA_SA(W_WCN, 0) = th0; MCA_SA(W_PSI, 0) = psi;
                MCA_SA(W_H, 0) = mcm::exp(MW_INFIL * psi / MCA_SA(W_RTS, 0));
                const double K = (MCA_GA(G_KS, 1)) * mcm::pow(psi_e / psi, (MCA_GA(G_NEXP, 1)));
                MCA_SA(W_K, 0) = K;
                const double rt = (MCA_GA(G_ROOTZ, 1)) / K + (MCA_GA(G_ROOTR, 1));
                MCA_SA(W_SUP, 0) = rt; MCA_SA(W_SUB, 0) = psi / rt; MCA_SA(W_DIA, 0) = 1.0 / rt;
            }
        }
        W.k++;
    }
    if (W.k < m.niter_moist) {                                        /* moist_substep_start */
        W.cn[5]++;
        double psum = 0.0, rsum = 0.0;
        MC_ROLLED for (int r = 0; r <= L - 2; r++) { psum += MCA_SA(W_SUB, r); rsum += MCA_SA(W_DIA, r); }      /* in-order sums of psi/rt and 1/rt */
        W.psi1 = MCA_SA(W_PSI, 0);                                    /* water_potential[1] = water_potential[2], never updated again */
        const double psi_bar = psum / rsum, R_bar = 1.0 / rsum;
        const double psi_leaf = stomatal_newton(0.0, psi_bar, R_bar, W.TP);       /* campbell.jl:270-284 */
        W.TR = W.TP / (1.0 + mcm::pow(psi_leaf / m.PC, m.SP));
        W.psi_leaf = psi_leaf; W.RLTR = m.RL * W.TR;
        W.counter = 0;
        W.mflags = MF_FIRST;
        async_request0(B, T0_MF, o, 0);
    } else {
        async_enqueue1(B, T1_SITE, o);
    }
}

/* SITE: load / day / pass / hour boundaries of solve_soil! (simulation.jl:355-819), the serial glue of solve_site_coop2 as a resumable state machine */
template <int NA, int SLOTS>
MC_DEV_NOINLINE void async_task_site(AsyncBlock<NA, SLOTS>& B, const KArgs& a, double* ga, int o)
{
    constexpr int SP = AsyncBlock<NA, SLOTS>::SP;
    const DevModel& m = MC_CONST_MODEL;
    const int N = m.N;
    const int nrows = a.ndays * 24;
    const bool dynamic = m.dynamic != 0, nonconsec = m.nonconsec != 0;
    const int64_t S = a.S;
    double* sa = B.arr; SlotState& Z = B.z(o);
    SiteRef r; r.a = &a; r.s = Z.s;
    const bool multi = nonconsec ? (m.iters_per_day > 1 || m.conv_kind != MC_CONV_FIXED) : (m.spinup != 0);       /* a day may take several passes */

    auto tins_mean = [&](int jl) { double ssum = 0.0; MC_ROLLED for (int h = 0; h < 24; h++) ssum += r.hourly(a.f_T, jl * 24 + h); ssum += r.hourly(a.f_T, jl * 24 + 23); return ssum / 25.0; };
    auto set_env = [&](const EnvInstant& e) {
        Z.e_lai = e.lai; Z.e_Tref = e.T_ref; Z.e_P = e.P; Z.e_rh = e.rh; Z.e_wind = e.wind; Z.e_zen = e.zenith; Z.e_rain_daily = e.rain_daily; Z.e_Tdeep = e.T_deep;
    };
    auto set_hour_params = [&](const EnvInstant& e) {
        Z.shade = e.shade; Z.eps_surf = e.eps_surf; Z.eps_cloud = e.eps_cloud;
        Z.albedo = e.albedo; Z.soil_wetness = dynamic ? Z.wet_dyn : e.wet_daily;
        Z.sky_T = sky_temperature(Z.svf, e.P, e.rh, e.T_ref, e.eps_cloud, e.cloud, e.shade, a.has_lw != 0, e.lw);
    };
    auto put_potential = [&](int row) {
        if (!a.o_psi) return;
        MC_ROLLED for (int i = 0; i < N; i++)
            r.put_node(a.o_psi, row, i, r.node(a.psi_e, i) * mcm::pow((1.0 - r.node(a.bd, i) / r.node(a.md, i)) / MCA_SA(A_TH, i), r.node(a.bb, i)));
    };
    auto put_humidity = [&](int U) {                                          /* initialise_soil_humidity!, dynamic.jl:32-35: always row 1 of the buffer */
        if (!m.dynamic || !a.o_h) return;
        MC_ROLLED for (int i = 0; i < N; i++) {
            const double psi = r.node(a.psi_e, i) * mcm::pow((1.0 - r.node(a.bd, i) / r.node(a.md, i)) / MCA_SA(A_TH, i), r.node(a.bb, i));
            r.put_node(a.o_h, 0, i, jclamp(mcm::exp(MW_INFIL * psi / (m.R * MCA_SA(U, i))), 0.0, 1.0));
        }
    };
    auto put_props = [&](int row, int U, double P) {
        if (!a.o_lam && !a.o_cp && !a.o_rho) return;
        MC_ROLLED for (int i = 0; i < N; i++) {
            double lam, cp, rho;
            soil_properties_full(MCA_SA(U, i), P, MCA_SA(A_TH, i), r.node(a.bd, i), r.node(a.md, i), r.node(a.lam_m, i), r.node(a.c_m, i), lam, cp, rho);
            r.put_node(a.o_lam, row, i, lam); r.put_node(a.o_cp, row, i, cp); r.put_node(a.o_rho, row, i, rho);
        }
    };

    for (;;) {
        switch (Z.pc) {
        case PC_LOAD: {
            const unsigned long long idx = mca_grab(a.next_site);
            if ((int64_t)idx >= a.s_end - a.s_begin) { Z.pc = PC_DEAD; mca_add(&B.alive, -1); return; }
            const int64_t s = a.s_begin + (int64_t)idx;
            Z.s = (int32_t)s; r.s = s;
            Z.kswap = 0; Z.uswap = 0;
            MC_ROLLED for (int i = 0; i < N; i++) {
                MCA_SA(A_UA, i) = a.T[(int64_t)i * S + s]; MCA_SA(A_TH, i) = a.theta[(int64_t)i * S + s]; MCA_GA(G_SUMPH, i) = a.sum_phase[(int64_t)i * S + s];
                const double f = a.fsal[(int64_t)i * S + s]; MCA_SA(A_K1A, i) = f; MCA_SA(A_K1B, i) = f;
                /* site constants of node i, once per residency (the per-site kernels rebuild them every hour) */
                const double bd = r.node(a.bd, i), md = r.node(a.md, i), b = r.node(a.bb, i);
                const double nexp = 2.0 + 3.0 / b;
                MCA_GA(G_PM, i) = bd / md; MCA_GA(G_LM, i) = r.node(a.lam_m, i); MCA_GA(G_CPM, i) = bd / md * r.node(a.c_m, i); MCA_GA(G_BD, i) = bd;
                MCA_GA(G_PSIE, i) = -fabs(r.node(a.psi_e, i)); MCA_GA(G_SAT, i) = 1.0 - bd / md; MCA_GA(G_BB, i) = b; MCA_GA(G_KS, i) = r.node(a.ks, i);
                MCA_GA(G_NEXP, i) = nexp;
                if (i >= 1) {                                                   /* profile row i + 1: thickness depth[i+2] - depth[i] of the reference's 1-based grid */
                    const int row = i + 1;
                    const double rd = r.node(a.rootd, i);
                    const double thick = (row == N ? 1e20 : m.depths[row - 1]) - (row == 2 ? -1e10 : (row == 3 ? 0.0 : m.depths[row - 3]));
                    if (rd > 0.0) {
                        MCA_GA(G_ROOTR, i) = m.root_res / (rd * thick / 2.0);
                        MCA_GA(G_ROOTZ, i) = (1.0 - nexp) * mcm::log(PI_D * m.root_radius * m.root_radius * rd) / (4.0 * PI_D * rd * thick / 2.0);
                    } else { MCA_GA(G_ROOTR, i) = 1e20; MCA_GA(G_ROOTZ, i) = 0.0; }
                }
            }
            Z.pool = a.pool[s]; Z.wet_dyn = a.wetness[s]; Z.L_ode = a.L_ode[s]; Z.L_prof = a.L_prof[s]; Z.tdeep_stale = a.tdeep_prev[s];
            Z.flags = a.flags[s];
            Z.site_albedo = MC_LDG(a.albedo + s);
            Z.z0 = MC_LDG(a.z0 + s); Z.slope = MC_LDG(a.slope + s); Z.svf = MC_LDG(a.svf + s); Z.lz = mcm::log(m.zref / Z.z0);
            Z.iflags = 0; Z.iter = 0; Z.accept = 0; Z.u_modified = 0; Z.t = 0; Z.dt = 0; Z.dtpropose = 0; Z.qold = 1e-4; Z.q11 = 1; Z.tstop = 1440.0;
            Z.st = 0; Z.guard = 0; Z.need_initdt = 0; Z.u_stale = 0; Z.stage = ST_NONE; Z.clampflag = 0;
            Z.sky_T = 0.0; Z.T_out_bottom = 0.0; Z.jl = 0; Z.hr = 0; Z.is_last = 0; Z.have_infil = 0;
            for (int k = 0; k < 7; k++) Z.cn[k] = 0;
            if (a.day0 == 0) {
                /* init(): SciMLBase.init evaluates the RHS once at prototype parameters to seed fsalfirst (simulation.jl:196-238,853-860) */
                async_stage_knots(Z, a, 0, 0);
                const EnvInstant env = get_instant(r, 0, 0, Z.site_albedo);
                set_env(env);
                Z.shade = env.shade; Z.eps_surf = env.eps_surf; Z.eps_cloud = env.eps_cloud; Z.soil_wetness = 0.0; Z.albedo = Z.site_albedo; Z.L_prev = Z.L_ode;
                MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(A_UB, i) = a.has_init_T ? MCA_SA(A_UA, i) : 280.0;      /* uparr of uswap = false */
                Z.pc = PC_INIT_A;
                async_request0(B, T0_HOIST, o, 0);
                return;
            }
            Z.pc = PC_DAY_START;
            continue;
        }
        case PC_INIT_A:
            Z.cn_s0 = Z.cn[0]; Z.cn_s3 = Z.cn[3];                                /* the prototype evaluation is not counted */
            Z.st = 10; Z.stage = ST_REFRESH; Z.req_t = 0.0; Z.dtpub = 0.0;
            Z.pc = PC_INIT_B;
            async_publish_rhs(B, a, o);
            return;
        case PC_INIT_B: {
            /* solve_soil! pre-loop (simulation.jl:395-452) */
            Z.cn[0] = Z.cn_s0; Z.cn[3] = Z.cn_s3;
            const int U = Z.uarr();
            if (!a.has_init_T) { const double t_init = tins_mean(0); MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(U, i) = t_init; }
            MC_ROLLED for (int i = 0; i < N; i++) MCA_GA(G_SUMPH, i) = 0.0;
            put_props(0, U, 101325.0);
            put_potential(0);
            MC_ROLLED for (int i = 0; i < N; i++) { r.put_node(a.o_T, 0, i, MCA_SA(U, i)); r.put_node(a.o_theta, 0, i, MCA_SA(A_TH, i)); }
            put_humidity(U);
            Z.tdeep_stale = Z.e_Tdeep;
            Z.pool = 0.0;
            Z.pc = PC_DAY_START;
            continue;
        }
        case PC_DAY_START: {
            const int jl = Z.jl; const int64_t j = a.day0 + jl;
            const int U = Z.uarr();
            Z.niter = nonconsec ? m.iters_per_day : ((m.spinup && j == 0) ? m.conv_iters : 1);
            int noon_i = 0; double best = r.hourly(a.f_zen, jl * 24);
            MC_ROLLED for (int i = 1; i < 24; i++) { const double zz = r.hourly(a.f_zen, jl * 24 + i); if (zz < best) { best = zz; noon_i = i; } }
            Z.midnight_i = (noon_i + 12) % 24;
            if (nonconsec) {                                                      /* per-day resets (simulation.jl:480-516) */
                MC_ROLLED for (int i = 0; i < N; i++) MCA_GA(G_SUMPH, i) = 0.0;
                if (!dynamic && a.has_presc) { MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(A_TH, i) = MC_LDG(a.presc + ((int64_t)jl * N + i) * S + Z.s); }
                else { MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(A_TH, i) = r.node(a.init_theta, i); }
                const double t_reset = tins_mean(jl);
                if (!a.has_init_T) { MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(U, i) = t_reset; }
                else { MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(U, i) = r.node(a.init_T, i); }
            }
            MCA_SA(U, N - 1) = Z.tdeep_stale;
            if (multi) { MC_ROLLED for (int i = 0; i < N; i++) { const double v = MCA_SA(U, i); MCA_GA(G_TPREV, i) = v; MCA_GA(G_TDAY, i) = v; MCA_GA(G_SPDAY, i) = MCA_GA(G_SUMPH, i); } }
            Z.pass_iter = 0; Z.is_last = 0;
            Z.pc = PC_PASS_START;
            continue;
        }
        case PC_PASS_START: {                                                     /* outer fixed point over day passes (simulation.jl:530) */
            const int jl = Z.jl; const int U = Z.uarr(), UP = Z.uparr();
            Z.pass_iter++; Z.cn[6]++;
            {   /* is_converged */
                const int iter = Z.pass_iter, niter = Z.niter;
                bool last;
                if (m.conv_kind == MC_CONV_FIXED) last = iter >= niter;
                else if (iter >= niter) last = true;
                else if (iter <= 1) last = niter <= 1;
                else { double mx = 0.0; MC_ROLLED for (int i = 0; i < N; i++) mx = fmax(mx, fabs(MCA_SA(U, i) - MCA_GA(G_TPREV, i))); last = mx < m.conv_tol; }
                Z.is_last = last;
            }
            if (multi) {
                MC_ROLLED for (int i = 0; i < N; i++) MCA_GA(G_TPREV, i) = MCA_SA(U, i);
                if (!nonconsec) { MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(U, i) = MCA_GA(G_TDAY, i); }
                MC_ROLLED for (int i = 0; i < N; i++) MCA_GA(G_SUMPH, i) = MCA_GA(G_SPDAY, i);
            }
            const int day_row = jl * 24;
            if (Z.is_last) {                                                      /* day row 1 = state at the start of the output pass (:557-592) */
                MC_ROLLED for (int i = 0; i < N; i++) { r.put_node(a.o_T, day_row, i, MCA_SA(U, i)); r.put_node(a.o_theta, day_row, i, MCA_SA(A_TH, i)); }
                r.put_scalar(a.o_pool, day_row, Z.pool);
                put_potential(day_row);
                put_humidity(U);
                const EnvInstant e0 = get_instant(r, jl, day_row, Z.site_albedo);
                r.put_scalar(a.o_sky, day_row, sky_temperature(Z.svf, e0.P, e0.rh, e0.T_ref, e0.eps_cloud, e0.cloud, e0.shade, a.has_lw != 0, e0.lw));
                put_props(day_row, U, e0.P);
            }
            const EnvInstant env = get_instant(r, jl, day_row, Z.site_albedo);
            set_env(env);
            MCA_SA(U, N - 1) = env.T_deep;
            set_hour_params(env);
            /* reinit!(integrator, u0; t0=0, reset_dt=true, reinit_cache=false): k1 (fsalfirst) deliberately stays stale */
            MC_ROLLED for (int i = 0; i < N; i++) MCA_SA(UP, i) = MCA_SA(U, i);
            Z.u_stale = 0;
            Z.t = 0.0; Z.iter = 0; Z.u_modified = 0; Z.accept = 0; Z.qold = 1e-4; Z.q11 = 1.0; Z.tstop = 1440.0;
            Z.need_initdt = 1;                                                    /* the two init-dt RHS calls join hour 0's evaluations */
            Z.hr = 0;
            Z.pc = PC_HOUR_START;
            continue;
        }
        case PC_HOUR_START: {
            const int U = Z.uarr();
            if (Z.is_last) { MC_ROLLED for (int k = 0; k < N; k++) MCA_GA(G_TBEF, k) = MCA_SA(U, k); }
            async_stage_knots(Z, a, Z.jl, Z.hr);
            Z.pc = PC_ODE_BEGIN;
            async_request0(B, T0_HOIST, o, 0);
            return;
        }
        case PC_ODE_BEGIN:
            Z.L_prev = Z.L_ode;
            /* step_to_hour! (simulation.jl:888-903) */
            Z.target = (Z.hr + 1) * 60.0; Z.tstop = Z.target; Z.st = 0;
            Z.pc = PC_HOUR_POST;
            if (async_ode_advance(B, o)) { async_publish_rhs(B, a, o); return; }
            continue;
        case PC_HOUR_POST: {
            const int jl = Z.jl, i = Z.hr, step = jl * 24 + i;
            const int U = Z.uarr();
            if (Z.is_last) {                                                      /* phase_transition!, latent_heat.jl:21-112 (sequential over nodes) */
                const double tol = 1.0e-4;
                MC_ROLLED for (int n = 0; n < N; n++) {
                    double mean, mean_past;
                    if (n < N - 1) { mean = 0.5 * (MCA_SA(U, n) + MCA_SA(U, n + 1)); mean_past = 0.5 * (MCA_GA(G_TBEF, n) + MCA_GA(G_TBEF, n + 1)); }
                    else { mean = MCA_SA(U, n); mean_past = MCA_GA(G_TBEF, n); }
                    const double th = MCA_SA(A_TH, n);
                    double mass = 0.0;
                    if (th > 0.0) mass = (n < N - 1 ? (m.depths[n + 1] - m.depths[n]) : ((m.depths[n] + 1.0) - m.depths[n])) * 1000.0 * th;
                    if (th <= 0.0 || mass <= 0.0) { MCA_GA(G_SUMPH, n) = 0.0; continue; }
                    const double max_latent = LATENT_FUSION * mass;
                    if ((mean_past - T0C > tol) && (mean - T0C <= -tol)) {
                        double sp = MCA_GA(G_SUMPH, n) + (mean_past - mean) * mass * WATER_CP;
                        if (sp >= max_latent) sp = max_latent;
                        MCA_GA(G_SUMPH, n) = sp;
                        MCA_SA(U, n) = T0C; if (n < N - 1) MCA_SA(U, n + 1) = T0C;
                    } else if ((mean_past - T0C < -tol) && (mean - T0C >= tol)) {
                        double sp = MCA_GA(G_SUMPH, n) - (mean - mean_past) * mass * WATER_CP;
                        if (sp <= 0.0) sp = 0.0;
                        else { MCA_SA(U, n) = T0C; if (n < N - 1) MCA_SA(U, n + 1) = T0C; }
                        MCA_GA(G_SUMPH, n) = sp;
                    }
                }
                Z.T_out_bottom = MCA_SA(U, N - 1);
            }
            if (i < 23) {
                const EnvInstant env = get_instant(r, jl, step + 1, Z.site_albedo);
                set_env(env);
                MCA_SA(U, N - 1) = env.T_deep;
                set_hour_params(env);
            }
            Z.u_modified = 1;
            {   /* init_surface_fluxes! (monin_obukhov.jl:247-264): hourly warm start of the Obukhov length */
                const double T_air = r.hourly(a.f_T, step);
                const double wind = jmax(r.hourly(a.f_wind, step), 0.1);
                const double zen = jmin(90.0, nearbyint(r.hourly(a.f_zen, step) * 1000.0) / 1000.0);
                const double Tsf = MCA_SA(U, 0);
                if (T_air < Tsf && zen < 90.0) {
                    const Obukhov ob = calc_obukhov(wind, Z.z0, m.zref, Z.lz, T_air - Tsf, calc_rho_cp((Tsf + T_air) / 2.0), Z.L_ode >= 0.0 ? -0.3 : Z.L_ode);
                    Z.cn[3] += ob.iters;
                    Z.L_ode = ob.L;
                }
            }
            Z.have_infil = 0;
            Z.pc = PC_HOUR_END;
            if (Z.is_last) {
                double rain;
                if (m.rain_hourly) rain = r.hourly(a.f_rain, step);
                else rain = (i == Z.midnight_i) ? Z.e_rain_daily : 0.0;
                Z.pool = jclamp(Z.pool + rain, 0.0, m.max_pool);
                if (dynamic) {
                    /* soil_water_balance! (campbell.jl:378-495): MOST profile -> potential evaporation -> pool wetting -> niter_moist implicit sub-steps */
                    Z.have_infil = 1;
                    const double T_surf = MCA_SA(U, 0), T_air = Z.e_Tref, P = Z.e_P, rh = Z.e_rh;
                    const double z = m.zref, hloc = m.zloc, z0 = Z.z0;
                    const double lz = mcm::log(z / z0);
                    const double dT = T_air - T_surf;
                    const double rho_cp = calc_rho_cp((T_surf + T_air) / 2.0);
                    double T_loc, Qc;
                    if (T_air >= T_surf || Z.e_zen >= 90.0) {
                        const double ustar = jmax(m.kappa * Z.e_wind / lz, 1.0e-6);
                        const double Stb = 0.64 / lz, Sts = sublayer_stanton(z0, ustar);
                        Qc = convective_flux(rho_cp, dT, ustar, Stb, Sts);
                        const double Tz0 = (T_air * Stb + T_surf * Sts) / (Stb + Sts);
                        if (hloc < z0) T_loc = T_surf + (Tz0 - T_surf) * (hloc / z0);
                        else if (hloc == z) T_loc = T_air;
                        else T_loc = Tz0 + (T_air - Tz0) * mcm::log(hloc / z0) / lz;
                    } else {
                        const Obukhov ob = calc_obukhov(Z.e_wind, z0, z, lz, dT, rho_cp, Z.L_prof);
                        Z.cn[3] += ob.iters;
                        Z.L_prof = ob.L; Qc = ob.Q;
                        const double Tz0 = (T_air * ob.Stb + T_surf * ob.Sts) / (ob.Stb + ob.Sts);
                        if (hloc < z0) T_loc = T_surf + (Tz0 - T_surf) * (hloc / z0);
                        else if (hloc == z) T_loc = T_air;
                        else {
                            const double x = calc_phi_m(hloc, m.gamma, ob.L);
                            const double tl = jmax(mcm::log(hloc / z0) - calc_psi_h(x), 1e-6);
                            T_loc = Tz0 + (T_air - Tz0) * tl / (lz - ob.psi_h);
                        }
                    }
                    /* local relative humidity and potential evaporation (campbell.jl:416-440) */
                    const double es_air = goff_gratch(T_air);
                    const double e_ref = es_air * rh;
                    const double rh_ref_prof = jclamp(e_ref / es_air, 0.0, 1.0);
                    const double rh_loc = jclamp((es_air * rh_ref_prof) / goff_gratch(T_loc), 0.0, 0.99);
                    const double hc = jmax(fabs(Qc / (T_surf - T_air)), 0.5);
                    const WetAir wa = wet_air(T_air, rh, P);
                    const double hd = (hc / (wa.specific_heat * wa.density)) * m.schmidt_factor;
                    const double Tsc = T_surf < T_LO ? T_LO : T_surf;
                    const double Qe = (hd * (vapour_density_sat(Tsc) - wa.vapour_density)) * latent_heat(Tsc);      /* saturated=true */
                    const double EPot = jmax(1e-7, Qe / latent_heat(T_surf));
                    Z.sat0 = (MCA_GA(G_SAT, 0));
                    Z.half = (m.depths[1] - m.depths[0]) / 2.0;
                    {   double th0 = MCA_SA(A_TH, 0); const double before = th0;
                        Z.pool = wet_surface_node(&th0, Z.pool, Z.sat0, Z.half);
                        if (th0 != before) MCA_SA(A_TH, 0) = th0; }
                    Z.EPot = EPot;
                    Z.EP = mcm::exp(-0.82 * Z.e_lai) * EPot; Z.TP = EPot - Z.EP; Z.rh_loc = rh_loc;
                    Z.rts1 = m.R * MCA_SA(U, 0);
                    Z.k = 0; Z.counter = 0; Z.evaporation = 0.0; Z.surface_water_flux = 0.0; Z.TR = 0.0; Z.K2 = 0.0; Z.K3 = 0.0;
                    if (m.niter_moist > 0) {
                        Z.mflags = MF_S0;
                        async_request0(B, T0_MU, o, 0);
                        return;
                    }
                }
            }
            continue;
        }
        case PC_HOUR_END: {
            const int jl = Z.jl, i = Z.hr, step = jl * 24 + i;
            const int U = Z.uarr();
            if (Z.have_infil) {
                if (m.niter_moist > 0) MCA_SA(A_TH, N - 1) = (MCA_GA(G_SAT, N - 1));      /* water_content[L+1] = saturation of the last row */
                Z.wet_dyn = jclamp(jmax(0.0, Z.evaporation) / (Z.EPot * m.moist_dt), 0.0, 1.0);
            }
            const int out_row = step + 1;
            if (Z.is_last && out_row < nrows && !(i == 23 && nonconsec)) {
                r.put_scalar(a.o_pool, out_row, Z.pool);
                MC_ROLLED for (int k = 0; k < N; k++) { r.put_node(a.o_T, out_row, k, k == N - 1 ? Z.T_out_bottom : MCA_SA(U, k)); r.put_node(a.o_theta, out_row, k, MCA_SA(A_TH, k)); }
                r.put_scalar(a.o_sky, out_row, Z.sky_T);
                if (Z.have_infil && (a.o_psi || a.o_h)) {
                    MC_ROLLED for (int k = 0; k < N - 1; k++) { r.put_node(a.o_psi, out_row, k, MCA_SA(W_PSI, k)); r.put_node(a.o_h, out_row, k, MCA_SA(W_H, k)); }
                    r.put_node(a.o_psi, out_row, N - 1, (MCA_GA(G_PSIE, N - 1))); r.put_node(a.o_h, out_row, N - 1, MCA_SA(W_H, N - 2));      /* row L+1: psi_e[L], h[L+1] = h[L] */
                }
                put_props(out_row, U, r.hourly(a.f_P, out_row));
            }
            Z.hr = i + 1;
            if (Z.hr < 24) { Z.pc = PC_HOUR_START; continue; }
            if (!Z.is_last) { Z.pc = PC_PASS_START; continue; }
            Z.tdeep_stale = Z.e_Tdeep;
            Z.jl = jl + 1;
            Z.pc = Z.jl < a.ndays ? PC_DAY_START : PC_STORE;
            continue;
        }
        case PC_STORE: {
            const int64_t s = Z.s;
            const int U = Z.uarr(), K1 = Z.k1arr();
            int flags = Z.flags;
            MC_ROLLED for (int i = 0; i < N; i++) {
                const double Tv = MCA_SA(U, i), th = MCA_SA(A_TH, i);
                a.T[(int64_t)i * S + s] = Tv; a.theta[(int64_t)i * S + s] = th; a.sum_phase[(int64_t)i * S + s] = MCA_GA(G_SUMPH, i);
                if (!(isfinite(Tv) && isfinite(th))) flags |= MC_STATUS_NAN;
                a.fsal[(int64_t)i * S + s] = MCA_SA(K1, i);
            }
            a.pool[s] = Z.pool; a.wetness[s] = Z.wet_dyn; a.L_ode[s] = Z.L_ode; a.L_prof[s] = Z.L_prof; a.tdeep_prev[s] = Z.tdeep_stale;
            a.flags[s] = flags | Z.iflags;
            if (a.counters) { for (int k = 0; k < 7; k++) mca_gadd(a.counters + k, (unsigned long long)Z.cn[k]); }
            Z.pc = PC_LOAD;
            continue;
        }
        default:
            return;
        }
    }
}

/* ---- the worker loop ---- */
struct AsyncClaim { int chan, type, total, nseg; unsigned seg[4]; unsigned bits[2]; };

/* Pick a task type and claim up to MCA_LANES items of it (executed by one lane per warp).  Node-item types are preferred while a full warp of
 * them is on offer; single-lane tasks are taken in batches where possible (they unlock the node items) and always when nothing else is. */
template <int NA, int SLOTS>
MC_DEV AsyncClaim async_try_claim(AsyncBlock<NA, SLOTS>& B, int salt, int warp)
{
    const int N = MC_CONST_MODEL.N;
    AsyncClaim c; c.chan = -1; c.type = 0; c.total = 0; c.nseg = 0; c.bits[0] = 0; c.bits[1] = 0;
    if (mca_ld(&B.nreq) <= 0) return c;                                    /* nothing on offer: one load instead of a scan of every mask */
    unsigned m0[T0_TYPES][2], m1[T1_TYPES][2];
    int best0 = -1, n0 = 0, best1 = -1, n1 = 0;
    for (int t = 0; t < T0_TYPES; t++) {
        m0[t][0] = mca_ldu(&B.mask0[t][0]); m0[t][1] = SLOTS > 32 ? mca_ldu(&B.mask0[t][1]) : 0u;
        const int cnt = (mca_popc(m0[t][0]) + mca_popc(m0[t][1])) * async_items(t, N);
        if (cnt > n0) { n0 = cnt; best0 = t; }
    }
    for (int t = 0; t < T1_TYPES; t++) {
        m1[t][0] = mca_ldu(&B.mask1[t][0]); m1[t][1] = SLOTS > 32 ? mca_ldu(&B.mask1[t][1]) : 0u;
        const int cnt = mca_popc(m1[t][0]) + mca_popc(m1[t][1]);
        if (cnt > n1) { n1 = cnt; best1 = t; }
    }
    if (best0 < 0 && best1 < 0) return c;
    bool take0 = best0 >= 0 && (n0 >= MCA_LANES || best1 < 0 || n1 < B.ser_min);
    if (B.affinity) {
        /* Warp w issues from scheduler w & 3.  Keeping the warps of one scheduler in one task class keeps that scheduler's instruction cache on one
         * code region: schedulers 0-1 take the ODE node items, scheduler 2 the soil-water node items, scheduler 3 the single-lane tasks; a warp whose
         * class has nothing on offer falls back to the general rule above. */
        const int cls = warp & 3;
        if (cls <= 1) {
            int bt = -1, bn = 0;
            for (int t = T0_COND; t <= T0_HOIST; t++) { const int cnt = (mca_popc(m0[t][0]) + mca_popc(m0[t][1])) * async_items(t, N); if (cnt > bn) { bn = cnt; bt = t; } }
            if (bt >= 0) { best0 = bt; take0 = true; }
        } else if (cls == 2) {
            int bt = -1, bn = 0;
            for (int t = T0_MU; t <= T0_MF; t++) { const int cnt = (mca_popc(m0[t][0]) + mca_popc(m0[t][1])) * async_items(t, N); if (cnt > bn) { bn = cnt; bt = t; } }
            if (bt >= 0) { best0 = bt; take0 = true; }
        } else if (best1 >= 0) take0 = false;
    }
    if (take0) {
        const int t = best0, n = async_items(t, N);
        int got = 0;
        for (int w0 = 0; w0 < 2 && got < MCA_LANES && c.nseg < 4; w0++) {
            const int w = (w0 + salt) & 1;
            unsigned bits = m0[t][w];
            while (bits && got < MCA_LANES && c.nseg < 4) {
                const int b = mca_ctz(bits); bits &= bits - 1;
                const int o = w * 32 + b;
                for (;;) {
                    const int rq = mca_ld(&B.req0[o]);
                    const int cur = rq & 0xffff;
                    if ((rq >> 16) != t || cur >= n) break;
                    const int take = (n - cur) < (MCA_LANES - got) ? (n - cur) : (MCA_LANES - got);
                    if (mca_cas(&B.req0[o], rq, rq + take) == rq) {
                        if (cur + take >= n) { mca_and32(&B.mask0[t][w], ~(1u << b)); mca_add(&B.nreq, -1); }      /* the last items are taken: withdraw the offer (before they run, so before any re-request) */
                        c.seg[c.nseg++] = (unsigned)o | ((unsigned)cur << 8) | ((unsigned)take << 16);
                        got += take;
                        break;
                    }
                }
            }
        }
        if (got) { c.chan = 0; c.type = t; c.total = got; mca_fence(); }
        return c;
    }
    {
        const int t = best1;
        int got = 0;
        for (int w = 0; w < 2 && got < MCA_LANES; w++) {
            unsigned want = m1[t][w];
            if (!want) continue;
            while (mca_popc(want) > MCA_LANES - got) want &= want - 1;       /* no more than the lanes left */
            const unsigned old = mca_and32(&B.mask1[t][w], ~want);
            c.bits[w] = old & want;
            got += mca_popc(c.bits[w]);
        }
        if (got) { c.chan = 1; c.type = t; c.total = got; mca_add(&B.nreq, -got); mca_fence(); }
        return c;
    }
}

/* a pass of slot o lost `cnt` items: the lane that retires the last one publishes what follows */
template <int NA, int SLOTS>
MC_DEV void async_retire(AsyncBlock<NA, SLOTS>& B, int type0_or_surf, int o, int cnt)
{
    mca_fence();
    if (mca_add(&B.pend[o], -cnt) != cnt) return;
    mca_fence();
    switch (type0_or_surf) {
    case T0_COND: async_request0(B, T0_P2, o, 0); break;                   /* also reached from SURF (type passed as T0_COND) */
    case T0_P2: async_enqueue1(B, T1_CTRL, o); break;
    case T0_HOIST: async_enqueue1(B, T1_SITE, o); break;
    case T0_MU: if (B.z(o).mflags & (MF_S0 | MF_FINAL)) async_enqueue1(B, T1_SUB, o); else async_request0(B, T0_MF, o, 0); break;
    default: async_enqueue1(B, T1_THOMAS, o); break;                       /* T0_MF */
    }
}

#if !defined(__CUDACC__)
MC_DEV int mca_nth_bit(unsigned lo, unsigned hi, int n)
{
    for (int b = 0; b < 64; b++) { const bool set = b < 32 ? ((lo >> b) & 1u) : ((hi >> (b - 32)) & 1u); if (set && n-- == 0) return b; }
    return -1;
}
#endif

#if defined(__CUDACC__)
#define MCA_IS_LEADER ((threadIdx.x & 31) == 0)
#else
#define MCA_IS_LEADER true
#endif

/* body of one worker warp; on the host (tests/hostemu) a "warp" is an OS thread that runs its MCA_LANES lanes one after the other */
template <int NA, int SLOTS>
MC_DEV void async_worker(AsyncBlock<NA, SLOTS>& B, const KArgs& a, double* ga, int warp)
{
    int salt = warp;
    unsigned long long wst[MCA_STATS];                                     /* this warp's census (leader lane) */
    for (int k = 0; k < MCA_STATS; k++) wst[k] = 0;
    long long t_last_work = mca_clock();
    int backoff = 0;                                                       /* idle sleep doubles from the base up to 16x while nothing turns up: idle warps must not eat the issue slots of working ones */
    for (;;) {
        AsyncClaim c;
#if defined(__CUDACC__)
        const int lane = threadIdx.x & 31;
        if (lane == 0) c = async_try_claim(B, salt, warp);
        c.chan = __shfl_sync(0xffffffffu, c.chan, 0); c.type = __shfl_sync(0xffffffffu, c.type, 0); c.total = __shfl_sync(0xffffffffu, c.total, 0);
        c.nseg = __shfl_sync(0xffffffffu, c.nseg, 0);
        c.seg[0] = __shfl_sync(0xffffffffu, c.seg[0], 0); c.seg[1] = __shfl_sync(0xffffffffu, c.seg[1], 0);
        c.seg[2] = __shfl_sync(0xffffffffu, c.seg[2], 0); c.seg[3] = __shfl_sync(0xffffffffu, c.seg[3], 0);
        c.bits[0] = __shfl_sync(0xffffffffu, c.bits[0], 0); c.bits[1] = __shfl_sync(0xffffffffu, c.bits[1], 0);
#else
        c = async_try_claim(B, salt, warp);
#endif
        salt++;
        if (c.chan < 0) {
            if (mca_ld(&B.alive) <= 0 || mca_ld(&B.abort)) break;
            const long long ti = mca_clock();
            mca_idle((a.async_idle_ns > 0 ? a.async_idle_ns : 250) << backoff);
            if (backoff < 4) backoff++;
            const long long tn = mca_clock();
            wst[3 * MCA_STAT_TYPES] += 1; wst[3 * MCA_STAT_TYPES + 1] += (unsigned long long)(tn - ti);
            if (tn - t_last_work > (1ll << 37)) { B.abort = 1; mca_fence(); }          /* ~70 s without work: give up rather than hang the device (the host reports it) */
            continue;
        }
        backoff = 0;
        const long long t_begin = mca_clock();
        if (c.chan == 0) {
#if defined(__CUDACC__)
            {
#else
            for (int lane = 0; lane < MCA_LANES; lane++) {
#endif
                if (lane < c.total) {
                    int rem = lane, o = 0, item = 0;
                    for (int j = 0; j < c.nseg; j++) {
                        const int cnt = (int)(c.seg[j] >> 16);
                        if (rem < cnt) { o = (int)(c.seg[j] & 0xff); item = (int)((c.seg[j] >> 8) & 0xff) + rem; break; }
                        rem -= cnt;
                    }
                    switch (c.type) {
                    case T0_COND: async_item_cond(B, ga, o, item); break;
                    case T0_P2: async_item_p2(B, o, item); break;
                    case T0_HOIST: async_item_hoist(B, ga, o, item); break;
                    case T0_MU: async_item_mu(B, ga, o, item); break;
                    default: async_item_mf(B, ga, o, item); break;
                    }
                }
            }
#if defined(__CUDACC__)
            __syncwarp();
            if (lane < c.nseg) async_retire(B, c.type, (int)(c.seg[lane] & 0xff), (int)(c.seg[lane] >> 16));
#else
            for (int j = 0; j < c.nseg; j++) async_retire(B, c.type, (int)(c.seg[j] & 0xff), (int)(c.seg[j] >> 16));
#endif
        } else {
#if defined(__CUDACC__)
            {
#else
            for (int lane = 0; lane < MCA_LANES; lane++) {
#endif
                if (lane < c.total) {
#if defined(__CUDACC__)
                    const int n_lo = __popc(c.bits[0]);
                    const int o = lane < n_lo ? (int)__fns(c.bits[0], 0, lane + 1) : 32 + (int)__fns(c.bits[1], 0, lane - n_lo + 1);
#else
                    const int o = mca_nth_bit(c.bits[0], c.bits[1], lane);
#endif
                    switch (c.type) {
                    case T1_SURF: async_task_surf(B, a, o); async_retire(B, T0_COND, o, 1); break;
                    case T1_CTRL: async_task_ctrl(B, a, o); break;
                    case T1_THOMAS: async_task_thomas(B, ga, o); break;
                    case T1_SUB: async_task_sub(B, ga, o); break;
                    default: async_task_site(B, a, ga, o); break;
                    }
                }
            }
        }
        {
            const int k = 3 * (c.chan == 0 ? c.type : T0_TYPES + c.type);
            t_last_work = mca_clock();
            wst[k] += 1; wst[k + 1] += (unsigned long long)c.total; wst[k + 2] += (unsigned long long)(t_last_work - t_begin);
        }
    }
    if (a.async_stats && MCA_IS_LEADER) { for (int k = 0; k < MCA_STATS; k++) if (wst[k]) mca_stat_add(a.async_stats + k, wst[k]); }
    if (MCA_IS_LEADER && warp == 0 && mca_ld(&B.abort) && a.flags) mca_stat_add(a.counters + 7, 1ull);      /* counters[7]: watchdog aborts */
}

/* block set-up: every slot starts by asking for a site */
template <int NA, int SLOTS>
MC_DEV void async_block_init(AsyncBlock<NA, SLOTS>& B, int tid, int nthreads, int active_slots, int ser_min = 0, int affinity = 0)
{
    const int n = active_slots <= 0 || active_slots > SLOTS ? SLOTS : active_slots;
    for (int o = tid; o < SLOTS; o += nthreads) { B.z(o).pc = o < n ? PC_LOAD : PC_DEAD; B.req0[o] = 0; B.pend[o] = 0; }
    if (tid == 0) {
        for (int t = 0; t < T0_TYPES; t++) { B.mask0[t][0] = 0; B.mask0[t][1] = 0; }
        for (int t = 0; t < T1_TYPES; t++) { B.mask1[t][0] = 0; B.mask1[t][1] = 0; }
        B.mask1[T1_SITE][0] = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
        B.mask1[T1_SITE][1] = n > 32 ? (n >= 64 ? 0xffffffffu : ((1u << (n - 32)) - 1u)) : 0u;
        B.alive = n;
        B.ser_min = ser_min > 0 ? ser_min : MCA_LANES / 4;
        B.abort = 0; B.affinity = affinity;
        B.nreq = n;                                                        /* the initial SITE requests */
    }
}
