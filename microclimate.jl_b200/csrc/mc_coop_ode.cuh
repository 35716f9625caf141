/*
 * mc_coop_ode.cuh - Block-cooperative ODE phase: resumable Tsit5 state machine, item passes over (site, node), per-block scratch.
 * Part of mc_device.cuh (included from there, inside namespace mcd, in this order); not a stand-alone header.
 */

/* ==========================================================================================================================
 * Block-cooperative ODE phase (NoSnow).  Every thread still OWNS one site, but the node loops of the ODE hour — above all the
 * Campbell-de Vries conductivity of every node at every RHS evaluation (3 Goff-Gratch evaluations + pow + de Vries weights,
 * ~1.5 k instructions) — are served by the WHOLE BLOCK over a flat list of (site, node) items of the sites that currently need an
 * RHS evaluation.  Sites that have already reached the hour boundary no longer idle their lanes: they serve items of the sites
 * still stepping, and every warp of the block shares one short instruction window.  The integrator is the SAME arithmetic as
 * tsit5_step / determine_initdt above, written as a resumable state machine (protothread style): each yield publishes an RHS request
 * and returns; the next call resumes right after it.
 * ========================================================================================================================== */
#ifndef MC_COOP_TB
#define MC_COOP_TB 1024     /* one block per SM: every warp of the SM is in the same phase (measured 128 / 256 / 512 / 1024 threads: 135 / 148 / 156 / 163 k site-days/s) */
#endif
constexpr int COOP_TB = MC_COOP_TB;

struct RhsCtx { double P, T_air, wind, zen, solar, cloud, rh, szen, fd; };

/* Block-level list building: every thread with `pred` appends its id to `list`, positions from `*counter` (shared memory).  On the GPU one
 * shared-memory atomic per warp (ballot + popc); on the host (tests/hostemu, one or several OS threads per "block") a plain atomic. */
MC_DEV void block_list_append(bool pred, int* list, int* counter, int tid)
{
#if defined(__CUDACC__)
    const unsigned bal = __ballot_sync(0xffffffffu, pred);
    const int lane = tid & 31;
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(counter, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pred) list[base + __popc(bal & ((1u << lane) - 1u))] = tid;
#else
    if (pred) list[__atomic_fetch_add(counter, 1, __ATOMIC_RELAXED)] = tid;
#endif
}
MC_DEV void block_count_add(bool pred, int* counter, int tid)
{
#if defined(__CUDACC__)
    const unsigned bal = __ballot_sync(0xffffffffu, pred);
    if ((tid & 31) == 0 && bal) atomicAdd(counter, __popc(bal));
#else
    (void)tid;
    if (pred) __atomic_fetch_add(counter, 1, __ATOMIC_RELAXED);
#endif
}
MC_DEV void block_flag_or(int* flag, int v)
{
#if defined(__CUDACC__)
    atomicOr(flag, v);
#else
    __atomic_fetch_or(flag, v, __ATOMIC_RELAXED);
#endif
}

/* ==========================================================================================================================
 * The integrator's vectors (k1..k7, u, uprev) live in the block's scratch instead of thread-local memory and EVERY node loop of
 * the ODE hour is an item pass over (requesting site, node):
 *
 *     owners : resume the Tsit5 state machine -> publish (stage, dt, pressure terms)                                   | barrier
 *     pass 1 : stage vector x = uprev + dt * sum(a_sj k_j)  ->  clamped T  ->  Campbell-de Vries conductivity          | barrier
 *     pass 2 : conduction du_i of nodes 1..N-1 (+ the error-estimate term of stage 7); owners: surface energy balance   | barrier
 *
 * Only scalar control (step-size controller, in-order error sum) and the surface energy balance remain per-site work.  The copies
 * uprev <- u and k1 <- k7 of an accepted step become role swaps of two pairs of physical arrays (uswap / kswap), done lazily at
 * the next loop header exactly where OrdinaryDiffEq copies.  Every floating-point operation and every summation order is that of
 * tsit5_step / determine_initdt / soil_rhs above.
 * ========================================================================================================================== */
enum { PK_A = 0, PK_2, PK_3, PK_4, PK_5, PK_6, PK_B, PU_A, PU_B, P_T, P_LAM, P_Q, P_TH, P_PM, P_PG, P_FW, P_LM, P_CAP, P_ARRAYS };
enum { ST_NONE = 0, ST_INIT_F0, ST_INIT_F1, ST_REFRESH, ST_2, ST_3, ST_4, ST_5, ST_6, ST_7 };

struct Coop2Shared {
    int list[COOP_TB]; int count[3]; int clampflag[COOP_TB];      /* count: three rotating list counters (see coop2_rounds) */
    double P[COOP_TB], p0p[COOP_TB], pp0[COOP_TB], dt[COOP_TB];
    unsigned char code[COOP_TB];        /* stage | kswap << 4 | uswap << 5 */
};

struct Own2 {                           /* what the owner keeps of its integrator */
    double t, dt, dtpropose, qold, q11, tstop, target, w0, w1, w2, st_t, st_dt, req_t, q0;
    int st, guard, iter, flags, stage, par;
    bool accept, u_modified, need_initdt, u_stale, kswap, uswap;
    MC_MEMBER int k1arr() const { return kswap ? PK_B : PK_A; }
    MC_MEMBER int k7arr() const { return kswap ? PK_A : PK_B; }
    MC_MEMBER int uarr() const { return uswap ? PU_B : PU_A; }
    MC_MEMBER int uparr() const { return uswap ? PU_A : PU_B; }
};

MC_DEV int stage_out_array(int stage, bool kswap)
{
    switch (stage) {
    case ST_INIT_F0: case ST_2: return PK_2;
    case ST_INIT_F1: case ST_3: return PK_3;
    case ST_4: return PK_4;
    case ST_5: return PK_5;
    case ST_6: return PK_6;
    case ST_7: return kswap ? PK_A : PK_B;
    default: return kswap ? PK_B : PK_A;          /* ST_REFRESH -> k1 */
    }
}

#define MC_SCR(arr, node, t) scr[((size_t)((arr) * NA + (node))) * MC_TB + (t)]

/* pass 1, one (site o, node i) item */
template <int NA>
MC_DEV void coop2_item_stage(const Coop2Shared& sh, Coop2Shared& shw, double* scr, int o, int i)
{
    using namespace tsit5;
    const DevModel& m = MC_CONST_MODEL;
    const int code = sh.code[o], stage = code & 15;
    const bool kswap = (code >> 4) & 1, uswap = (code >> 5) & 1;
    const int K1 = kswap ? PK_B : PK_A, U = uswap ? PU_B : PU_A, UP = uswap ? PU_A : PU_B;
    const double dt = sh.dt[o];
    const double up = MC_SCR(UP, i, o);
    double x;
    switch (stage) {
    case ST_INIT_F1: x = up + dt * MC_SCR(PK_2, i, o); break;
    case ST_2: x = up + dt * a21 * MC_SCR(K1, i, o); break;
    case ST_3: x = up + dt * (a31 * MC_SCR(K1, i, o) + a32 * MC_SCR(PK_2, i, o)); break;
    case ST_4: x = up + dt * (a41 * MC_SCR(K1, i, o) + a42 * MC_SCR(PK_2, i, o) + a43 * MC_SCR(PK_3, i, o)); break;
    case ST_5: x = up + dt * (a51 * MC_SCR(K1, i, o) + a52 * MC_SCR(PK_2, i, o) + a53 * MC_SCR(PK_3, i, o) + a54 * MC_SCR(PK_4, i, o)); break;
    case ST_6: x = up + dt * (a61 * MC_SCR(K1, i, o) + a62 * MC_SCR(PK_2, i, o) + a63 * MC_SCR(PK_3, i, o) + a64 * MC_SCR(PK_4, i, o) + a65 * MC_SCR(PK_5, i, o)); break;
    case ST_7:
        x = up + dt * (a71 * MC_SCR(K1, i, o) + a72 * MC_SCR(PK_2, i, o) + a73 * MC_SCR(PK_3, i, o) + a74 * MC_SCR(PK_4, i, o) + a75 * MC_SCR(PK_5, i, o) + a76 * MC_SCR(PK_6, i, o));
        MC_SCR(U, i, o) = x;
        break;
    default: x = up; break;                       /* ST_INIT_F0, ST_REFRESH: the RHS at the current state (u == uprev there) */
    }
    const double Tc = jclamp(x, T_LO, m.T_hi);
    if (stage == ST_7 && Tc != x) block_flag_or(&shw.clampflag[o], 1);
    MC_SCR(P_T, i, o) = Tc;
    if (i < m.N - 1)
        MC_SCR(P_LAM, i, o) = soil_conductivity(Tc, sh.P[o], sh.p0p[o], sh.pp0[o], MC_SCR(P_TH, i, o), MC_SCR(P_PM, i, o), MC_SCR(P_PG, i, o),
                                                MC_SCR(P_FW, i, o), MC_SCR(P_LM, i, o)) / m.dzg[i];      /* the conductance lambda / dz of the layer below node i */
}

/* the error-estimate term of one node (stage 7), given its fresh k7 */
template <int NA>
MC_DEV double coop2_error_term(const Coop2Shared& sh, double* scr, int o, int i, double k7)
{
    using namespace tsit5;
    const DevModel& m = MC_CONST_MODEL;
    const int code = sh.code[o];
    const bool kswap = (code >> 4) & 1, uswap = (code >> 5) & 1;
    const int K1 = kswap ? PK_B : PK_A, U = uswap ? PU_B : PU_A, UP = uswap ? PU_A : PU_B;
    const double dt = sh.dt[o];
    const double ut = dt * (bt1 * MC_SCR(K1, i, o) + bt2 * MC_SCR(PK_2, i, o) + bt3 * MC_SCR(PK_3, i, o) + bt4 * MC_SCR(PK_4, i, o) + bt5 * MC_SCR(PK_5, i, o)
                            + bt6 * MC_SCR(PK_6, i, o) + bt7 * k7);
    return ut / (m.abstol + fmax(fabs(MC_SCR(UP, i, o)), fabs(MC_SCR(U, i, o))) * m.reltol);
}

/* pass 2, one (site o, node i >= 1) item: conduction (heat_transport_1d.jl:192-210) */
template <int NA>
MC_DEV void coop2_item_conduct(const Coop2Shared& sh, double* scr, int o, int i)
{
    const DevModel& m = MC_CONST_MODEL;
    const int code = sh.code[o], stage = code & 15;
    double du = 0.0;
    if (i < m.N - 1) {
        const double T_up = MC_SCR(P_T, i - 1, o), T_i = MC_SCR(P_T, i, o), T_dn = MC_SCR(P_T, i + 1, o);
        const double G_up = MC_SCR(P_LAM, i - 1, o);
        const double G_i = MC_SCR(P_LAM, i, o);
        du = 60.0 * (G_up * (T_up - T_i) + G_i * (T_dn - T_i)) / MC_SCR(P_CAP, i, o);
    }
    MC_SCR(stage_out_array(stage, (code >> 4) & 1), i, o) = du;
    if (stage == ST_7) MC_SCR(P_Q, i, o) = coop2_error_term<NA>(sh, scr, o, i, du);
}

/* owner: forcing at the stage time + pressure terms for pass 1 */
MC_DEV void coop2_pre(const HourPar& hp, double t, RhsCtx& R, Coop2Shared& sh)
{
    const DevModel& m = MC_CONST_MODEL; const int tid = MC_TID;
    const double xk = t / 60.0;
    int ik = (int)floor(xk);
    ik = ik < 0 ? 0 : (ik > 23 ? 23 : ik);
    const double dk = xk - (double)ik;
    R.P = knot_lerp(hp.fP, hp.S, ik, dk);
    R.T_air = knot_lerp(hp.fT, hp.S, ik, dk);
    R.wind = jmax(knot_lerp(hp.fwind, hp.S, ik, dk), 0.1);
    R.zen = jmin(90.0, nearbyint(knot_lerp(hp.fzen, hp.S, ik, dk) * 1000.0) / 1000.0);
    R.solar = jmax(0.0, knot_lerp(hp.fsolar, hp.S, ik, dk));
    R.cloud = jclamp(knot_lerp(hp.fcloud, hp.S, ik, dk), 0.0, 1.0);
    R.rh = jclamp(knot_lerp(hp.frh, hp.S, ik, dk), 0.0, 1.0);
    R.szen = 0.0; R.fd = 0.0;
    if (hp.slope > 0.0 && R.zen < 90.0) {
        R.szen = jmin(90.0, knot_lerp(hp.fszen, hp.S, ik, dk));
        R.fd = jclamp(knot_lerp(hp.fdiff, hp.S, ik, dk), 0.0, 1.0);
    }
    sh.P[tid] = R.P; sh.p0p[tid] = m.P0 / R.P; sh.pp0[tid] = R.P / m.P0;
}

/* owner: surface energy balance of node 0 (heat_transport_1d.jl:212-302), same arithmetic as soil_rhs */
template <int NA>
MC_DEV_NOINLINE double coop2_surface(const HourPar& hp, const RhsCtx& R, const double* scr, Counters& cn)
{
    const DevModel& m = MC_CONST_MODEL; const int tid = MC_TID;
    const double P = R.P, T_air = R.T_air, wind = R.wind, zen = R.zen, solar = R.solar, cloud = R.cloud, rh = R.rh;
    const double Ts_node = MC_SCR(P_T, 0, tid);
    const double G_up = MC_SCR(P_LAM, 0, tid);
    const double Q_cond = G_up * (MC_SCR(P_T, 1, tid) - Ts_node);
    double Q_solar = (1.0 - hp.albedo) * solar * (1.0 - hp.shade);
    if (hp.slope > 0.0 && zen < 90.0) {
        const double cz = mcm::cos(zen * PI_D / 180.0), csz = mcm::cos(R.szen * PI_D / 180.0);
        Q_solar = jmax(0.0, (((1.0 - R.fd) * Q_solar) / cz) * csz) + (R.fd * Q_solar) * hp.svf;
    }
    const WetAir wa = wet_air(T_air, rh, P);
    const double atm = atmospheric_radiation(wa.vapour_pressure, T_air);
    const double cloud_rad = m.sigma * hp.eps_cloud * mcm::pow4(T_air - 2.0);
    const double hill_rad = m.sigma * hp.eps_cloud * mcm::pow4(T_air);
    const double lw_sky = (atm * (1.0 - cloud) + cloud_rad * cloud) * (1.0 - hp.shade);
    const double incoming = (lw_sky + hp.shade * hill_rad) * hp.svf + hill_rad * (1.0 - hp.svf);
    const double Q_ir = incoming - ((1.0 - hp.shade) * m.sigma * hp.eps_surf) * mcm::pow4(Ts_node) - hp.shade * hill_rad;
    double Ts = Ts_node;
    if (fabs(Ts - T0C) > 100.0) Ts = T_air + 1.0;
    if (Ts == T_air) Ts = Ts + 0.1;
    const double dT = T_air - Ts;
    const double rho_cp = calc_rho_cp((Ts + T_air) / 2.0);
    double Qc;
    if (T_air >= Ts || zen >= 90.0) {
        const double ustar = jmax(m.kappa * wind / hp.lz, 1.0e-6);
        Qc = convective_flux(rho_cp, dT, ustar, 0.64 / hp.lz, sublayer_stanton(hp.z0, ustar));
    } else {
        const Obukhov o = calc_obukhov(wind, hp.z0, m.zref, hp.lz, dT, rho_cp, hp.L_prev >= 0.0 ? -0.3 : hp.L_prev);
        cn.n_obukhov += o.iters;
        Qc = o.Q;
    }
    const double hc = dT == 0.0 ? 0.5 : jmax(fabs(Qc / dT), 0.5);
    const double hd = (hc / (wa.specific_heat * wa.density)) * m.schmidt_factor;
    const double Tsc = Ts < T_LO ? T_LO : Ts;
    const double Qe = (hp.soil_wetness * hd * (vapour_density_sat(Tsc) - wa.vapour_density)) * latent_heat(Tsc);
    return 60.0 * (Q_solar + Q_ir + Q_cond + Qc + 0.0 - Qe) / MC_SCR(P_CAP, 0, tid);
}

#define MC_YIELD2(label, stg, tt) do { I.st = (label); I.stage = (stg); I.req_t = (tt); return true; case (label):; } while (0)

/* Resumable adaptive integration to I.target (see coop_advance): returns true with the next stage published in (stage, req_t). */
template <int NA>
MC_DEV_NOINLINE bool coop2_advance(Own2& I, Coop2Shared& sh, double* scr, Counters& cn)
{
    using namespace tsit5;
    const DevModel& m = MC_CONST_MODEL; const int N = m.N; const int tid = MC_TID;
    switch (I.st) {
    case 0:
        if (I.need_initdt) {                                             /* Hairer-Wanner initial step: two RHS evaluations */
            I.need_initdt = false;
            { double s0 = 0.0; const int UP = I.uparr();
              MC_ROLLED for (int i = 0; i < N; i++) { const double ui = MC_SCR(UP, i, tid); const double sk = m.abstol + fabs(ui) * m.reltol; const double q = ui / sk; s0 += q * q; }
              I.w0 = sqrt(s0 / N); }
            MC_YIELD2(1, ST_INIT_F0, I.t);
            { bool finite = true;
              MC_ROLLED for (int i = 0; i < N; i++) if (!isfinite(MC_SCR(PK_2, i, tid))) finite = false;
              if (!finite) { I.dt = 1e-12; goto initdt_done; } }
            { double s1 = 0.0; const int UP = I.uparr();
              MC_ROLLED for (int i = 0; i < N; i++) { const double sk = m.abstol + fabs(MC_SCR(UP, i, tid)) * m.reltol; const double q = MC_SCR(PK_2, i, tid) / sk; s1 += q * q; }
              I.w1 = sqrt(s1 / N); }
            I.w2 = (I.w0 < 1e-5 || I.w1 < 1e-5) ? 1e-6 : (I.w0 / I.w1) / 100.0;
            I.w2 = fmin(I.w2, 1440.0);
            sh.dt[tid] = I.w2;
            MC_YIELD2(2, ST_INIT_F1, I.t + I.w2);
            { double s2 = 0.0; const int UP = I.uparr();
              MC_ROLLED for (int i = 0; i < N; i++) { const double sk = m.abstol + fabs(MC_SCR(UP, i, tid)) * m.reltol; const double q = (MC_SCR(PK_3, i, tid) - MC_SCR(PK_2, i, tid)) / sk; s2 += q * q; }
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
                        I.uswap = !I.uswap; I.u_stale = true;            /* uprev <- u */
                        I.dt = I.dtpropose;
                        if (I.u_modified) { MC_YIELD2(3, ST_REFRESH, I.t); }
                        else I.kswap = !I.kswap;                          /* k1 <- k7 (FSAL) */
                    }
                } else if (I.u_modified) {
                    const int U = I.uarr(), UP = I.uparr();
                    MC_ROLLED for (int i = 0; i < N; i++) MC_SCR(UP, i, tid) = MC_SCR(U, i, tid);
                }
                I.iter++;
                I.dt = fmin(1440.0, I.dt);
                I.dt = fmin(I.dt, I.tstop - I.t);
                if (!(I.dt > 1e-12) || ++I.guard > 100000) {
                    I.flags |= (I.dt != I.dt) ? MC_STATUS_NAN : (I.guard > 100000 ? MC_STATUS_STEP_LIMIT : MC_STATUS_DT_UNDERFLOW);
                    I.t = I.tstop; I.accept = true; I.dtpropose = 1.0;
                    const int K1 = I.k1arr(), K7 = I.k7arr(), U = I.uarr(), UP = I.uparr();
                    MC_ROLLED for (int i = 0; i < N; i++) MC_SCR(K7, i, tid) = MC_SCR(K1, i, tid);
                    if (I.u_stale) { MC_ROLLED for (int i = 0; i < N; i++) MC_SCR(U, i, tid) = MC_SCR(UP, i, tid); I.u_stale = false; }
                    sh.clampflag[tid] = 1;                               /* force the exact check below */
                    break;
                }
                I.st_t = I.t; I.st_dt = I.dt;
                sh.dt[tid] = I.st_dt;
                MC_YIELD2(4, ST_2, I.st_t + c1 * I.st_dt);
                MC_YIELD2(5, ST_3, I.st_t + c2 * I.st_dt);
                MC_YIELD2(6, ST_4, I.st_t + c3 * I.st_dt);
                MC_YIELD2(7, ST_5, I.st_t + c4 * I.st_dt);
                MC_YIELD2(8, ST_6, I.st_t + I.st_dt);
                sh.clampflag[tid] = 0;
                MC_YIELD2(9, ST_7, I.st_t + I.st_dt);
                {
                    double ss = 0.0;
                    ss += I.q0 * I.q0;
                    MC_ROLLED for (int i0 = 1; i0 < N; i0 += 6) {      /* in-order sum; the loads of a batch are issued together */
                        double qv[6];
                        _Pragma("unroll") for (int k = 0; k < 6; k++) qv[k] = i0 + k < N ? MC_SCR(P_Q, i0 + k, tid) : 0.0;
                        _Pragma("unroll") for (int k = 0; k < 6; k++) if (i0 + k < N) ss += qv[k] * qv[k];
                    }
                    const double EEst = sqrt(ss / N);
                    I.u_modified = false;
                    double q;
                    if (EEst == 0.0) q = 0.1;
                    else {
                        I.q11 = fastpow(EEst, 7.0 / 50.0);
                        q = I.q11 / fastpow(I.qold, 2.0 / 25.0);
                        q = fmax(0.1, fmin(5.0, q / 0.9));
                    }
                    I.accept = EEst <= 1.0;
                    if (EEst != EEst) { I.flags |= MC_STATUS_NAN; I.accept = true; q = 1.0; }
                    if (I.accept) {
                        cn.n_accept++;
                        I.qold = fmax(EEst, 1e-4);
                        const double dtnew = I.dt / q;
                        double ttmp = I.t + I.dt;
                        if (fabs(ttmp - I.tstop) < 100.0 * 2.220446049250313e-16 * fmax(I.t, I.tstop)) ttmp = I.tstop;
                        I.t = ttmp;
                        I.dtpropose = fmin(1440.0, dtnew);
                        I.u_stale = false;
                        break;
                    }
                    cn.n_reject++;
                    I.u_stale = true;                                     /* u <- uprev: the array of u now holds the rejected proposal */
                }
            }
            if (sh.clampflag[tid]) {   /* re-clamp the state after every internal step (simulation.jl:891-900); pass 1 flags candidates */
                const int U = I.uarr(), UP = I.uparr();
                bool changed = false;
                MC_ROLLED for (int k = 0; k < N; k++) { const double v = MC_SCR(U, k, tid); const double c = jclamp(v, T_LO, m.T_hi); if (c != v) { MC_SCR(U, k, tid) = c; changed = true; } }
                if (changed) { MC_ROLLED for (int k = 0; k < N; k++) MC_SCR(UP, k, tid) = MC_SCR(U, k, tid); I.u_modified = true; }
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

/* Drive every site of the block through the RHS evaluations its state machine requests until all have reached their target. */
template <int NA>
MC_DEV void coop2_rounds(Own2& I, bool pending, const HourPar& hp, Coop2Shared& sh, double* scr, Counters& cn)
{
    RhsCtx R;
    const DevModel& m = MC_CONST_MODEL; const int N = m.N;
    const int tid = MC_TID, TB = MC_TB;
    /* The request list of round r is counted in sh.count[r % 3]; after the round's barrier thread 0 clears the counter of round r + 2.  Three
     * counters, cleared two rounds ahead, make every access race-free without extra barriers: the counter a thread is still reading (round r)
     * is not written before round r + 3's appends, and the counter of round r + 1 was cleared during round r - 1.  The rotation index is
     * block-uniform and carried across calls and site batches (every thread of a block runs the same number of rounds). */
    int par = I.par;
    for (;;) {
        if (pending) {
            coop2_pre(hp, I.req_t, R, sh);
            sh.code[tid] = (unsigned char)(I.stage | (I.kswap ? 16 : 0) | (I.uswap ? 32 : 0));
        }
        block_list_append(pending, sh.list, &sh.count[par], tid);        /* the requesting sites of this round */
        MC_BLOCK_SYNC();
        const int A = sh.count[par];
        if (tid == 0) sh.count[par == 0 ? 2 : par - 1] = 0;      /* (par + 2) % 3 */
        par = par == 2 ? 0 : par + 1;
        if (A == 0) break;
        /* item j = node * A + (index into the list); j / A by multiplication (exact for j < 2^20, A <= 1024) */
        const unsigned long long magic = (0x100000000ull + (unsigned)A - 1ull) / (unsigned)A;
        {   /* pass 1: stage vector, clamp, conductivity */
            const int total = A * N;
            MC_ROLLED for (int j = tid; j < total; j += TB) { const int node = (int)(((unsigned long long)j * magic) >> 32); coop2_item_stage<NA>(sh, sh, scr, sh.list[j - node * A], node); }
        }
        MC_BLOCK_SYNC();
        {   /* pass 2: conduction of nodes 1..N-1 */
            const int total = A * (N - 1);
            MC_ROLLED for (int j = tid; j < total; j += TB) { const int node = (int)(((unsigned long long)j * magic) >> 32); coop2_item_conduct<NA>(sh, scr, sh.list[j - node * A], node + 1); }
        }
        if (pending) {
            cn.n_rhs++;
            const double du0 = coop2_surface<NA>(hp, R, scr, cn);
            MC_SCR(stage_out_array(I.stage, I.kswap), 0, tid) = du0;
            if (I.stage == ST_7) I.q0 = coop2_error_term<NA>(sh, scr, tid, 0, du0);
        }
        MC_BLOCK_SYNC();
        if (pending) pending = coop2_advance<NA>(I, sh, scr, cn);
    }
    I.par = par;
}
