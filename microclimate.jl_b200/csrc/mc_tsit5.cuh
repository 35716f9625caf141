/*
 * mc_tsit5.cuh - Adaptive Tsit5 with OrdinaryDiffEq's PI controller: per-site (thread-local) form.
 * Part of mc_device.cuh (included from there, inside namespace mcd, in this order); not a stand-alone header.
 */

/* ------------------------------------------------------------------------------------------
 * Adaptive Tsit5 with OrdinaryDiffEq's PI controller (third-party arithmetic restated; see the
 * oracle header for provenance).  Integrator state lives in thread-local arrays.
 * ------------------------------------------------------------------------------------------ */
namespace tsit5 {
constexpr double c1 = 0.161, c2 = 0.327, c3 = 0.9, c4 = 0.9800255409045097;
constexpr double a21 = 0.161;
constexpr double a31 = -0.008480655492356989, a32 = 0.335480655492357;
constexpr double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
constexpr double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365, a54 = -0.09249506636175525;
constexpr double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159, a64 = -0.071584973281401, a65 = -0.028269050394068383;
constexpr double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996, a74 = 1.379008574103742, a75 = -3.290069515436081, a76 = 2.324710524099774;
constexpr double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469, bt3 = 0.007880878010261995,
                 bt4 = -0.1447110071732629, bt5 = 0.5823571654525552, bt6 = -0.45808210592918697, bt7 = 0.015151515151515152;
}

/* FastPower.fastpower(x, y): Float32 log2 rational fit + exp2, as the PI controller uses it */
MC_DEV float fastlog2f(float x)
{
    union { float f; uint32_t u; } v; v.f = x;
    const uint32_t ux = v.u;
    const uint32_t ex = (ux & 0x7F800000u) >> 23;
    float signif, fexp;
    if (ux & 0x00400000u) { v.u = (ux & 0x007FFFFFu) | 0x3f000000u; signif = v.f; fexp = (float)ex - 126.0f; }
    else                  { v.u = (ux & 0x007FFFFFu) | 0x3f800000u; signif = v.f; fexp = (float)ex - 127.0f; }
    signif = signif - 1.0f;
#if defined(__CUDACC__)
    return __fadd_rn(fexp, __fdiv_rn(__fmul_rn(signif, __fadd_rn(__fmul_rn(0.338953f, signif), 2.198599f)), __fadd_rn(signif, 1.523692f)));
#else
    return fexp + signif * (0.338953f * signif + 2.198599f) / (signif + 1.523692f);
#endif
}
MC_DEV double fastpow(double x, double y)
{
#if defined(__CUDACC__)
    const float l = __fmul_rn(fastlog2f((float)x), (float)y);
#else
    const float l = fastlog2f((float)x) * (float)y;
#endif
    return (double)(float)mcm::exp2((double)l);
}

template <int NA>
struct Integ {
    double u[NA], uprev[NA], k1[NA], k2[NA], k3[NA], k4[NA], k5[NA], k6[NA], k7[NA], tmp[NA];
    double t, dt, dtpropose, qold, q11, tstop;
    int iter, flags; bool accept, u_modified;
};

MC_DEV double rms_scaled(const double* x, const double* sk, int n) { double s = 0.0; MC_ROLLED for (int i = 0; i < n; i++) { const double q = x[i] / sk[i]; s += q * q; } return sqrt(s / n); }

/* Hairer-Wanner automatic initial step (OrdinaryDiffEqCore initdt.jl), order 5 */
template <int NA, int NS>
MC_DEV_NOINLINE double determine_initdt(Integ<NA + NS>& I, const HourSoil<NA>& hs, const HourPar& hp, Counters& cn)
{
    const DevModel& m = MC_CONST_MODEL; const int N = m.N + NS;       /* ODE size */
    double* sk = I.tmp; double* f0 = I.k2; double* f1 = I.k3; double* u1 = I.k4;
    double s0 = 0.0;
    MC_ROLLED for (int i = 0; i < N; i++) { sk[i] = m.abstol + fabs(I.u[i]) * m.reltol; const double q = I.u[i] / sk[i]; s0 += q * q; }
    const double d0 = sqrt(s0 / N);
    rhs_dispatch<NA, NS>(hs, hp, I.u, I.t, f0, cn);
    bool finite = true;
    MC_ROLLED for (int i = 0; i < N; i++) if (!isfinite(f0[i])) finite = false;
    if (!finite) return 1e-12;
    const double d1 = rms_scaled(f0, sk, N);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
    dt0 = fmin(dt0, 1440.0);
    MC_ROLLED for (int i = 0; i < N; i++) u1[i] = I.u[i] + dt0 * f0[i];
    rhs_dispatch<NA, NS>(hs, hp, u1, I.t + dt0, f1, cn);
    double s2 = 0.0;
    MC_ROLLED for (int i = 0; i < N; i++) { const double q = (f1[i] - f0[i]) / sk[i]; s2 += q * q; }
    const double d2 = sqrt(s2 / N) / dt0;
    const double md = fmax(d1, d2);
    double dt1;
    if (md <= 1e-15) dt1 = fmax(1e-6, dt0 * 1e-3);
    else dt1 = mcm::exp10(-(2.0 + mcm::log10(md)) / 5.0);
    return fmin(fmin(100.0 * dt0, dt1), 1440.0);
}

/* one accepted step: loopheader!/perform_step!/loopfooter! until accepted */
template <int NA, int NS>
MC_DEV_NOINLINE void tsit5_step(Integ<NA + NS>& I, const HourSoil<NA>& hs, const HourPar& hp, Counters& cn)
{
    using namespace tsit5;
    const DevModel& m = MC_CONST_MODEL; const int N = m.N + NS;       /* ODE size */
    int guard = 0;
    for (;;) {
        if (I.iter > 0) {
            if (!I.accept) {
                I.dt = I.dt / fmin(5.0, I.q11 / 0.9);
            } else {
                MC_UNROLL4 for (int i = 0; i < N; i++) I.uprev[i] = I.u[i];
                I.dt = I.dtpropose;
                if (I.u_modified) rhs_dispatch<NA, NS>(hs, hp, I.u, I.t, I.k1, cn);
                else MC_UNROLL4 for (int i = 0; i < N; i++) I.k1[i] = I.k7[i];
            }
        } else if (I.u_modified) {
            MC_UNROLL4 for (int i = 0; i < N; i++) I.uprev[i] = I.u[i];
        }
        I.iter++;
        I.dt = fmin(1440.0, I.dt);
        I.dt = fmin(I.dt, I.tstop - I.t);
        if (!(I.dt > 1e-12) || ++guard > 100000) {
            I.flags |= (I.dt != I.dt) ? MC_STATUS_NAN : (guard > 100000 ? MC_STATUS_STEP_LIMIT : MC_STATUS_DT_UNDERFLOW);
            I.t = I.tstop; I.accept = true; I.dtpropose = 1.0;
            MC_UNROLL4 for (int i = 0; i < N; i++) I.k7[i] = I.k1[i];
            return;
        }
        const double t = I.t, dt = I.dt;
        MC_UNROLL4 for (int i = 0; i < N; i++) I.tmp[i] = I.uprev[i] + dt * a21 * I.k1[i];
        rhs_dispatch<NA, NS>(hs, hp, I.tmp, t + c1 * dt, I.k2, cn);
        MC_UNROLL4 for (int i = 0; i < N; i++) I.tmp[i] = I.uprev[i] + dt * (a31 * I.k1[i] + a32 * I.k2[i]);
        rhs_dispatch<NA, NS>(hs, hp, I.tmp, t + c2 * dt, I.k3, cn);
        MC_UNROLL4 for (int i = 0; i < N; i++) I.tmp[i] = I.uprev[i] + dt * (a41 * I.k1[i] + a42 * I.k2[i] + a43 * I.k3[i]);
        rhs_dispatch<NA, NS>(hs, hp, I.tmp, t + c3 * dt, I.k4, cn);
        MC_UNROLL4 for (int i = 0; i < N; i++) I.tmp[i] = I.uprev[i] + dt * (a51 * I.k1[i] + a52 * I.k2[i] + a53 * I.k3[i] + a54 * I.k4[i]);
        rhs_dispatch<NA, NS>(hs, hp, I.tmp, t + c4 * dt, I.k5, cn);
        MC_UNROLL4 for (int i = 0; i < N; i++) I.tmp[i] = I.uprev[i] + dt * (a61 * I.k1[i] + a62 * I.k2[i] + a63 * I.k3[i] + a64 * I.k4[i] + a65 * I.k5[i]);
        rhs_dispatch<NA, NS>(hs, hp, I.tmp, t + dt, I.k6, cn);
        MC_UNROLL4 for (int i = 0; i < N; i++) I.u[i] = I.uprev[i] + dt * (a71 * I.k1[i] + a72 * I.k2[i] + a73 * I.k3[i] + a74 * I.k4[i] + a75 * I.k5[i] + a76 * I.k6[i]);
        rhs_dispatch<NA, NS>(hs, hp, I.u, t + dt, I.k7, cn);
        double ss = 0.0;
        MC_UNROLL4 for (int i = 0; i < N; i++) {
            const double ut = dt * (bt1 * I.k1[i] + bt2 * I.k2[i] + bt3 * I.k3[i] + bt4 * I.k4[i] + bt5 * I.k5[i] + bt6 * I.k6[i] + bt7 * I.k7[i]);
            const double q = ut / (m.abstol + fmax(fabs(I.uprev[i]), fabs(I.u[i])) * m.reltol);
            ss += q * q;
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
            return;
        }
        cn.n_reject++;
        MC_UNROLL4 for (int i = 0; i < N; i++) I.u[i] = I.uprev[i];
    }
}
