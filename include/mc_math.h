/*
 * mc_math.h — portable, bit-reproducible double-precision elementary functions.
 *
 * WHY: the soil path is ill-conditioned at its discontinuities (adaptive step accept/reject, the 1 %
 * Obukhov stopping rule, the |residual| <= 10 J/kg stomatal Newton exit, Float32 step-size controller):
 * a 1-ulp difference in ONE libm call is amplified to 1e-2 K / 1e-2 m3/m3 within simulated days (the
 * CPU oracle shows this against itself, tests/test_conditioning.py).  CUDA's libm and the host's libm
 * differ in the last bit, so parity through vendor libms can only ever be statistical.  Every
 * transcendental below is built only from IEEE-754 + - * / sqrt fma and integer bit operations,
 * which are correctly rounded on both x86-64 and sm_100a, so the CUDA engine and the CPU oracle
 * produce IDENTICAL bits when compiled without floating-point contraction (-fmad=false / -ffp-contract=off).
 * The reference itself (Julia) uses its own pure-Julia libm, not glibc; agreement of these functions
 * with glibc to <= 1-2 ulp (pow/exp10: a few 1e-16 relative) is pinned in tests/test_mc_math.py.
 *
 * Algorithms: the classic fdlibm kernels (exp, log, atan, sin/cos kernels), restated; pow and exp10
 * are exp(y*log x) with the rounding error of the product carried through an fma.
 *
 * Shared by microclimate.jl_b200/csrc (device + host-emulation) and oracle/ (host).
 */
#ifndef MC_MATH_H
#define MC_MATH_H
#ifndef MCM_DIV_NOINLINE
#define MCM_DIV_NOINLINE 0
#endif

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define MCM_FN __host__ __device__ __forceinline__
/* the big kernels (exp, log, atan) are real functions on the device: ONE copy each in the instruction cache instead of
 * one per call site (the soil right-hand side alone has ~20 call sites) */
#define MCM_HEAVY __host__ __device__ __noinline__
#else
#define MCM_FN static inline
#define MCM_HEAVY static inline
#endif

namespace mcm {

#if defined(__CUDA_ARCH__) && MCM_DIV_NOINLINE
__device__ __noinline__ double ddiv(double a, double b) { return a / b; }
#define MCM_DIV(a, b) mcm::ddiv((a), (b))
#else
#define MCM_DIV(a, b) ((a) / (b))
#endif

#if defined(__CUDA_ARCH__)
MCM_FN uint64_t d2u(double x) { return (uint64_t)__double_as_longlong(x); }
MCM_FN double u2d(uint64_t u) { return __longlong_as_double((long long)u); }
MCM_FN double ffma(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
MCM_FN uint64_t d2u(double x) { union { double d; uint64_t u; } v; v.d = x; return v.u; }
MCM_FN double u2d(uint64_t u) { union { double d; uint64_t u; } v; v.u = u; return v.d; }
MCM_FN double ffma(double a, double b, double c) { return __builtin_fma(a, b, c); }
#endif
MCM_FN int32_t hi_word(double x) { return (int32_t)(d2u(x) >> 32); }
MCM_FN uint32_t lo_word(double x) { return (uint32_t)(d2u(x) & 0xffffffffu); }
MCM_FN double with_hi(double x, int32_t hi) { return u2d(((uint64_t)(uint32_t)hi << 32) | (d2u(x) & 0xffffffffu)); }

/* coefficient table: on the device it lives in constant memory so DFMA reads c[bank][offset] operands directly
 * instead of materialising 64-bit immediates (two UMOVs per constant per use) */
#if defined(__CUDA_ARCH__)
#define MCM_TAB mcm_tab_c
#else
#define MCM_TAB mcm_tab_h
#endif
#define MCM_TAB_INIT { \
    /* 0..11: 1/n!, n = 2..13 */ \
    0.5, 1.6666666666666666e-01, 4.1666666666666664e-02, 8.3333333333333332e-03, 1.3888888888888889e-03, 1.9841269841269841e-04, \
    2.4801587301587302e-05, 2.7557319223985893e-06, 2.7557319223985888e-07, 2.5052108385441720e-08, 2.0876756987868100e-09, 1.6059043836821613e-10, \
    /* 12..14: ln2_hi, ln2_lo, 1/ln2 */ \
    6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.44269504088896338700e+00, \
    /* 15..21: Lg1..Lg7 */ \
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01, \
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01 }
#if defined(__CUDACC__)
__constant__ double mcm_tab_c[22] = MCM_TAB_INIT;
#endif
static const double mcm_tab_h[22] = MCM_TAB_INIT;

#if defined(__CUDA_ARCH__)
MCM_FN double rne(double x) { return rint(x); }
#else
MCM_FN double rne(double x) { return __builtin_nearbyint(x); }
#endif

/* ---- exp ---------------------------------------------------------------------------------- */
/* rare arguments (|x| >= 690, NaN): the classic fdlibm e_exp.c evaluation incl. overflow / gradual underflow */
MCM_HEAVY double exp_slow(double x)
{
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return u2d(0x7ff0000000000000ULL);
    if (x < -7.45133219101941108420e+02) return 0.0;
    const int k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
    const double hi = x - (double)k * ln2HI;
    const double lo = (double)k * ln2LO;
    x = hi - lo;
    const double t = x * x;
    const double c = x - t * ffma(t, ffma(t, ffma(t, ffma(t, P5, P4), P3), P2), P1);
    const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) {
        if (k == 1024) return y * 2.0 * 8.98846567431158e307;   /* 2^1023 */
        return u2d(d2u(y) + ((uint64_t)(int64_t)k << 52));
    }
    return u2d(d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52)) * 9.33263618503218878990e-302;   /* 2^-1000 */
}

/* exp(x) = 2^k * exp(r), r = x - k ln2, |r| <= 0.347: degree-13 Taylor polynomial evaluated Estrin-style (division-free,
 * branch-free, short dependency chain); truncation 4e-18, total error < 1 ulp */
MCM_FN bool exp_fast_ok(double x) { return fabs(x) < 690.0; }
MCM_FN double exp_core(double x)
{
    const double kd = rne(x * MCM_TAB[14]);
    const double hi = x - kd * MCM_TAB[12];          /* exact: ln2_hi has 21 trailing zero bits, |k| < 2^10 */
    const double lo = kd * MCM_TAB[13];
    const double r = hi - lo;
    const double r2 = r * r;
    const double r4 = r2 * r2;
    const double q01 = ffma(r, MCM_TAB[1], MCM_TAB[0]);
    const double q23 = ffma(r, MCM_TAB[3], MCM_TAB[2]);
    const double q45 = ffma(r, MCM_TAB[5], MCM_TAB[4]);
    const double q67 = ffma(r, MCM_TAB[7], MCM_TAB[6]);
    const double q89 = ffma(r, MCM_TAB[9], MCM_TAB[8]);
    const double qab = ffma(r, MCM_TAB[11], MCM_TAB[10]);
    const double a = ffma(r2, q23, q01);
    const double b = ffma(r2, q67, q45);
    const double c = ffma(r2, qab, q89);
    const double q = ffma(r4, ffma(r4, c, b), a);
    const double y = 1.0 + ffma(r2, q, r);
    return u2d(d2u(y) + ((uint64_t)(int64_t)kd << 52));
}
MCM_HEAVY double exp(double x)
{
    if (!exp_fast_ok(x)) return exp_slow(x);
    return exp_core(x);
}

/* ---- log ---------------------------------------------------------------------------------- */
/* rare arguments (zero, negative, subnormal, inf, NaN) */
MCM_HEAVY double log_slow(double x)
{
    if (x != x) return x + x;
    if (x == 0.0) return -u2d(0x7ff0000000000000ULL);
    if (x < 0.0) return u2d(0x7ff8000000000000ULL);
    if (x == u2d(0x7ff0000000000000ULL)) return x;
    /* positive subnormal: scale by 2^54 and correct */
    const double xs = x * 1.80143985094819840000e+16;
    int32_t hx = hi_word(xs);
    int k = (hx >> 20) - 1023 - 54;
    hx &= 0x000fffff;
    const int32_t i0 = (hx + 0x95f64) & 0x100000;
    const double m = with_hi(xs, hx | (i0 ^ 0x3ff00000));
    k += (i0 >> 20);
    const double f = m - 1.0, dk = (double)k;
    const double s = MCM_DIV(f, 2.0 + f), z = s * s, w = z * z;
    const double t1 = w * ffma(w, ffma(w, MCM_TAB[20], MCM_TAB[18]), MCM_TAB[16]);
    const double t2 = z * ffma(w, ffma(w, ffma(w, MCM_TAB[21], MCM_TAB[19]), MCM_TAB[17]), MCM_TAB[15]);
    const double R = t2 + t1, hfsq = 0.5 * f * f;
    return dk * MCM_TAB[12] - ((hfsq - (s * (hfsq + R) + dk * MCM_TAB[13])) - f);
}

/* fdlibm e_log.c kernel (the always-hfsq form of the later k_log.h), branch-free for normal positive x */
MCM_FN bool log_fast_ok(double x) { return x >= 2.2250738585072014e-308 && x < u2d(0x7ff0000000000000ULL); }
MCM_FN double log_core(double x)
{
    int32_t hx = hi_word(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int32_t i0 = (hx + 0x95f64) & 0x100000;
    const double m = with_hi(x, hx | (i0 ^ 0x3ff00000));      /* m in [sqrt(2)/2, sqrt(2)) */
    k += (i0 >> 20);
    const double f = m - 1.0, dk = (double)k;
    const double s = MCM_DIV(f, 2.0 + f), z = s * s, w = z * z;
    const double t1 = w * ffma(w, ffma(w, MCM_TAB[20], MCM_TAB[18]), MCM_TAB[16]);
    const double t2 = z * ffma(w, ffma(w, ffma(w, MCM_TAB[21], MCM_TAB[19]), MCM_TAB[17]), MCM_TAB[15]);
    const double R = t2 + t1, hfsq = 0.5 * f * f;
    return dk * MCM_TAB[12] - ((hfsq - (s * (hfsq + R) + dk * MCM_TAB[13])) - f);
}
MCM_HEAVY double log(double x)
{
    if (!log_fast_ok(x)) return log_slow(x);
    return log_core(x);
}

MCM_FN double log10(double x) { return log(x) * 4.34294481903251816668e-01; }

/* exp(a*b) with the rounding error of the product a*b (and an optional low word of b) carried through */
MCM_FN double exp_prod(double a, double b, double b_lo)
{
    const double p = a * b;
    const double e = ffma(a, b, -p) + a * b_lo;
    const double r = exp(p);
    return ffma(r, e, r);
}
MCM_FN double exp10(double x) { return exp_prod(x, 2.302585092994046, -2.1707562233822494e-16); }
/* the pieces of exp10 for callers that evaluate several independent arguments side by side (instruction-level parallelism):
 * exp10(x) == exp10_finish(exp(exp10_arg(x)), x, exp10_arg(x)) with exactly the operations of exp10 above */
MCM_FN double exp10_arg(double x) { return x * 2.302585092994046; }
MCM_FN double exp10_finish(double r, double x, double p) { const double e = ffma(x, 2.302585092994046, -p) + x * -2.1707562233822494e-16; return ffma(r, e, r); }
MCM_FN double exp2(double x) { return exp_prod(x, 0.6931471805599453, 2.3190468138462996e-17); }

/* pow for the cases the soil path uses: x > 0 (any y), x == 0, and x < 0 with integer-valued y */
#ifndef MCM_POW_NOINLINE
#define MCM_POW_NOINLINE 1      /* one shared copy on the device: at 32 warps/SM instruction fetch is the top stall, code size matters */
#endif
#if MCM_POW_NOINLINE
MCM_HEAVY
#else
MCM_FN
#endif
double pow(double x, double y)
{
    if (y == 0.0) return 1.0;
    if (x != x || y != y) return x + y;
    if (x == 0.0) return y < 0.0 ? u2d(0x7ff0000000000000ULL) : 0.0;
    if (x < 0.0) {
        const double yi = floor(y);
        if (yi != y) return u2d(0x7ff8000000000000ULL);
        const double r = exp_prod(y, log(-x), 0.0);
        return (yi * 0.5 != floor(yi * 0.5)) ? -r : r;
    }
    return exp_prod(y, log(x), 0.0);
}

/* ---- atan ---------------------------------------------------------------------------------- */
MCM_HEAVY double atan(double x)
{
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
                 aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
                 aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
    const int32_t hx = hi_word(x);
    const int32_t ix = hx & 0x7fffffff;
    double hi_c, lo_c; int id;
    if (ix >= 0x44100000) {                         /* |x| >= 2^66 */
        if (x != x) return x + x;
        const double r = 1.57079632679489655800e+00 + 6.12323399573676603587e-17;
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3fdc0000) {                          /* |x| < 0.4375 */
        if (ix < 0x3e200000) return x;              /* |x| < 2^-29 */
        id = -1; hi_c = 0.0; lo_c = 0.0;
    } else {
        x = x < 0.0 ? -x : x;
        if (ix < 0x3ff30000) {                      /* |x| < 1.1875 */
            if (ix < 0x3fe60000) { id = 0; hi_c = 4.63647609000806093515e-01; lo_c = 2.26987774529616870924e-17; x = (2.0 * x - 1.0) / (2.0 + x); }
            else { id = 1; hi_c = 7.85398163397448278999e-01; lo_c = 3.06161699786838301793e-17; x = (x - 1.0) / (x + 1.0); }
        } else {
            if (ix < 0x40038000) { id = 2; hi_c = 9.82793723247329054082e-01; lo_c = 1.39033110312309984516e-17; x = (x - 1.5) / (1.0 + 1.5 * x); }
            else { id = 3; hi_c = 1.57079632679489655800e+00; lo_c = 6.12323399573676603587e-17; x = -1.0 / x; }
        }
    }
    const double z = x * x;
    const double w = z * z;
    const double s1 = z * ffma(w, ffma(w, ffma(w, ffma(w, ffma(w, aT10, aT8), aT6), aT4), aT2), aT0);
    const double s2 = w * ffma(w, ffma(w, ffma(w, ffma(w, aT9, aT7), aT5), aT3), aT1);
    if (id < 0) return x - x * (s1 + s2);
    const double r = hi_c - ((x * (s1 + s2) - lo_c) - x);
    return hx < 0 ? -r : r;
}

/* ---- cos on [-3pi/4, 3pi/4] (the path only needs zenith angles in [0, 90] degrees) ---------- */
MCM_FN double kernel_cos(double x, double y)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const int32_t ix = hi_word(x) & 0x7fffffff;
    const double z = x * x;
    const double r = z * ffma(z, ffma(z, ffma(z, ffma(z, ffma(z, C6, C5), C4), C3), C2), C1);
    if (ix < 0x3FD33333) return 1.0 - (0.5 * z - (z * r - x * y));
    double qx;
    if (ix > 0x3fe90000) qx = 0.28125;
    else qx = u2d((uint64_t)(uint32_t)(ix - 0x00200000) << 32);
    const double hz = 0.5 * z - qx;
    const double a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}
MCM_FN double kernel_sin(double x, double y)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double v = z * x;
    const double r = ffma(z, ffma(z, ffma(z, ffma(z, S6, S5), S4), S3), S2);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
MCM_FN double cos(double x)
{
    const double ax = x < 0.0 ? -x : x;
    if (ax <= 0.78539816339744830962) return kernel_cos(ax, 0.0);
    if (ax <= 2.35619449019234492885) {             /* one step of pi/2 reduction: cos(x) = -sin(x - pi/2) */
        const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
        const double z = ax - pio2_1;
        const double y0 = z - pio2_1t;
        const double y1 = (z - y0) - pio2_1t;
        return -kernel_sin(y0, y1);
    }
    return u2d(0x7ff8000000000000ULL);              /* out of the supported range on purpose */
}

MCM_FN double pow4(double x) { const double x2 = x * x; return x2 * x2; }

/* ---- sin / cos for |x| <= 64 (solar geometry: day angles <= 2 pi * 366/365, hour angles in [-pi, pi]) ------------------------
 * One Cody-Waite reduction by n * pi/2 with the two-word split of fdlibm's e_rem_pio2.c (first iteration): n * pio2_1 is exact for
 * |n| < 2^20 (pio2_1 has 33 significant bits), the tail is carried as y1.  Outside the range: NaN on purpose. */
MCM_FN int rem_pio2_small(double x, double* y0, double* y1)
{
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    const double fn = rne(x * invpio2);
    const double r = x - fn * pio2_1;
    const double w = fn * pio2_1t;
    *y0 = r - w;
    *y1 = (r - *y0) - w;
    return (int)fn;
}
MCM_FN double sin_r(double x)
{
    const double ax = x < 0.0 ? -x : x;
    if (!(ax <= 64.0)) return u2d(0x7ff8000000000000ULL);
    if (ax <= 0.78539816339744830962) return kernel_sin(x, 0.0);
    double y0, y1;
    const int n = rem_pio2_small(x, &y0, &y1) & 3;
    const double a0 = y0 < 0.0 ? -y0 : y0, a1 = y0 < 0.0 ? -y1 : y1;     /* kernel_cos wants a non-negative argument */
    if (n == 0) return kernel_sin(y0, y1);
    if (n == 1) return kernel_cos(a0, a1);
    if (n == 2) return -kernel_sin(y0, y1);
    return -kernel_cos(a0, a1);
}
MCM_FN double cos_r(double x)
{
    const double ax = x < 0.0 ? -x : x;
    if (!(ax <= 64.0)) return u2d(0x7ff8000000000000ULL);
    if (ax <= 0.78539816339744830962) return kernel_cos(ax, 0.0);
    double y0, y1;
    const int n = rem_pio2_small(x, &y0, &y1) & 3;
    const double a0 = y0 < 0.0 ? -y0 : y0, a1 = y0 < 0.0 ? -y1 : y1;
    if (n == 0) return kernel_cos(a0, a1);
    if (n == 1) return -kernel_sin(y0, y1);
    if (n == 2) return -kernel_cos(a0, a1);
    return kernel_sin(y0, y1);
}
/* sine / cosine of an angle in DEGREES, |deg| <= 720 (Julia's sind / cosd are exact at multiples of 90: so are these) */
MCM_FN void sincosd(double deg, double* s, double* c)
{
    double d = deg;
    while (d >= 360.0) d -= 360.0;                  /* exact: Sterbenz / small integers times 360 */
    while (d < 0.0) d += 360.0;
    int q = 0;                                      /* d = 90 q + r, r in [-45, 45] */
    if (d > 315.0) { d -= 360.0; }
    else if (d > 225.0) { d -= 270.0; q = 3; }
    else if (d > 135.0) { d -= 180.0; q = 2; }
    else if (d > 45.0) { d -= 90.0; q = 1; }
    const double r = d * 1.74532925199432957692e-02;        /* pi / 180 */
    const double sr = kernel_sin(r, 0.0), cr = kernel_cos(r < 0.0 ? -r : r, 0.0);
    if (q == 0) { *s = sr; *c = cr; }
    else if (q == 1) { *s = cr; *c = -sr; }
    else if (q == 2) { *s = -sr; *c = -cr; }
    else { *s = -cr; *c = sr; }
}

/* ---- acos (fdlibm e_acos.c, restated) ---------------------------------------------------------------------------------------- */
MCM_FN double acos_rational(double z)
{
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01,
                 qS4 = 7.70381505559019352791e-02;
    const double p = z * ffma(z, ffma(z, ffma(z, ffma(z, ffma(z, pS5, pS4), pS3), pS2), pS1), pS0);
    const double q = ffma(z, ffma(z, ffma(z, ffma(z, qS4, qS3), qS2), qS1), 1.0);
    return p / q;
}
MCM_FN double acos(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17, pi = 3.14159265358979311600e+00;
    const double ax = x < 0.0 ? -x : x;
    if (!(ax <= 1.0)) return u2d(0x7ff8000000000000ULL);
    if (ax == 1.0) return x > 0.0 ? 0.0 : pi + 2.0 * pio2_lo;
    if (ax < 0.5) {
        if (ax <= 5.5511151231257827e-17) return pio2_hi + pio2_lo;         /* |x| <= 2^-54 */
        const double r = acos_rational(x * x);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (x < 0.0) {                                  /* x <= -0.5 */
        const double z = (1.0 + x) * 0.5;
        const double s = sqrt(z);
        const double r = acos_rational(z);
        const double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    const double z = (1.0 - x) * 0.5;               /* x >= 0.5 */
    const double s = sqrt(z);
    const double df = u2d(d2u(s) & 0xffffffff00000000ULL);
    const double c = (z - df * df) / (s + df);
    const double r = acos_rational(z);
    const double w = r * s + c;
    return 2.0 * (df + w);
}

}  // namespace mcm
#endif /* MC_MATH_H */
