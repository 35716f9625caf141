/*
 * opcount.cpp - the CPU oracle (oracle/mc_oracle.cpp, unchanged) compiled with an OPERATION-COUNTING scalar in place of `double`.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE (like everything under oracle/): produces the source-level FP64 operation counts that bench.py's
 * algorithmic roofline numerator uses (W_alg = sum_k count_k * w_k).  Every + - * / sqrt the reference algorithm performs on a
 * double is tallied (fma = 2), every call into the math kernel (exp, log, log10, exp10, exp2, pow, atan, cos) is tallied as ONE call of its
 * class and later folded at the survey's fixed weights (exp/exp10/exp2/log/log10 = 20, pow = 40, atan/cos/tanh = 25); comparisons, min/max,
 * abs, floor and copies are not counted.  Tallies are kept per REGION of the path - the regions are the events the device counters count
 * (RHS evaluation, Obukhov iteration, Newton iteration, infiltration sub-step setup, integrator step control, everything else per site-hour) -
 * so weights per event follow by division (tools/walg_weights.py).  The values computed are bit-identical to the plain oracle's (the counted
 * scalar wraps a double and performs the same operation), which the tool asserts.
 *
 * How: all standard headers and mc_math.h are included first with the real `double`; then `double` is #defined to the counting type and the
 * oracle source is included verbatim.  The C structs of microclimate_b200.h keep their layout (the counting type is one double).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../include/mc_math.h"

namespace opc {
enum Region { R_OTHER = 0, R_RHS, R_OBUKHOV, R_NEWTON, R_INFIL, R_STEP, R_STOMATAL, R_COUNT };
enum Op { O_ADD = 0, O_MUL, O_DIV, O_SQRT, O_FMA, O_EXP, O_LOG, O_POW, O_TRIG, O_COUNT };
static thread_local int region = R_OTHER;
static thread_local unsigned long long tally[R_COUNT][O_COUNT];
static unsigned long long total[R_COUNT][O_COUNT];
inline void hit(int op) { tally[region][op]++; }
struct Scope { int prev; explicit Scope(int r) : prev(region) { region = r; } ~Scope() { region = prev; } };

struct real {
    double v;
    real() = default;
    real(double x) : v(x) {}
    real(float x) : v(x) {}
    real(int x) : v(x) {}
    real(long x) : v((double)x) {}
    real(long long x) : v((double)x) {}
    real(unsigned x) : v(x) {}
    real(unsigned long x) : v((double)x) {}
    real(unsigned long long x) : v((double)x) {}
    explicit operator int() const { return (int)v; }
    explicit operator long() const { return (long)v; }
    explicit operator long long() const { return (long long)v; }
    explicit operator float() const { return (float)v; }
    explicit operator bool() const { return v != 0.0; }
    real& operator+=(real o) { hit(O_ADD); v += o.v; return *this; }
    real& operator-=(real o) { hit(O_ADD); v -= o.v; return *this; }
    real& operator*=(real o) { hit(O_MUL); v *= o.v; return *this; }
    real& operator/=(real o) { hit(O_DIV); v /= o.v; return *this; }
};
inline real operator+(real a, real b) { hit(O_ADD); return real(a.v + b.v); }
inline real operator-(real a, real b) { hit(O_ADD); return real(a.v - b.v); }
inline real operator*(real a, real b) { hit(O_MUL); return real(a.v * b.v); }
inline real operator/(real a, real b) { hit(O_DIV); return real(a.v / b.v); }
inline real operator-(real a) { return real(-a.v); }
inline real operator+(real a) { return a; }
inline bool operator<(real a, real b) { return a.v < b.v; }
inline bool operator>(real a, real b) { return a.v > b.v; }
inline bool operator<=(real a, real b) { return a.v <= b.v; }
inline bool operator>=(real a, real b) { return a.v >= b.v; }
inline bool operator==(real a, real b) { return a.v == b.v; }
inline bool operator!=(real a, real b) { return a.v != b.v; }
}  // namespace opc

/* uncounted helpers the oracle calls unqualified or through std:: */
namespace std {
inline opc::real fabs(opc::real a) { return opc::real(::fabs(a.v)); }
inline opc::real sqrt(opc::real a) { opc::hit(opc::O_SQRT); return opc::real(::sqrt(a.v)); }
inline opc::real floor(opc::real a) { return opc::real(::floor(a.v)); }
inline bool isfinite(opc::real a) { return ::std::isfinite(a.v); }
inline bool isnan(opc::real a) { return a.v != a.v; }
inline opc::real nearbyint(opc::real a) { return opc::real(::nearbyint(a.v)); }
inline opc::real max(double a, opc::real b) { return a < b.v ? b : opc::real(a); }
inline opc::real max(opc::real a, double b) { return a.v < b ? opc::real(b) : a; }
inline opc::real min(double a, opc::real b) { return b.v < a ? b : opc::real(a); }
inline opc::real min(opc::real a, double b) { return b < a.v ? opc::real(b) : a; }
template <> struct numeric_limits<opc::real> {
    static opc::real quiet_NaN() { return opc::real(numeric_limits<double>::quiet_NaN()); }
    static opc::real infinity() { return opc::real(numeric_limits<double>::infinity()); }
    static opc::real epsilon() { return opc::real(numeric_limits<double>::epsilon()); }
    static opc::real max() { return opc::real(numeric_limits<double>::max()); }
};
}  // namespace std
inline opc::real fabs(opc::real a) { return opc::real(::fabs(a.v)); }
inline opc::real sqrt(opc::real a) { opc::hit(opc::O_SQRT); return opc::real(::sqrt(a.v)); }
inline opc::real floor(opc::real a) { return opc::real(::floor(a.v)); }
inline opc::real nearbyint(opc::real a) { return opc::real(::nearbyint(a.v)); }
inline opc::real fmax(opc::real a, opc::real b) { return opc::real(::fmax(a.v, b.v)); }
inline opc::real fmin(opc::real a, opc::real b) { return opc::real(::fmin(a.v, b.v)); }
inline bool isfinite(opc::real a) { return std::isfinite(a.v); }
/* the math kernel: one call of its class each (their internals are not source-level operations of the reference algorithm) */
namespace mcm {
inline opc::real exp(opc::real a) { opc::hit(opc::O_EXP); return opc::real(mcm::exp(a.v)); }
inline opc::real exp2(opc::real a) { opc::hit(opc::O_EXP); return opc::real(mcm::exp2(a.v)); }
inline opc::real exp10(opc::real a) { opc::hit(opc::O_EXP); return opc::real(mcm::exp10(a.v)); }
inline opc::real log(opc::real a) { opc::hit(opc::O_LOG); return opc::real(mcm::log(a.v)); }
inline opc::real log10(opc::real a) { opc::hit(opc::O_LOG); return opc::real(mcm::log10(a.v)); }
inline opc::real pow(opc::real a, opc::real b) { opc::hit(opc::O_POW); return opc::real(mcm::pow(a.v, b.v)); }
inline opc::real atan(opc::real a) { opc::hit(opc::O_TRIG); return opc::real(mcm::atan(a.v)); }
inline opc::real cos(opc::real a) { opc::hit(opc::O_TRIG); return opc::real(mcm::cos(a.v)); }
inline opc::real pow4(opc::real a) { opc::hit(opc::O_MUL); opc::hit(opc::O_MUL); const double x2 = a.v * a.v; return opc::real(x2 * x2); }
inline opc::real ffma(opc::real a, opc::real b, opc::real c) { opc::hit(opc::O_FMA); return opc::real(mcm::ffma(a.v, b.v, c.v)); }
}  // namespace mcm

#define MCO_OPCOUNT 1
#define MCO_REGION(r) opc::Scope mco_region_scope_(opc::r)
#define double opc::real
#include "../mc_oracle.cpp"
#undef double

extern "C" {
/* add the calling thread's tallies to the totals is done at the end of mco_step_days (single call per run in the tool: the tool runs with 1 thread) */
void opc_reset(void) { std::memset(opc::total, 0, sizeof(opc::total)); std::memset(opc::tally, 0, sizeof(opc::tally)); }
void opc_collect(unsigned long long* out /* [R_COUNT][O_COUNT] */)
{
    for (int r = 0; r < opc::R_COUNT; r++) for (int o = 0; o < opc::O_COUNT; o++) out[r * opc::O_COUNT + o] = opc::tally[r][o] + opc::total[r][o];
}
int opc_regions(void) { return opc::R_COUNT; }
int opc_ops(void) { return opc::O_COUNT; }
}
