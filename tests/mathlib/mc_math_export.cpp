/* test harness: exports include/mc_math.h functions over arrays for ulp comparisons against the host libm */
#include "../../include/mc_math.h"
extern "C" void mcm_apply(int fn, const double* x, const double* y, double* out, long n)
{
    for (long i = 0; i < n; i++) {
        switch (fn) {
        case 0: out[i] = mcm::exp(x[i]); break;
        case 1: out[i] = mcm::log(x[i]); break;
        case 2: out[i] = mcm::log10(x[i]); break;
        case 3: out[i] = mcm::exp10(x[i]); break;
        case 4: out[i] = mcm::pow(x[i], y[i]); break;
        case 5: out[i] = mcm::atan(x[i]); break;
        case 6: out[i] = mcm::cos(x[i]); break;
        case 7: out[i] = mcm::exp2(x[i]); break;
        case 8: out[i] = mcm::sin_r(x[i]); break;
        case 9: out[i] = mcm::cos_r(x[i]); break;
        case 10: out[i] = mcm::acos(x[i]); break;
        case 11: { double sn, cs; mcm::sincosd(x[i], &sn, &cs); out[i] = sn; break; }
        case 12: { double sn, cs; mcm::sincosd(x[i], &sn, &cs); out[i] = cs; break; }
        default: out[i] = 0.0;
        }
    }
}
