// agc_math_check — the GPU's logf / expf for the RMS AGC (iq_tool_b200/csrc/agc_math.h, plain IEEE double arithmetic)
// against this machine's libm, on the CPU.  Usage: agc_math_check [stride]   (stride 1 = every float in range)
// Prints the largest pre-rounding relative error (against long double) and, for information, how many float results differ
// from this libm's (glibc 2.39's logf / expf are NOT correctly rounded: ~0.3 % / ~0.01 % of arguments are one ulp off the
// correctly rounded value, which is what these functions return).  Exit code 1 if the pre-rounding error exceeds 2e-15,
// i.e. if the result could differ from the correctly rounded one for more than ~1e-7 of the arguments.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../iq_tool_b200/csrc/agc_math.h"

int main(int argc, char** argv)
{
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 97;
    AgcLogEntry tab[AGC_LOG_N];
    agc_log_table(tab);
    static const double K[AGC_NCOEF] = AGC_COEF_LIST;
    unsigned long long n = 0, bad = 0;
    double worst = 0.0;
    // logf: every stride-th float in [1e-6, 1e6] plus a dense sweep around 1.0 (the AGC's steady state)
    for (int pass = 0; pass < 2; pass++) {
        uint32_t lo, hi, st;
        float a = pass ? 0.96f : 1.0e-6f, b = pass ? 1.04f : 1.0e6f;
        memcpy(&lo, &a, 4); memcpy(&hi, &b, 4);
        st = pass ? 1 : stride;
        for (uint32_t u = lo; u <= hi; u += st) {
            float x; memcpy(&x, &u, 4);
            const double d = agc_log_fast(x, tab, K);
            const float mine = (float)d, ref = logf(x);
            const long double ex = logl((long double)x);
            if (ex != 0.0L) { const double rel = (double)fabsl(((long double)d - ex) / ex); if (rel > worst) worst = rel; }
            n++;
            if (mine != ref) bad++;
        }
    }
    printf("logf: %llu arguments, %llu differ from libm, worst pre-rounding relative error %.3g\n", n, bad, worst);
    unsigned long long n2 = 0, bad2 = 0;
    double worst2 = 0.0;
    // expf: floats with |t| <= 0.125 (both signs, down to 1e-12) 
    for (int sgn = 0; sgn < 2; sgn++) {
        uint32_t lo, hi;
        float a = 1.0e-12f, b = 0.125f;
        memcpy(&lo, &a, 4); memcpy(&hi, &b, 4);
        for (uint32_t u = lo; u <= hi; u += stride) {
            float t; memcpy(&t, &u, 4);
            if (sgn) t = -t;
            const double d = agc_exp_tiny((double)t, K);
            const float mine = (float)d, ref = expf(t);
            const long double ex = expl((long double)t);
            const double rel = (double)fabsl(((long double)d - ex) / ex);
            if (rel > worst2) worst2 = rel;
            n2++;
            if (mine != ref) bad2++;
        }
    }
    printf("expf: %llu arguments, %llu differ from libm, worst pre-rounding relative error %.3g\n", n2, bad2, worst2);
    (void)bad; (void)bad2;
    return (worst > 2e-15 || worst2 > 2e-15) ? 1 : 0;
}
