// Host twin of the generated device code (quadruped_landing_b200/csrc/rk4_dual_gen.h), used by
// tests/test_rk4_gen.py to check -- on the CPU, bit for bit -- that the sparsity-exploiting
// straight-line code equals the oracle's dense 20-wide dual evaluation.
// Build: g++ -O2 -ffp-contract=off -fPIC -shared (done by the test).
#include <cstring>

#include <cmath>
#include "host_consts.h"
#include "../../quadruped_landing_b200/csrc/rk4_dual_gen.h"

extern "C" {

// xn[15], J[15*20] column-major (J[i + 15*j]); entries outside the pattern are set to 0.
void host_rk4_jac(int mode, const double* x, const double* u, double g, double mb, double mf, double lb,
                  double* xn, double* J)
{
    const HostConsts K{g, mb, mf, mb * (lb * lb) / 12};
    double jv[QL_NJ_MODE1 > QL_NJ_MODE3 ? QL_NJ_MODE1 : QL_NJ_MODE3];
    std::memset(J, 0, sizeof(double) * 300);
#define SCATTER(M)                                                                                         \
    ql_rk4_jac_mode##M(x, u, K, xn, jv);                                                                   \
    for (int n = 0; n < QL_NJ_MODE##M; ++n) J[QL_PAT_I_MODE##M[n] + 15 * QL_PAT_J_MODE##M[n]] = jv[n];       \
    for (int n = 0; n < QL_NJC_MODE##M; ++n) J[QL_CPAT_I_MODE##M[n] + 15 * QL_CPAT_J_MODE##M[n]] = QL_CPAT_V_MODE##M[n];
    if (mode == 1) { SCATTER(1) } else if (mode == 2) { SCATTER(2) } else { SCATTER(3) }
#undef SCATTER
}

void host_rk4(int mode, const double* x, const double* u, double g, double mb, double mf, double lb, double* xn)
{
    const HostConsts K{g, mb, mf, mb * (lb * lb) / 12};
    if (mode == 1) ql_rk4_mode1(x, u, K, xn);
    else if (mode == 2) ql_rk4_mode2(x, u, K, xn);
    else ql_rk4_mode3(x, u, K, xn);
}

// The kernel divides by the model constants with a precomputed reciprocal and two FMAs
// (q = a*r; e = fma(-q, b, a); q' = fma(e, r, q)).  Counts how many of n pseudo-random numerators give a
// result different from the IEEE quotient a / b (expected: 0).
long long host_count_fastdiv_mismatches(double b, long long n, unsigned long long seed)
{
    const double r = 1.0 / b;
    long long bad = 0;
    unsigned long long s = seed ? seed : 88172645463325252ULL;
    for (long long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        // random sign, mantissa and an exponent in [-40, 40]
        const double m = 1.0 + (double)(s >> 12) * (1.0 / 4503599627370496.0);
        const int e = (int)((s >> 3) % 81) - 40;
        const double a = std::ldexp((s & 1) ? -m : m, e);
        const double q = a * r;
        const double rem = std::fma(-q, b, a);
        const double q1 = std::fma(rem, r, q);
        if (q1 != a / b) ++bad;
    }
    return bad;
}
}
