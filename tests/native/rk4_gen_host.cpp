// Host twin of the generated device code (quadruped_landing_b200/csrc/rk4_dual_gen.h), used by
// tests/test_rk4_gen.py to check -- on the CPU, bit for bit -- that the sparsity-exploiting
// straight-line code equals the oracle's dense 20-wide dual evaluation.
// Build: g++ -O2 -ffp-contract=off -fPIC -shared (done by the test).
#include <cstring>

#define QL_ADD(a, b) ((a) + (b))
#define QL_SUB(a, b) ((a) - (b))
#define QL_MUL(a, b) ((a) * (b))
#define QL_DIV(a, b) ((a) / (b))
#define QL_FN static inline
#define QL_ST(ptr, off, val) ((ptr)[(off)] = (val))
#include "../../quadruped_landing_b200/csrc/rk4_dual_gen.h"

extern "C" {

// xn[15], J[15*20] column-major (J[i + 15*j]); entries outside the pattern are set to 0.
void host_rk4_jac(int mode, const double* x, const double* u, double g, double mb, double mf, double lb,
                  double* xn, double* J)
{
    const double Ib = mb * (lb * lb) / 12;
    double jv[QL_NJ_MODE1 > QL_NJ_MODE3 ? QL_NJ_MODE1 : QL_NJ_MODE3];
    std::memset(J, 0, sizeof(double) * 300);
    if (mode == 1) {
        ql_rk4_jac_mode1(x, u, g, mb, mf, Ib, xn, jv);
        for (int n = 0; n < QL_NJ_MODE1; ++n) J[QL_PAT_I_MODE1[n] + 15 * QL_PAT_J_MODE1[n]] = jv[n];
    } else if (mode == 2) {
        ql_rk4_jac_mode2(x, u, g, mb, mf, Ib, xn, jv);
        for (int n = 0; n < QL_NJ_MODE2; ++n) J[QL_PAT_I_MODE2[n] + 15 * QL_PAT_J_MODE2[n]] = jv[n];
    } else {
        ql_rk4_jac_mode3(x, u, g, mb, mf, Ib, xn, jv);
        for (int n = 0; n < QL_NJ_MODE3; ++n) J[QL_PAT_I_MODE3[n] + 15 * QL_PAT_J_MODE3[n]] = jv[n];
    }
}

void host_rk4(int mode, const double* x, const double* u, double g, double mb, double mf, double lb, double* xn)
{
    const double Ib = mb * (lb * lb) / 12;
    if (mode == 1) ql_rk4_mode1(x, u, g, mb, mf, Ib, xn);
    else if (mode == 2) ql_rk4_mode2(x, u, g, mb, mf, Ib, xn);
    else ql_rk4_mode3(x, u, g, mb, mf, Ib, xn);
}

// Patch a run image the way the kernel does: p[grp] = run + shift[grp].
void host_patch(int mode, const double* jv, double* run, const int* shift, int jump)
{
    double* p[7];
    for (int i = 0; i < 7; ++i) p[i] = run + shift[i];
    if (mode == 1) ql_patch_mode1(jv, p, jump != 0);
    else if (mode == 2) ql_patch_mode2(jv, p, jump != 0);
    else ql_patch_mode3(jv, p, jump != 0);
}
}
