// true_run.h -- writes one knot's run of the SPARSE_TRUE value stream (layout.h: ql_true_run_off).
// Included after rk4_dual_gen.h by the kernel and by the CPU emulation harness; the includer defines
//   QL_ST(ptr, off, val)   store one double at ptr + off doubles
//   QL_PADD(ptr, n)        ptr advanced by n doubles
// PTR is `double*` on the host and a 32-bit shared-memory byte address on the device.
//
// Run of knot k < N, in the reference's column-major order restricted to structural non-zeros:
//   state column j : [diagonal entry: +1 init block (k == 1) | -1 (-I block, k >= 2)] [RK4 pattern rows of
//                    column j] [extra row: contact-first / contact-other / body-pos]
//   control column : [RK4 pattern rows] [final-ctrl row (k == N-1, columns 16 and 18)]
// Knot N: state column j = [term diagonal +1 (j < 14)] [-I diagonal -1] [extras].
#pragma once

template <typename PTR>
QL_FN void ql_true_write_run(const QlClass& c, int k, const double* jv, double jtheta, PTR run)
{
    const int e4 = ql_e4(c, k), e6 = ql_e6(c, k), fc = ql_fc(c, k);
    PTR p[7];
    p[0] = run;
    p[1] = QL_PADD(run, 1);
    p[2] = QL_PADD(run, 2);
    p[3] = QL_PADD(run, 2 + e4);
    p[4] = QL_PADD(run, 2 + e4 + e6);
    p[5] = QL_PADD(run, 2 + e4 + e6 + fc);
    p[6] = QL_PADD(run, 2 + e4 + e6 + 2 * fc);
    if (k == c.N) {
        QL_UNROLL
        for (int j = 0; j < QL_NX - 1; ++j) {
            QL_ST(p[ql_col_group(j)], 2 * j, 1.0);          // jac_term diagonal       constraints.jl:229
            QL_ST(p[ql_col_group(j)], 2 * j + 1, -1.0);     // -I diagonal             constraints.jl:200
        }
        QL_ST(p[4], 28, -1.0);                              // column 15 (time): no terminal row
        QL_ST(p[0], 4, 1.0);                                // body-pos d/dyb
        QL_ST(p[1], 6, jtheta);                             // body-pos d/dtheta
        if (e4) QL_ST(p[2], 10, 1.0);
        if (e6) QL_ST(p[3], 14, 1.0);
        return;
    }
    const double dg = (k == 1) ? 1.0 : -1.0;
#define QL_TRUE_EXTRAS(TAG)                                         \
    QL_ST(p[0], QL_TRUE_X1_##TAG, 1.0);                             \
    QL_ST(p[1], QL_TRUE_X2_##TAG, jtheta);                          \
    if (e4) QL_ST(p[2], QL_TRUE_X4_##TAG, 1.0);                     \
    if (e6) QL_ST(p[3], QL_TRUE_X6_##TAG, 1.0);                     \
    if (fc) { QL_ST(p[4], QL_TRUE_X16_##TAG, 1.0); QL_ST(p[5], QL_TRUE_X18_##TAG, 1.0); }
    if (k >= c.k_trans) {
        ql_store_true_mode3(jv, p, dg);
        QL_TRUE_EXTRAS(MODE3)
    } else if (k == c.k_trans - 1) {
        if (c.init_mode == 1) { ql_store_true_mode1_jump(jv, p, dg); QL_TRUE_EXTRAS(MODE1_JUMP) }
        else { ql_store_true_mode2_jump(jv, p, dg); QL_TRUE_EXTRAS(MODE2_JUMP) }
    } else {
        if (c.init_mode == 1) { ql_store_true_mode1(jv, p, dg); QL_TRUE_EXTRAS(MODE1) }
        else { ql_store_true_mode2(jv, p, dg); QL_TRUE_EXTRAS(MODE2) }
    }
#undef QL_TRUE_EXTRAS
}

// One knot's run of the VALS stream (layout.h: ql_vals_run_off): the value-dependent entries only, in column-major
// order -- jv entries (rows masked by the jump Jacobian dropped) with d/dtheta of the body-clearance row after column 2.
template <typename PTR>
QL_FN void ql_vals_write_run(const QlClass& c, int k, const double* jv, double jtheta, PTR run)
{
    if (k == c.N) { QL_ST(run, 0, jtheta); return; }
    if (k >= c.k_trans) ql_store_vals_mode3(jv, jtheta, run);
    else if (k == c.k_trans - 1) {
        if (c.init_mode == 1) ql_store_vals_mode1_jump(jv, jtheta, run);
        else ql_store_vals_mode2_jump(jv, jtheta, run);
    } else {
        if (c.init_mode == 1) ql_store_vals_mode1(jv, jtheta, run);
        else ql_store_vals_mode2(jv, jtheta, run);
    }
}

// host-side enumeration helper: pattern (i, j) lists per variant, used by the structure generator
struct QlTruePattern {
    int n;
    const unsigned char* I;
    const unsigned char* J;
};
static inline QlTruePattern ql_true_pattern(int mode, bool jump)
{
    if (mode == 3) return {QL_TRUE_NPAT_MODE3, QL_TRUE_I_MODE3, QL_TRUE_J_MODE3};
    if (mode == 1) return jump ? QlTruePattern{QL_TRUE_NPAT_MODE1_JUMP, QL_TRUE_I_MODE1_JUMP, QL_TRUE_J_MODE1_JUMP}
                               : QlTruePattern{QL_TRUE_NPAT_MODE1, QL_TRUE_I_MODE1, QL_TRUE_J_MODE1};
    return jump ? QlTruePattern{QL_TRUE_NPAT_MODE2_JUMP, QL_TRUE_I_MODE2_JUMP, QL_TRUE_J_MODE2_JUMP}
                : QlTruePattern{QL_TRUE_NPAT_MODE2, QL_TRUE_I_MODE2, QL_TRUE_J_MODE2};
}
