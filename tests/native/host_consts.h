// Host-side definitions the generated header (csrc/rk4_dual_gen.h) expects from its includer.
// Plain IEEE operations; compile with -ffp-contract=off.  TEST HARNESS ONLY.
#pragma once
#define QL_ADD(a, b) ((a) + (b))
#define QL_SUB(a, b) ((a) - (b))
#define QL_MUL(a, b) ((a) * (b))
#define QL_DIV_MB(a) ((a) / K.mb)
#define QL_DIV_MF(a) ((a) / K.mf)
#define QL_DIV_IB(a) ((a) / K.Ib)
#define QL_DIV_SIX(a) ((a) / 6.0)
#define QL_FN static inline
#define QL_ST(ptr, off, val) ((ptr)[(off)] = (val))
#define QL_PADD(ptr, n) ((ptr) + (n))
struct HostConsts {
    double g, mb, mf, Ib;
};
