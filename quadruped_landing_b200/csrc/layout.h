// layout.h -- closed-form index maps of the planar-quadruped landing NLP, shared by the host
// planner and the CUDA kernels.  Knot numbers `k` are 1-BASED like the reference; everything
// returned is a 0-BASED offset.
//
//   Z layout        src/nlp.jl:38-39      knot k: x at 20(k-1)..+14, u at 20(k-1)+15..+19
//   g layout        src/nlp.jl:48-63      init | term | dyn | contact-first | contact-other | final-ctrl | body-pos
//   J value order   src/moi.jl:31-33      column-major (row fastest) filter of what jac_c!
//                                         (src/constraints.jl:212-291) assigns  == SPARSE_BLOCK
//
// Because columns belong to knots, the SPARSE_BLOCK value stream is the concatenation of one
// "run" per knot.  Inside knot k's run, state column j (0..14) holds, in row order,
//   [init 15 rows if k==1 | term 14 rows if k==N] [-I 15 rows if k>=2] [RK4 15 rows if k<N]
//   [contact-first row] [contact-other row] [body-pos row]           (the "extras", <=1 per column)
// and control column j (15..19, k<N only) holds [RK4 15 rows] [final-ctrl row if k==N-1, j in {16,18}].
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define QL_HD __host__ __device__ __forceinline__
#define QL_UNROLL _Pragma("unroll")
#else
#define QL_HD static inline
#define QL_UNROLL
#endif

#define QL_NX 15
#define QL_NU 5
#define QL_NZK 20
#define QL_LANES 32           // knots per pass: one lane per knot
#define QL_JBUF 1064          // doubles per J staging buffer: two runs (<= 529 + 531) + parity, rounded to 16 B
#define QL_MAX_N 1024
#define QL_NCOST 41           // cost fields per knot: Q[15] q[15] R[5] r[5] c

struct QlClass {
    int N, k_trans, init_mode;            // 1-based meaning, as in HybridNLP (nlp.jl:16-19)
    int n_nlp, m_nlp, nnz;                // nlp.jl:72,63 ; entries jac_c! assigns (SPARSE_BLOCK)
    int nnz_true;                         // structurally non-zero entries (SPARSE_TRUE)
    int nnz_vals;                         // value-dependent entries (VALS stream of the host path)
    int c_term, c_dyn, c_cfirst, c_cother, c_fctrl, c_body;   // 0-based first row of each g block (c_init = 0)
    int npass;                            // ceil(N / 32)
    double g, mb, mf, lb;                 // planar_quadruped.jl:11-20
    double Ib;                            // mb * lb^2 / 12   (planar_quadruped.jl:41)
    double mbg;                           // mb * g           (constraints.jl:154)
    double half_lb;                       // lb / 2           (constraints.jl:109,270-272)
};

QL_HD void ql_class_init(QlClass* c, int N, int k_trans, int init_mode,
                         double g, double mb, double mf, double lb)
{
    c->N = N; c->k_trans = k_trans; c->init_mode = init_mode;
    c->n_nlp = QL_NX * N + QL_NU * (N - 1);
    c->c_term = QL_NX;
    c->c_dyn = c->c_term + (QL_NX - 1);
    c->c_cfirst = c->c_dyn + QL_NX * (N - 1);
    c->c_cother = c->c_cfirst + N;
    c->c_fctrl = c->c_cother + (N - k_trans + 1);
    c->c_body = c->c_fctrl + 1;
    c->m_nlp = c->c_body + N;
    c->nnz = 529 * N - k_trans - 87;
    c->npass = (N + QL_LANES - 1) / QL_LANES;
    c->g = g; c->mb = mb; c->mf = mf; c->lb = lb;
    c->Ib = mb * (lb * lb) / 12;
    c->mbg = mb * g;
    c->half_lb = lb / 2;
    c->nnz_true = 0;                      // filled by ql_class_finish (needs the helpers below)
    c->nnz_vals = 0;
}

// ---- per-knot extras -----------------------------------------------------------------------
// e4: column 4 (y1) carries an extra row: contact-first (init_mode 1) or contact-other (init_mode 2, k>=k_trans)
// e6: column 6 (y2) likewise with the roles swapped          (constraints.jl:235-256)
QL_HD int ql_e4(const QlClass& c, int k) { return (c.init_mode == 1) ? 1 : (k >= c.k_trans ? 1 : 0); }
QL_HD int ql_e6(const QlClass& c, int k) { return (c.init_mode == 2) ? 1 : (k >= c.k_trans ? 1 : 0); }
QL_HD int ql_fc(const QlClass& c, int k) { return k == c.N - 1 ? 1 : 0; }     // constraints.jl:259-260

// offset of knot k's run in the value stream
QL_HD int ql_run_off(const QlClass& c, int k)
{
    int o = 528 * (k - 1);
    if (k > c.k_trans) o += k - c.k_trans;       // one contact-other entry per earlier knot >= k_trans
    if (k == c.N) o += 2;                        // the two final-ctrl entries of knot N-1
    return o;
}
QL_HD int ql_run_len(const QlClass& c, int k)
{
    return (k == c.N ? c.nnz : ql_run_off(c, k + 1)) - ql_run_off(c, k);
}
// rows per state column before the extras
QL_HD int ql_col_width(const QlClass& c, int k) { return k == c.N ? 29 : 30; }

// shift (number of extras) preceding column group grp of knot k; groups:
//   0: cols 0-1   1: col 2   2: cols 3-4   3: cols 5-6   4: cols 7-16   5: cols 17-18   6: col 19
QL_HD int ql_col_group(int j)
{
    return j <= 1 ? 0 : j == 2 ? 1 : j <= 4 ? 2 : j <= 6 ? 3 : j <= 16 ? 4 : j <= 18 ? 5 : 6;
}
QL_HD int ql_group_shift(const QlClass& c, int k, int grp)
{
    const int e4 = ql_e4(c, k), e6 = ql_e6(c, k), fc = ql_fc(c, k);
    switch (grp) {
    case 0: return 0;
    case 1: return 1;
    case 2: return 2;
    case 3: return 2 + e4;
    case 4: return 2 + e4 + e6;
    case 5: return 2 + e4 + e6 + fc;
    default: return 2 + e4 + e6 + 2 * fc;
    }
}
// start of column j inside knot k's run
QL_HD int ql_col_start(const QlClass& c, int k, int j)
{
    const int cw = ql_col_width(c, k);
    const int sh = ql_group_shift(c, k, ql_col_group(j));
    return j < QL_NX ? cw * j + sh : cw * QL_NX + QL_NX * (j - QL_NX) + sh;
}
// offset of RK4-block entry (i, j) inside the run of a knot k < N (what rk4_dual_gen.h's patch code uses)
QL_HD int ql_rk4_pos(const QlClass& c, int k, int i, int j) { return ql_col_start(c, k, j) + (j < QL_NX ? 15 : 0) + i; }
// offset of the body-pos d/dtheta entry (the only other value-dependent entry) inside knot k's run
QL_HD int ql_theta_pos(const QlClass& c, int k) { return ql_col_start(c, k, 2) + ql_col_width(c, k); }

// Constants of knot k's run (everything jac_c! assigns that does not depend on Z), into a zeroed image.
// Split by "slot" so that 16 lanes can write one knot's constants in parallel: slot j < 15 = the identity-block
// entries of state column j, slot 15 = the unit entries of the extra rows.
QL_HD void ql_write_run_constants_slot(const QlClass& c, int k, double* run, int j)
{
    const int cw = ql_col_width(c, k);
    const int e4 = ql_e4(c, k), e6 = ql_e6(c, k);
    if (j < QL_NX) {
        const int cs = cw * j + (j >= 2) + (j >= 3) + (j >= 5 ? e4 : 0) + (j >= 7 ? e6 : 0);
        if (k == 1) run[cs + j] = 1.0;                         // jac_init .= I(n)            constraints.jl:228
        if (k == c.N && j < QL_NX - 1) run[cs + j] = 1.0;      // jac_term .= I(n)[1:n-1,:]   constraints.jl:229
        if (k >= 2) run[cs + (k == c.N ? QL_NX - 1 : 0) + j] = -1.0;   // D[ci, xi[k+1]] .= -I(n)   :200
        return;
    }
    run[cw * 1 + cw] = 1.0;                                    // body-pos d/dyb              constraints.jl:267
    if (e4) run[cw * 4 + 2 + cw] = 1.0;                        // contact row on y1           :237 / :254
    if (e6) run[cw * 6 + 2 + e4 + cw] = 1.0;                   // contact row on y2           :241 / :249
    if (k == c.N - 1) {                                        // final-ctrl row              :259-260
        const int cb = cw * QL_NX + 2 + e4 + e6;
        run[cb + 15 * 1 + 15] = 1.0;                           // after the RK4 rows of control column 16 (F1y)
        run[cb + 15 * 3 + 1 + 15] = 1.0;                       // ... and of control column 18 (F2y)
    }
}
QL_HD void ql_write_run_constants(const QlClass& c, int k, double* run)
{
    for (int j = 0; j <= QL_NX; ++j) ql_write_run_constants_slot(c, k, run, j);
}

// ---- SPARSE_TRUE: only structurally non-zero entries (identity blocks as diagonals, RK4 blocks as their
// mode-specific pattern: 71 / 71 / 57 entries for modes 1 / 2 / 3, 56 at the jump knot) ------------------
// run length of a knot without its extras; the per-mode counts come from rk4_dual_gen.h (QL_TRUE_LEN_*):
//   86 = 15 diagonal + 71 pattern (initial mode), 71 = 15 + 56 (jump knot), 72 = 15 + 57 (mode 3)
#define QL_TRUE_LEN_INIT 86
#define QL_TRUE_LEN_JUMPK 71
#define QL_TRUE_LEN_M3 72
#define QL_TRUE_LEN_LAST 29      // knot N: 14 term-diagonal + 15 (-I)-diagonal entries
#define QL_TRUE_PBUF 1476        // doubles of the half-pass staging buffer: 16 x (86 + 6) + parity, rounded to 16 B

QL_HD int ql_true_base_len(const QlClass& c, int k)
{
    if (k == c.N) return QL_TRUE_LEN_LAST;
    if (k >= c.k_trans) return QL_TRUE_LEN_M3;
    return k == c.k_trans - 1 ? QL_TRUE_LEN_JUMPK : QL_TRUE_LEN_INIT;
}
// offset of knot k's run in the SPARSE_TRUE value stream
QL_HD int ql_true_run_off(const QlClass& c, int k)
{
    const int km = k - 1;                                                // knots before k
    const int kj = c.k_trans - 1;                                        // the jump knot (0: none)
    int n_init = kj - 1; if (n_init < 0) n_init = 0; if (n_init > km) n_init = km;
    const int n_jump = (kj >= 1 && kj <= km) ? 1 : 0;
    int hi = km < c.N - 1 ? km : c.N - 1;
    int n_m3 = hi - c.k_trans + 1; if (n_m3 < 0) n_m3 = 0;
    int o = n_init * QL_TRUE_LEN_INIT + n_jump * QL_TRUE_LEN_JUMPK + n_m3 * QL_TRUE_LEN_M3;
    o += 3 * km;                                                         // body-pos x2 + contact-first per knot
    if (k > c.k_trans) o += k - c.k_trans;                               // contact-other entries
    if (k == c.N) o += 2;                                                // final-ctrl entries of knot N-1
    return o;
}
QL_HD int ql_true_nnz(const QlClass& c) { return ql_true_run_off(c, c.N) + QL_TRUE_LEN_LAST + 2 + ql_e4(c, c.N) + ql_e6(c, c.N); }


// ---- VALS: only the VALUE-DEPENDENT entries, column-major order (what changes between two evaluations of a
// SPARSE_BLOCK / SPARSE_TRUE row: the jv entries of the RK4 block + the body-clearance d/dtheta entry).  This is
// the stream host-pointer batches ship over PCIe; host threads merge it into the caller's rows (hostrows.cpp).
// Per-knot counts come from rk4_dual_gen.h (QL_VALS_LEN_*; checked by static_assert in qlnlp_kernels.cuh).
#define QL_VALS_LEN_INIT 56      // 55 jv + d/dtheta
#define QL_VALS_LEN_JUMPK 49     // jump knot: the 7 entries of masked rows are constant zeros
#define QL_VALS_LEN_M3 42        // 41 jv + d/dtheta
#define QL_VALS_LEN_LAST 1       // knot N: d/dtheta only
QL_HD int ql_vals_len(const QlClass& c, int k)
{
    if (k == c.N) return QL_VALS_LEN_LAST;
    if (k >= c.k_trans) return QL_VALS_LEN_M3;
    return k == c.k_trans - 1 ? QL_VALS_LEN_JUMPK : QL_VALS_LEN_INIT;
}
QL_HD int ql_vals_run_off(const QlClass& c, int k)
{
    const int km = k - 1;
    const int kj = c.k_trans - 1;
    int n_init = kj - 1; if (n_init < 0) n_init = 0; if (n_init > km) n_init = km;
    const int n_jump = (kj >= 1 && kj <= km) ? 1 : 0;
    int hi = km < c.N - 1 ? km : c.N - 1;
    int n_m3 = hi - c.k_trans + 1; if (n_m3 < 0) n_m3 = 0;
    return n_init * QL_VALS_LEN_INIT + n_jump * QL_VALS_LEN_JUMPK + n_m3 * QL_VALS_LEN_M3;
}
QL_HD int ql_vals_nnz(const QlClass& c) { return ql_vals_run_off(c, c.N) + QL_VALS_LEN_LAST; }

// ---- Lagrangian Hessian (no reference counterpart): block diagonal, one 20x20 block per knot
// (15x15 for the last); values = lower triangle, column-major, restricted to the structural pattern of the knot's
// mode (rk4_dual_gen.h: QL_HESS_R/C_MODE*; checked by static_assert in qlnlp_hess.cuh).
#define QL_HESS_LEN_INIT 57      // modes 1 / 2 (the jump knot keeps the pattern; masked rows contribute zeros)
#define QL_HESS_LEN_M3 55
#define QL_HESS_LEN_LAST 15      // knot N: the diagonal (terminal cost + body-clearance d2/dtheta2)
#define QL_HESS_PBUF 1832        // doubles of the per-pass staging buffer: 32 x 57 + parity, rounded to 16 B
QL_HD int ql_hess_len(const QlClass& c, int k)
{
    if (k == c.N) return QL_HESS_LEN_LAST;
    return k >= c.k_trans ? QL_HESS_LEN_M3 : QL_HESS_LEN_INIT;
}
QL_HD int ql_hess_run_off(const QlClass& c, int k)
{
    const int km = k - 1;                                         // knots before k
    int n_init = c.k_trans - 1; if (n_init > km) n_init = km;     // ... of them in the initial mode
    int hi = km < c.N - 1 ? km : c.N - 1;
    int n_m3 = hi - (c.k_trans - 1); if (n_m3 < 0) n_m3 = 0;
    return n_init * QL_HESS_LEN_INIT + n_m3 * QL_HESS_LEN_M3;
}
QL_HD int ql_hess_nnz(const QlClass& c) { return ql_hess_run_off(c, c.N) + QL_HESS_LEN_LAST; }

QL_HD void ql_class_finish(QlClass* c) { c->nnz_true = ql_true_nnz(*c); c->nnz_vals = ql_vals_nnz(*c); }

// ---- segments: what one bulk store moves ---------------------------------------------------
// A segment is 1 or 2 consecutive knots of one pass.  Its image lives in a staging buffer at offset
// (start & 1) so that shared and global addresses agree modulo 16 B.  16-byte records: the kernel
// keeps the plan in shared memory and reads a segment with one 128-bit load.
struct QlSeg {
    int start;            // stream offset of knot k0's run
    int end;              // stream offset one past the last knot's run
    short k0;             // first knot (1-based)
    signed char nk;       // 1 or 2
    signed char buf;      // unused (the kernel alternates buffers with a running counter, ql_seg_buffer)
    short tmpl;           // template id: equal ids <=> identical constant image
    short pad;
};
// staging buffer of the n-th segment a warp streams (n counts across evaluations): strict alternation, so the two
// most recent bulk stores always come from different buffers whatever the number of segments per evaluation
QL_HD int ql_seg_buffer(unsigned n) { return (int)(n & 1u); }
