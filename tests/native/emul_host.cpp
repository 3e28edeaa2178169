// CPU emulation of the kernel's Jacobian value stream, built from the SAME headers the kernel
// uses (layout.h: run offsets, constants, positions; rk4_dual_gen.h: RK4 duals + patch code),
// following the kernel's segment plan: zero image -> run constants -> patch -> copy [start,end).
// tests/test_layout_cpu.py compares the result with the oracle entry by entry, which checks all
// of the kernel's index arithmetic without a GPU.  TEST HARNESS ONLY.
#include <cmath>
#include <cstring>
#include <vector>

#include "host_consts.h"
#include "../../quadruped_landing_b200/csrc/layout.h"
#include "../../quadruped_landing_b200/csrc/rk4_dual_gen.h"
#include "../../quadruped_landing_b200/csrc/true_run.h"

extern "C" {

// segs: [nseg][6] = k0 nk start end tmpl buf (from qlnlp_debug_segments).  Returns 0, or a
// negative code if the plan itself is inconsistent.
int emul_jac_stream(int N, int k_trans, int init_mode, double g, double mb, double mf, double lb,
                    const long long* segs, int nseg, const double* Z, double* out, int persist_templates)
{
    QlClass c;
    ql_class_init(&c, N, k_trans, init_mode, g, mb, mf, lb);
    std::vector<double> bufs[2] = {std::vector<double>(QL_JBUF, NAN), std::vector<double>(QL_JBUF, NAN)};
    int tmpl[2] = {-1, -1};
    unsigned segctr = 0;          // the kernel alternates the staging buffers with a running counter across evaluations
    int last_buf = -1;
    for (int rep = 0; rep < (persist_templates ? 3 : 1); ++rep) {   // later repetitions reuse the images
        for (int i = 0; i < c.nnz; ++i) out[i] = NAN;
        for (int s = 0; s < nseg; ++s) {
            const long long* sg = segs + 6 * s;
            const int k0 = (int)sg[0], nk = (int)sg[1], start = (int)sg[2], end = (int)sg[3], tm = (int)sg[4];
            const int bi = ql_seg_buffer(segctr++);
            if (bi < 0 || bi > 1 || nk < 1 || nk > 2) return -1;
            // the bulk store committed just before may still be reading ITS buffer (wait_group.read 1): the buffer
            // being rewritten must be the other one, also across the wrap from one evaluation to the next
            if (bi == last_buf) return -3;
            last_buf = bi;
            const int base = start & ~1;
            if (end - base > QL_JBUF) return -2;
            double* buf = bufs[bi].data();
            const HostConsts K{c.g, c.mb, c.mf, c.Ib};
            if (tmpl[bi] != tm) {
                std::memset(buf, 0, sizeof(double) * QL_JBUF);
                for (int k = k0; k < k0 + nk; ++k) {
                    double* run = buf + (ql_run_off(c, k) - base);
                    ql_write_run_constants(c, k, run);
                    if (k < N) {
                        double* p[7];
                        for (int grp = 0; grp < 7; ++grp) p[grp] = run + ql_group_shift(c, k, grp);
                        const bool jump = (k == k_trans - 1);
                        if (k >= k_trans) ql_const_mode3(p, jump);
                        else if (init_mode == 1) ql_const_mode1(p, jump);
                        else ql_const_mode2(p, jump);
                    }
                }
                tmpl[bi] = tm;
            }
            for (int k = k0; k < k0 + nk; ++k) {
                double* run = buf + (ql_run_off(c, k) - base);
                const double* x = Z + 20 * (k - 1);
                if (k < N) {
                    const double* u = x + 15;
                    double xn[15], jv[QL_NJ_MODE1];
                    double* p[7];
                    for (int grp = 0; grp < 7; ++grp) p[grp] = run + ql_group_shift(c, k, grp);
                    const bool jump = (k == k_trans - 1);
                    if (k >= k_trans) { ql_rk4_jac_mode3(x, u, K, xn, jv); ql_patch_mode3(jv, p, jump); }
                    else if (init_mode == 1) { ql_rk4_jac_mode1(x, u, K, xn, jv); ql_patch_mode1(jv, p, jump); }
                    else { ql_rk4_jac_mode2(x, u, K, xn, jv); ql_patch_mode2(jv, p, jump); }
                }
                const double th = x[2];
                run[ql_theta_pos(c, k)] = (th > 0) ? (-c.half_lb) * std::cos(th) : c.half_lb * std::cos(th);
            }
            for (int i = start; i < end; ++i) out[i] = buf[i - base];
        }
    }
    return 0;
}

// SPARSE_TRUE stream exactly as the kernel assembles it: every knot writes its whole run at ql_true_run_off.
int emul_true_stream(int N, int k_trans, int init_mode, double g, double mb, double mf, double lb,
                     const double* Z, double* out, int nout)
{
    QlClass c;
    ql_class_init(&c, N, k_trans, init_mode, g, mb, mf, lb);
    ql_class_finish(&c);
    if (nout != c.nnz_true) return -1;
    const HostConsts K{c.g, c.mb, c.mf, c.Ib};
    for (int i = 0; i < nout; ++i) out[i] = NAN;
    for (int k = 1; k <= N; ++k) {
        const double* x = Z + 20 * (k - 1);
        double xn[15], jv[QL_NJ_MODE1];
        if (k < N) {
            const double* u = x + 15;
            if (k >= k_trans) ql_rk4_jac_mode3(x, u, K, xn, jv);
            else if (init_mode == 1) ql_rk4_jac_mode1(x, u, K, xn, jv);
            else ql_rk4_jac_mode2(x, u, K, xn, jv);
        }
        const double th = x[2];
        const double jtheta = (th > 0) ? (-c.half_lb) * std::cos(th) : c.half_lb * std::cos(th);
        const int off = ql_true_run_off(c, k);
        const int len = (k == N ? c.nnz_true : ql_true_run_off(c, k + 1)) - off;
        if (off < 0 || off + len > nout) return -2;
        ql_true_write_run(c, k, jv, jtheta, out + off);
    }
    return 0;
}

// VALS stream exactly as the kernel assembles it: every knot writes its value-dependent entries at ql_vals_run_off.
int emul_vals_stream(int N, int k_trans, int init_mode, double g, double mb, double mf, double lb,
                     const double* Z, double* out, int nout)
{
    QlClass c;
    ql_class_init(&c, N, k_trans, init_mode, g, mb, mf, lb);
    ql_class_finish(&c);
    if (nout != c.nnz_vals) return -1;
    const HostConsts K{c.g, c.mb, c.mf, c.Ib};
    for (int i = 0; i < nout; ++i) out[i] = NAN;
    for (int k = 1; k <= N; ++k) {
        const double* x = Z + 20 * (k - 1);
        double xn[15], jv[QL_NJ_MODE1];
        if (k < N) {
            const double* u = x + 15;
            if (k >= k_trans) ql_rk4_jac_mode3(x, u, K, xn, jv);
            else if (init_mode == 1) ql_rk4_jac_mode1(x, u, K, xn, jv);
            else ql_rk4_jac_mode2(x, u, K, xn, jv);
        }
        const double th = x[2];
        const double jtheta = (th > 0) ? (-c.half_lb) * std::cos(th) : c.half_lb * std::cos(th);
        const int off = ql_vals_run_off(c, k);
        if (off < 0 || off + ql_vals_len(c, k) > nout) return -2;
        if (k < N && off + ql_vals_len(c, k) != ql_vals_run_off(c, k + 1)) return -3;
        ql_vals_write_run(c, k, jv, jtheta, out + off);
    }
    return 0;
}

// Lagrangian-Hessian stream exactly as the kernel assembles it (qlnlp_hess.cuh): every knot writes its block at
// ql_hess_run_off with the generated second-order code.  cost: knot-major Q[N][15] R[N][5] q[N][15] r[N][5].
int emul_hess_stream(int N, int k_trans, int init_mode, double g, double mb, double mf, double lb,
                     const double* Q, const double* R, const double* q, const double* r,
                     const double* Z, double sigma, const double* lambda, double* out, int nout)
{
    QlClass c;
    ql_class_init(&c, N, k_trans, init_mode, g, mb, mf, lb);
    ql_class_finish(&c);
    if (nout != ql_hess_nnz(c)) return -1;
    const HostConsts K{c.g, c.mb, c.mf, c.Ib};
    for (int i = 0; i < nout; ++i) out[i] = NAN;
    for (int k = 1; k <= N; ++k) {
        const double* x = Z + 20 * (k - 1);
        const double* u = x + 15;
        const int off = ql_hess_run_off(c, k);
        if (off < 0 || off + ql_hess_len(c, k) > nout) return -2;
        if (k < N && off + ql_hess_len(c, k) != ql_hess_run_off(c, k + 1)) return -3;
        const double as = c.half_lb * std::sin(x[2]);
        const double tt = lambda[c.c_body + (k - 1)] * ((x[2] > 0) ? as : -as);
        const double* Qk = Q + 15 * (k - 1);
        if (k == N) {
            for (int i = 0; i < 15; ++i) out[off + i] = (i == 2) ? sigma * Qk[i] + tt : sigma * Qk[i];
            continue;
        }
        double lam[15], hv[QL_NH_MODE1], ox[15], ou[4], hx[15], hu[4];
        for (int i = 0; i < 15; ++i) lam[i] = lambda[c.c_dyn + 15 * (k - 1) + i];
        if (k == k_trans - 1) { lam[4] = lam[6] = lam[10] = lam[11] = lam[12] = lam[13] = 0.0; }
        const double h = u[4];
        const double *Rk = R + 5 * (k - 1), *qk = q + 15 * (k - 1), *rk = r + 5 * (k - 1);
        for (int i = 0; i < 15; ++i) { ox[i] = sigma * (h * Qk[i]); hx[i] = sigma * (Qk[i] * x[i] + qk[i]); }
        for (int i = 0; i < 4; ++i) { ou[i] = sigma * (h * Rk[i]); hu[i] = sigma * (Rk[i] * u[i] + rk[i]); }
        const double hh = sigma * (2.0 * (Rk[4] * h + rk[4]) + h * Rk[4]);
        if (k >= k_trans) { ql_rk4_hess_mode3(x, u, K, lam, hv); ql_hess_store_mode3(hv, ox, ou, hx, hu, hh, tt, out + off); }
        else if (init_mode == 1) { ql_rk4_hess_mode1(x, u, K, lam, hv); ql_hess_store_mode1(hv, ox, ou, hx, hu, hh, tt, out + off); }
        else { ql_rk4_hess_mode2(x, u, K, lam, hv); ql_hess_store_mode2(hv, ox, ou, hx, hu, hh, tt, out + off); }
    }
    return 0;
}

// closed-form helpers exposed for direct checks
int emul_run_off(int N, int k_trans, int init_mode, int k)
{
    QlClass c;
    ql_class_init(&c, N, k_trans, init_mode, -9.81, 10, 0.1, 0.5);
    return ql_run_off(c, k);
}
int emul_rk4_pos(int N, int k_trans, int init_mode, int k, int i, int j)
{
    QlClass c;
    ql_class_init(&c, N, k_trans, init_mode, -9.81, 10, 0.1, 0.5);
    return ql_rk4_pos(c, k, i, j);
}
}
