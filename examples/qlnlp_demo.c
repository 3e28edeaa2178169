/*
 * qlnlp_demo.c -- the C ABI (include/qlnlp.h) driven from plain C, the way a solver's C callbacks would.
 *
 * Builds the reference's default landing problem (src/main.ipynb cells 2-7: N=61, k_trans=21, init_mode=1),
 * evaluates the four MOI callbacks at the initial guess and prints the numbers the Ipopt log shows at iteration 0
 * (src/main.ipynb:221-232).  The struct `ipopt_callbacks` shows the adaptor to Ipopt's C interface signatures
 * (Eval_F_CB, Eval_Grad_F_CB, Eval_G_CB, Eval_Jac_G_CB of IpStdCInterface.h); Ipopt itself is not needed to build.
 *
 *   gcc -O2 -Iinclude examples/qlnlp_demo.c -Lquadruped_landing_b200 -lqlnlp -lm -o qlnlp_demo
 *   LD_LIBRARY_PATH=quadruped_landing_b200 ./qlnlp_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qlnlp.h"

#define NK 61
#define KT 21

/* ---- Ipopt C-interface shaped callbacks (IpStdCInterface.h), user_data = qlnlp_handle ------------------- */
typedef double Number;
typedef int Index;
typedef int Bool;

/* Ipopt passes new_x = FALSE when x is the vector of the previous callback.  The library does the bookkeeping itself:
 * the first callback at a new x evaluates f, grad f, g and the Jacobian values with ONE launch and the callbacks that
 * follow at the same x are served from that evaluation (qlnlp.h, "single evaluations"), so the four callbacks of an
 * iterate cost one H2D + one kernel + one D2H.  new_x is honoured on top of that: when Ipopt says the vector is new
 * the whole bundle is requested at once through qlnlp_eval_all into the callback's own output. */
static Bool eval_f(Index n, Number* x, Bool new_x, Number* obj_value, void* user_data)
{
    (void)n;
    if (new_x) return qlnlp_eval_all((qlnlp_handle)user_data, x, obj_value, NULL, NULL, NULL) == QLNLP_OK;
    return qlnlp_eval_objective((qlnlp_handle)user_data, x, obj_value) == QLNLP_OK;
}
static Bool eval_grad_f(Index n, Number* x, Bool new_x, Number* grad_f, void* user_data)
{
    (void)n;
    if (new_x) return qlnlp_eval_all((qlnlp_handle)user_data, x, NULL, grad_f, NULL, NULL) == QLNLP_OK;
    return qlnlp_eval_objective_gradient((qlnlp_handle)user_data, x, grad_f) == QLNLP_OK;
}
static Bool eval_g(Index n, Number* x, Bool new_x, Index m, Number* g, void* user_data)
{
    (void)n; (void)m;
    if (new_x) return qlnlp_eval_all((qlnlp_handle)user_data, x, NULL, NULL, g, NULL) == QLNLP_OK;
    return qlnlp_eval_constraint((qlnlp_handle)user_data, x, g) == QLNLP_OK;
}
/* structure call (values == NULL) or value call; Ipopt's C interface uses 0-based or 1-based indices by option */
static Bool eval_jac_g(Index n, Number* x, Bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol, Number* values,
                       void* user_data)
{
    qlnlp_handle h = (qlnlp_handle)user_data;
    (void)n; (void)new_x; (void)m;
    if (!values) {
        int64_t* r = malloc(sizeof(int64_t) * (size_t)nele_jac);
        int64_t* c = malloc(sizeof(int64_t) * (size_t)nele_jac);
        const int ok = qlnlp_jacobian_structure(h, r, c) == QLNLP_OK;
        for (Index i = 0; ok && i < nele_jac; ++i) { iRow[i] = (Index)r[i]; jCol[i] = (Index)c[i]; }   /* index_style = 1 */
        free(r); free(c);
        return ok;
    }
    return qlnlp_eval_constraint_jacobian(h, x, values) == QLNLP_OK;
}

/* 0.5*x'Q*x for diagonal Q, folded left (quadratic_cost.jl:38) */
static double half_quad(const double* x, const double* d, int n)
{
    double r = (0.5 * (x[0] * d[0])) * x[0];
    for (int j = 1; j < n; ++j) r += (0.5 * (x[j] * d[j])) * x[j];
    return r;
}

int main(void)
{
    const qlnlp_model model = {-9.81, 10.0, 0.1, 0.5, 0.25, 0.25};
    const double lb = model.lb, l1 = model.l1, l2 = model.l2, dt = 0.009;
    double xinit[15] = {0}, xterm[15] = {0};
    xinit[0] = -lb / 2.5; xinit[1] = sqrt(l1 * l1 + l2 * l2) + 0.1; xinit[2] = -30 * M_PI / 180;
    xinit[5] = -lb; xinit[6] = 0.2; xinit[8] = -sqrt(2 * 9.81 * 2); xinit[9] = -M_PI / 2; xinit[13] = -1.0;
    xterm[0] = -lb / 2; xterm[1] = sqrt(l1 * l1 + l2 * l2); xterm[5] = -lb;

    /* reference trajectory (ref_traj.jl) and LQR cost tables (quadratic_cost.jl:33-42, main.ipynb:152-161) */
    static double Q[NK][15], R[NK][5], q[NK][15], r[NK][5], c[NK], Uref[NK - 1][5];
    const double Qd[15] = {10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 0};
    const double Rd[5] = {1e-3, 1e-2, 1e-3, 1e-2, 0.0};
    for (int k = 0; k < NK - 1; ++k) {
        memset(Uref[k], 0, sizeof Uref[k]);
        if (k < KT - 1) { Uref[k][1] = -model.mb * model.g; Uref[k][4] = 0.001; }
        else { Uref[k][1] = Uref[k][3] = -model.mb * model.g / 2; Uref[k][4] = 0.02; }
    }
    for (int k = 0; k < NK; ++k) {
        double xref[15], uref[5];
        memcpy(xref, xterm, sizeof xref);
        xref[14] = dt * k;
        memcpy(uref, Uref[k < NK - 1 ? k : 0], sizeof uref);
        for (int i = 0; i < 15; ++i) { Q[k][i] = Qd[i]; q[k][i] = (-Qd[i]) * xref[i]; }
        for (int i = 0; i < 5; ++i) { R[k][i] = (k < NK - 1) ? Rd[i] : Rd[i] * 0; r[k][i] = (-R[k][i]) * uref[i]; }
        c[k] = half_quad(xref, Q[k], 15) + half_quad(uref, R[k], 5);
    }

    qlnlp_problem_desc d;
    memset(&d, 0, sizeof d);
    d.N = NK; d.k_trans = KT; d.init_mode = 1; d.model = model;
    memcpy(d.x0, xinit, sizeof xinit); memcpy(d.xf, xterm, sizeof xterm);
    d.Q = &Q[0][0]; d.R = &R[0][0]; d.q = &q[0][0]; d.r = &r[0][0]; d.c = c;

    qlnlp_handle h;
    if (qlnlp_create(&d, 0, QLNLP_JAC_SPARSE_BLOCK, &h)) { fprintf(stderr, "create: %s\n", qlnlp_last_error()); return 1; }
    int64_t n, m, nnz, nnzb;
    qlnlp_dims(h, &n, &m, &nnz, &nnzb);
    printf("variables %lld  constraints %lld  jacobian nonzeros %lld (dense structure would be %lld)\n",
           (long long)n, (long long)m, (long long)nnz, (long long)(n * m));

    /* initial guess Z0 = packZ(nlp, Xguess, Uref), main.ipynb:181-196 */
    double* Z = calloc((size_t)n, sizeof(double));
    double t = 0.0;
    for (int k = 1; k <= NK; ++k) {
        double* x = Z + 20 * (k - 1);
        for (int i = 0; i < 14; ++i)
            x[i] = (k <= KT) ? xinit[i] + (xterm[i] - xinit[i]) / (KT - 1) * (k - 1) : xterm[i];
        x[14] = t;
        t += (k < KT) ? 0.001 : 0.02;
        if (k < NK) memcpy(x + 15, Uref[k - 1], sizeof Uref[0]);
    }

    double f, *grad = malloc(sizeof(double) * n), *g = malloc(sizeof(double) * m), *vals = malloc(sizeof(double) * nnz);
    Index* iRow = malloc(sizeof(Index) * nnz), *jCol = malloc(sizeof(Index) * nnz);
    /* one iterate as Ipopt drives it: new_x only on the first callback */
    if (!eval_f((Index)n, Z, 1, &f, h) || !eval_grad_f((Index)n, Z, 0, grad, h) || !eval_g((Index)n, Z, 0, (Index)m, g, h) ||
        !eval_jac_g((Index)n, Z, 0, (Index)m, (Index)nnz, iRow, jCol, NULL, h) ||
        !eval_jac_g((Index)n, Z, 0, (Index)m, (Index)nnz, NULL, NULL, vals, h)) {
        fprintf(stderr, "evaluation failed: %s\n", qlnlp_last_error());
        return 2;
    }
    double inf_pr = 0.0, gn = 0.0;
    for (int i = 0; i < 1032; ++i) inf_pr = fmax(inf_pr, fabs(g[i]));      /* equality rows */
    for (int i = 0; i < n; ++i) gn = fmax(gn, fabs(grad[i]));
    printf("objective(Z0) = %.16e   inf_pr(Z0) = %.3e   |grad|_inf = %.6e\n", f, inf_pr, gn);
    printf("first structure entries: (%d,%d) (%d,%d) ... last (%d,%d); J[0] = %g\n", iRow[0], jCol[0], iRow[1], jCol[1],
           iRow[nnz - 1], jCol[nnz - 1], vals[0]);
    qlnlp_destroy(h);
    return 0;
}
