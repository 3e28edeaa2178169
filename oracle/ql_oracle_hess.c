/*
 * ql_oracle_hess.c -- CPU oracle of the Lagrangian Hessian.
 * TEST INFRASTRUCTURE ONLY (see ql_oracle.h): never linked into or called by the product path.
 *
 * NO REFERENCE TARGET.  The reference advertises [:Grad, :Jac] only (src/moi.jl:26-28; Ipopt runs L-BFGS,
 * src/main.ipynb:219), so nothing upstream computes this matrix.  What is restated here is the mathematics:
 *
 *   H(Z; sigma, lambda) = sigma * Hess f(Z) + sum_r lambda_r * Hess g_r(Z)
 *
 * with f exactly as eval_f (src/costs.jl:6-16, i.e. INCLUDING d/dh of h*stagecost, which grad_f! omits -- quirk Q1
 * concerns the reference's gradient, not its objective) and g exactly as eval_c! (src/constraints.jl:145-158).
 * Only three kinds of rows are non-linear: the dynamics defects (RK4 of the mode's vector field, masked by the jump
 * map at k = k_trans-1; src/constraints.jl:23-37, src/planar_quadruped.jl:36-221,250-260), the body-clearance rows
 * yb - lb/2*|sin(theta)| (src/constraints.jl:98-113; second derivative of the branch the reference's Jacobian takes,
 * theta > 0, src/constraints.jl:269-273) and the objective.  Every one of them couples only the 20 variables of one
 * knot, so H is block diagonal with one 20x20 block per knot (15x15 for the last).
 *
 * Method: dense second-order forward mode -- every number carries its value, 20 first and 20x20 second partials --
 * pushed through the SAME dynamics text the rest of the oracle uses (ql_dyn_impl.inc).  Independent of the
 * generated, sparsity-exploiting device code (tools/gen_rk4_dual.py).
 */
#include "ql_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NX QLO_NX
#define NU QLO_NU
#define NZK (NX + NU)
#define NP 20

typedef struct {
    double v;
    double g[NP];
    double h[NP][NP];      /* full symmetric matrix */
} dual2;

static inline dual2 h_zero(void)
{
    dual2 r;
    memset(&r, 0, sizeof r);
    return r;
}
static inline dual2 h_add(dual2 a, dual2 b)
{
    dual2 r; int i, j;
    r.v = a.v + b.v;
    for (i = 0; i < NP; ++i) { r.g[i] = a.g[i] + b.g[i]; for (j = 0; j < NP; ++j) r.h[i][j] = a.h[i][j] + b.h[i][j]; }
    return r;
}
static inline dual2 h_sub(dual2 a, dual2 b)
{
    dual2 r; int i, j;
    r.v = a.v - b.v;
    for (i = 0; i < NP; ++i) { r.g[i] = a.g[i] - b.g[i]; for (j = 0; j < NP; ++j) r.h[i][j] = a.h[i][j] - b.h[i][j]; }
    return r;
}
static inline dual2 h_neg(dual2 a)
{
    dual2 r; int i, j;
    r.v = -a.v;
    for (i = 0; i < NP; ++i) { r.g[i] = -a.g[i]; for (j = 0; j < NP; ++j) r.h[i][j] = -a.h[i][j]; }
    return r;
}
/* (xy)'' = x'' y + x' y'^T + y' x'^T + x y'' */
static inline dual2 h_mul(dual2 x, dual2 y)
{
    dual2 r; int i, j;
    r.v = x.v * y.v;
    for (i = 0; i < NP; ++i) {
        r.g[i] = (y.v * x.g[i]) + (x.v * y.g[i]);
        for (j = 0; j < NP; ++j)
            r.h[i][j] = ((y.v * x.h[i][j]) + (x.v * y.h[i][j])) + ((x.g[i] * y.g[j]) + (x.g[j] * y.g[i]));
    }
    return r;
}
static inline dual2 h_divc(dual2 a, double c)
{
    dual2 r; int i, j;
    r.v = a.v / c;
    for (i = 0; i < NP; ++i) { r.g[i] = a.g[i] / c; for (j = 0; j < NP; ++j) r.h[i][j] = a.h[i][j] / c; }
    return r;
}
static inline dual2 h_addc(dual2 a, double c)
{
    dual2 r = a;
    r.v = a.v + c;
    return r;
}
static inline dual2 h_mulc(double c, dual2 a)
{
    dual2 r; int i, j;
    r.v = c * a.v;
    for (i = 0; i < NP; ++i) { r.g[i] = a.g[i] * c; for (j = 0; j < NP; ++j) r.h[i][j] = a.h[i][j] * c; }
    return r;
}

#define T dual2
#define NAME(f) f##_dual2
#define ADD(a, b) h_add((a), (b))
#define SUB(a, b) h_sub((a), (b))
#define MUL(a, b) h_mul((a), (b))
#define NEG(a) h_neg((a))
#define DIVC(a, c) h_divc((a), (c))
#define ADDC(a, c) h_addc((a), (c))
#define MULC(c, a) h_mulc((c), (a))
#define ZERO() h_zero()
#include "ql_dyn_impl.inc"
#undef T
#undef NAME

static void seed(dual2 *z, const double *x, const double *u)
{
    int i;
    for (i = 0; i < NZK; ++i) {
        z[i] = h_zero();
        z[i].v = (i < NX) ? x[i] : u[i - NX];
        z[i].g[i] = 1.0;
    }
}

/* H[i + 20*j] = sum_r lam[r] * d^2 rk4_r / dz_i dz_j   (z = [x; u], mode 1, 2, 3) */
void qlo_rk4_hessian(const qlo_model *m, int mode, const double *x, const double *u, const double *lam, double *H)
{
    dual2 z[NZK], out[NX];
    int i, j, r;
    seed(z, x, u);
    contact_dynamics_rk4_dual2(m, mode, z, z + NX, out);
    for (i = 0; i < NZK; ++i)
        for (j = 0; j < NZK; ++j) {
            double acc = 0.0;
            for (r = 0; r < NX; ++r) acc += lam[r] * out[r].h[i][j];
            H[i + NZK * j] = acc;
        }
}

/* quadratic_cost.jl:44-52 on second-order numbers (Q, R diagonal; folded left like half_quad / dot_n in ql_oracle.c) */
static dual2 h_half_quad(const dual2 *x, const double *d, int n)
{
    dual2 ret = h_mul(h_mulc(0.5, h_mulc(d[0], x[0])), x[0]);
    int j;
    for (j = 1; j < n; ++j) ret = h_add(ret, h_mul(h_mulc(0.5, h_mulc(d[j], x[j])), x[j]));
    return ret;
}
static dual2 h_dot(const double *a, const dual2 *x, int n)
{
    dual2 ret = h_mulc(a[0], x[0]);
    int j;
    for (j = 1; j < n; ++j) ret = h_add(ret, h_mulc(a[j], x[j]));
    return ret;
}

static const int JUMP_KEEP[NX] = {1, 1, 1, 1, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 1};      /* the jump MAP keeps the time (planar_quadruped.jl:252) */

/* Dense n_nlp x n_nlp Hessian of the Lagrangian, column-major, FULL symmetric matrix; the caller zeroes nothing
 * (every entry is written).  lambda has m_nlp entries in the order of eval_c! (nlp.jl:48-63). */
void qlo_hess_lagrangian_dense(const qlo_problem *p, const double *Z, double sigma, const double *lambda, double *H)
{
    const int64_t N = p->N, n = qlo_num_primals(p);
    const int64_t c_dyn = NX + (NX - 1);
    const int64_t c_body = qlo_num_duals(p) - N;
    const double a = p->model.lb / 2;
    int64_t k;
    int i, j;
    memset(H, 0, sizeof(double) * (size_t)n * (size_t)n);
    for (k = 0; k < N; ++k) {
        const double *x = Z + NZK * k, *u = x + NX;
        const int nv = (k < N - 1) ? NZK : NX;
        double blk[NZK][NZK];
        memset(blk, 0, sizeof blk);
        if (k < N - 1) {
            dual2 z[NZK], out[NX], cost;
            const int64_t k1 = k + 1;                                      /* 1-based knot */
            const int mode = (k1 >= p->k_trans) ? 3 : (int)p->init_mode;    /* constraints.jl:23-37 */
            const int jump = (k1 == p->k_trans - 1);
            const double *Q = p->Q + k * NX, *R = p->R + k * NU, *q = p->q + k * NX, *r = p->r + k * NU;
            seed(z, x, u);
            /* sigma * h_k * stagecost_k  (costs.jl:12, quadratic_cost.jl:46) */
            cost = h_addc(h_add(h_add(h_add(h_half_quad(z, Q, NX), h_dot(q, z, NX)), h_half_quad(z + NX, R, NU)),
                                h_dot(r, z + NX, NU)), p->c[k]);
            cost = h_mul(z[NZK - 1], cost);
            contact_dynamics_rk4_dual2(&p->model, mode, z, z + NX, out);
            for (i = 0; i < NZK; ++i)
                for (j = 0; j < NZK; ++j) {
                    double acc = sigma * cost.h[i][j];
                    int rr;
                    for (rr = 0; rr < NX; ++rr)
                        if (!jump || JUMP_KEEP[rr]) acc += lambda[c_dyn + NX * k + rr] * out[rr].h[i][j];
                    blk[i][j] = acc;
                }
        } else {
            const double *Q = p->Q + k * NX;                                /* termcost: 0.5 x'Qx + q'x + c */
            for (i = 0; i < NX; ++i) blk[i][i] = sigma * Q[i];
        }
        /* body clearance row: yb - a*|sin(theta)|; the reference's Jacobian uses -a*cos(theta) for theta > 0 and
         * +a*cos(theta) otherwise (constraints.jl:269-273): differentiate that once more */
        {
            const double th = x[2];
            blk[2][2] += lambda[c_body + k] * ((th > 0) ? a * sin(th) : -(a * sin(th)));
        }
        for (i = 0; i < nv; ++i)
            for (j = 0; j < nv; ++j) H[(NZK * k + i) + n * (NZK * k + j)] = blk[i][j];
    }
}
