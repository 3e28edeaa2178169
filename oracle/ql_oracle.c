/*
 * ql_oracle.c -- CPU oracle (see ql_oracle.h for the status header).
 * TEST INFRASTRUCTURE ONLY: never linked into or called by the product path.
 *
 * Build: gcc -O3 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * Every function cites the reference lines (relative to /root/reference/src/) it restates.
 */
#include "ql_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NX QLO_NX
#define NU QLO_NU
#define NZK (NX + NU)
#define NP 20 /* ForwardDiff chunk: the input of contactM_jacobian is an SVector{20} */

/* ------------------------------------------------------------------ scalar instantiation */
#define T double
#define NAME(f) f##_f64
#define ADD(a, b) ((a) + (b))
#define SUB(a, b) ((a) - (b))
#define MUL(a, b) ((a) * (b))
#define NEG(a) (-(a))
#define DIVC(a, c) ((a) / (c))
#define ADDC(a, c) ((a) + (c))
#define MULC(c, a) ((c) * (a))
#define ZERO() (0.0)
#include "ql_dyn_impl.inc"
#undef T
#undef NAME
#undef ADD
#undef SUB
#undef MUL
#undef NEG
#undef DIVC
#undef ADDC
#undef MULC
#undef ZERO

/* ------------------------------------------------------------------ dual instantiation
 * ForwardDiff 0.10.25 Dual{Tag,Float64,20}: dual.jl binary ops, partials.jl tuple ops. */
typedef struct {
    double v;
    double p[NP];
} dual;

static inline dual d_add(dual a, dual b)
{
    dual r; int i;
    r.v = a.v + b.v;
    for (i = 0; i < NP; ++i) r.p[i] = a.p[i] + b.p[i];
    return r;
}
static inline dual d_sub(dual a, dual b)
{
    dual r; int i;
    r.v = a.v - b.v;
    for (i = 0; i < NP; ++i) r.p[i] = a.p[i] - b.p[i];
    return r;
}
/* Dual*Dual: Dual(vx*vy, _mul_partials(px, py, vy, vx)) = (vy*px_i) + (vx*py_i) */
static inline dual d_mul(dual x, dual y)
{
    dual r; int i;
    r.v = x.v * y.v;
    for (i = 0; i < NP; ++i) r.p[i] = (y.v * x.p[i]) + (x.v * y.p[i]);
    return r;
}
static inline dual d_neg(dual a)
{
    dual r; int i;
    r.v = -a.v;
    for (i = 0; i < NP; ++i) r.p[i] = -a.p[i];
    return r;
}
static inline dual d_divc(dual a, double c)
{
    dual r; int i;
    r.v = a.v / c;
    for (i = 0; i < NP; ++i) r.p[i] = a.p[i] / c;
    return r;
}
static inline dual d_addc(dual a, double c)
{
    dual r = a;
    r.v = a.v + c;
    return r;
}
static inline dual d_mulc(double c, dual a)
{
    dual r; int i;
    r.v = c * a.v;
    for (i = 0; i < NP; ++i) r.p[i] = a.p[i] * c;
    return r;
}
static inline dual d_zero(void)
{
    dual r;
    memset(&r, 0, sizeof r);
    return r;
}

#define T dual
#define NAME(f) f##_dual
#define ADD(a, b) d_add((a), (b))
#define SUB(a, b) d_sub((a), (b))
#define MUL(a, b) d_mul((a), (b))
#define NEG(a) d_neg((a))
#define DIVC(a, c) d_divc((a), (c))
#define ADDC(a, c) d_addc((a), (c))
#define MULC(c, a) d_mulc((c), (a))
#define ZERO() d_zero()
#include "ql_dyn_impl.inc"
#undef T
#undef NAME

/* ------------------------------------------------------------------ public RK4 + Jacobian */
void qlo_rk4(const qlo_model *m, int mode, const double *x, const double *u, double *xn)
{
    contact_dynamics_rk4_f64(m, mode, x, u, xn);
}

/* planar_quadruped.jl:225-248: ForwardDiff.jacobian(z -> rk4(model, z[1:15], z[16:20]), [x;u]) */
void qlo_rk4_jacobian(const qlo_model *m, int mode, const double *x, const double *u,
                      double *xn, double *J)
{
    dual z[NZK], out[NX];
    int i, j;
    for (i = 0; i < NZK; ++i) {
        memset(&z[i], 0, sizeof(dual));
        z[i].v = (i < NX) ? x[i] : u[i - NX];
        z[i].p[i] = 1.0;
    }
    contact_dynamics_rk4_dual(m, mode, z, z + NX, out);
    for (i = 0; i < NX; ++i) {
        if (xn) xn[i] = out[i].v;
        for (j = 0; j < NZK; ++j) J[i + NX * j] = out[i].p[j];
    }
}

/* ------------------------------------------------------------------ index maps, nlp.jl:38-63 */
typedef struct {
    int64_t N, n_nlp, m_nlp;
    int64_t c_init, c_term, c_dyn, c_cfirst, c_cother, c_fctrl, c_body; /* 0-based block starts */
    int64_t n_cother;
    int64_t c_kin;            /* first kinematic row (== the reference's m_nlp); 2N rows when switched on */
} ql_dims;

static ql_dims make_dims(const qlo_problem *p)
{
    ql_dims d;
    d.N = p->N;
    d.n_nlp = NX * p->N + NU * (p->N - 1);                 /* nlp.jl:72 */
    d.c_init = 0;                                          /* 1:n */
    d.c_term = d.c_init + NX;                              /* n-1 rows */
    d.c_dyn = d.c_term + (NX - 1);                         /* (N-1)*n rows */
    d.c_cfirst = d.c_dyn + (p->N - 1) * NX;                /* N rows */
    d.c_cother = d.c_cfirst + p->N;                        /* N-k_trans+1 rows */
    d.n_cother = p->N - p->k_trans + 1;
    d.c_fctrl = d.c_cother + d.n_cother;                   /* 1 row */
    d.c_body = d.c_fctrl + 1;                              /* N rows */
    d.m_nlp = d.c_body + p->N;
    d.c_kin = d.m_nlp;                                     /* nlp.jl:60 (commented out): 2 rows per knot */
    if (p->kinematics) d.m_nlp += 2 * p->N;
    return d;
}
static inline int64_t xind(int64_t k) { return k * NZK; }        /* 0-based knot k -> first x index */
static inline int64_t uind(int64_t k) { return k * NZK + NX; }

int64_t qlo_num_primals(const qlo_problem *p) { return make_dims(p).n_nlp; }
int64_t qlo_num_duals(const qlo_problem *p) { return make_dims(p).m_nlp; }

/* ------------------------------------------------------------------ costs */
/* Adjoint(SVector)*SVector -> dot: ret accumulates a[j]*b[j] left to right. */
static double dot_n(const double *a, const double *b, int n)
{
    double ret = a[0] * b[0];
    int j;
    for (j = 1; j < n; ++j) ret += a[j] * b[j];
    return ret;
}
/* 0.5 * x'Q * x == dot(0.5 .* (x .* Qdiag), x)   (quadratic_cost.jl:46) */
static double half_quad(const double *x, const double *Qd, int n)
{
    double t[NX];
    int j;
    for (j = 0; j < n; ++j) t[j] = 0.5 * (x[j] * Qd[j]);
    return dot_n(t, x, n);
}

/* quadratic_cost.jl:33-42 */
void qlo_lqr_cost(const double *Qd, const double *Rd, const double *xf, const double *uf,
                  double *q, double *r, double *c)
{
    int i;
    for (i = 0; i < NX; ++i) q[i] = (-Qd[i]) * xf[i];      /* q = -Q * xf */
    for (i = 0; i < NU; ++i) r[i] = (-Rd[i]) * uf[i];      /* r = -R * uf */
    *c = half_quad(xf, Qd, NX) + half_quad(uf, Rd, NU);    /* 0.5*xf'Q*xf + 0.5*uf'R*uf */
}

/* quadratic_cost.jl:44-47: 0.5*x'Q*x + q'x + 0.5*u'R*u + r'u + c, summed left to right */
static double stagecost(const qlo_problem *p, int64_t k, const double *x, const double *u)
{
    const double *Q = p->Q + k * NX, *R = p->R + k * NU, *q = p->q + k * NX, *r = p->r + k * NU;
    return (((half_quad(x, Q, NX) + dot_n(q, x, NX)) + half_quad(u, R, NU)) + dot_n(r, u, NU)) + p->c[k];
}
/* quadratic_cost.jl:49-52 */
static double termcost(const qlo_problem *p, int64_t k, const double *x)
{
    const double *Q = p->Q + k * NX, *q = p->q + k * NX;
    return (half_quad(x, Q, NX) + dot_n(q, x, NX)) + p->c[k];
}

/* costs.jl:6-16 */
double qlo_eval_f(const qlo_problem *p, const double *Z)
{
    double J = 0.0;
    int64_t k;
    for (k = 0; k < p->N - 1; ++k) {
        const double *x = Z + xind(k), *u = Z + uind(k);
        const double hk = u[NU - 1];
        J += hk * stagecost(p, k, x, u);
    }
    J += termcost(p, p->N - 1, Z + xind(p->N - 1));
    return J;
}

/* costs.jl:23-34 (note quirk Q1: d/dh of h*stagecost is NOT included) */
void qlo_grad_f(const qlo_problem *p, const double *Z, double *grad)
{
    int64_t k;
    int i;
    for (k = 0; k < p->N - 1; ++k) {
        const double *x = Z + xind(k), *u = Z + uind(k);
        const double hk = u[NU - 1];
        const double *Q = p->Q + k * NX, *R = p->R + k * NU, *q = p->q + k * NX, *r = p->r + k * NU;
        for (i = 0; i < NX; ++i) grad[xind(k) + i] = hk * (Q[i] * x[i] + q[i]);
        for (i = 0; i < NU; ++i) grad[uind(k) + i] = hk * (R[i] * u[i] + r[i]);
    }
    k = p->N - 1;
    {
        const double *x = Z + xind(k);
        const double *Q = p->Q + k * NX, *q = p->q + k * NX;
        for (i = 0; i < NX; ++i) grad[xind(k) + i] = Q[i] * x[i] + q[i];
    }
}

/* ------------------------------------------------------------------ constraints */
/* planar_quadruped.jl:250-260 (both maps are identical) */
static void jump_map(const double *x, double *xn)
{
    xn[0] = x[0]; xn[1] = x[1]; xn[2] = x[2]; xn[3] = x[3];
    xn[4] = 0.0;
    xn[5] = x[5];
    xn[6] = 0.0;
    xn[7] = x[7]; xn[8] = x[8]; xn[9] = x[9];
    xn[10] = 0.0; xn[11] = 0.0; xn[12] = 0.0; xn[13] = 0.0;
    xn[14] = x[14];
}
/* planar_quadruped.jl:262-263 (quirk Q2: position 15 is 0) */
static const int JUMP_DIAG[NX] = {1, 1, 1, 1, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0};

/* mode of knot k (1-based) per constraints.jl:23-37 / :184-198: returns mode and whether to jump */
static int knot_mode(const qlo_problem *p, int64_t k1, int *jump)
{
    *jump = 0;
    if (k1 < p->k_trans - 1) return (p->init_mode == 1) ? 1 : 2;
    if (k1 == p->k_trans - 1) {
        *jump = 1;
        return (p->init_mode == 1) ? 1 : 2;
    }
    return 3;
}

static void eval_c_impl(const qlo_problem *p, const double *x0, const double *xf,
                        const double *Z, double *c)
{
    const ql_dims d = make_dims(p);
    const int64_t N = p->N;
    int64_t k;
    int i;

    /* constraints.jl:149-150 */
    for (i = 0; i < NX; ++i) c[d.c_init + i] = Z[xind(0) + i] - x0[i];
    for (i = 0; i < NX - 1; ++i) c[d.c_term + i] = Z[xind(N - 1) + i] - xf[i];

    /* dynamics_constraint!, constraints.jl:6-41 */
    for (k = 0; k < N - 1; ++k) {
        const double *x = Z + xind(k), *u = Z + uind(k), *xnext = Z + xind(k + 1);
        double xn[NX], xj[NX];
        int jump;
        const int mode = knot_mode(p, k + 1, &jump);
        contact_dynamics_rk4_f64(&p->model, mode, x, u, xn);
        if (jump) {
            jump_map(xn, xj);
            for (i = 0; i < NX; ++i) c[d.c_dyn + k * NX + i] = xj[i] - xnext[i];
        } else {
            for (i = 0; i < NX; ++i) c[d.c_dyn + k * NX + i] = xn[i] - xnext[i];
        }
    }
    /* contact_init_constraints!, constraints.jl:48-65 */
    for (k = 0; k < N; ++k) c[d.c_cfirst + k] = Z[xind(k) + ((p->init_mode == 1) ? 4 : 6)];
    /* contact_another_constraints!, constraints.jl:72-91 */
    for (k = 0; k < d.n_cother; ++k) {
        const int64_t kk = k + p->k_trans - 1; /* 0-based knot */
        c[d.c_cother + k] = Z[xind(kk) + ((p->init_mode == 1) ? 6 : 4)];
    }
    /* constraints.jl:154 */
    {
        const double *ul = Z + uind(N - 2);
        c[d.c_fctrl] = ul[1] + ul[3] + p->model.mb * p->model.g;
    }
    /* body_pos_constraints!, constraints.jl:98-113 */
    for (k = 0; k < N; ++k) {
        const double yb = Z[xind(k) + 1], theta = Z[xind(k) + 2];
        c[d.c_body + k] = yb - p->model.lb / 2 * fabs(sin(theta));
    }
    /* kinematics_constraints!, constraints.jl:115-138 (commented out upstream; opt-in here):
     * d[2k-1] = norm(pb - p1), d[2k] = norm(pb - p2); norm of a 2-vector = sqrt(dx*dx + dy*dy) */
    if (p->kinematics)
        for (k = 0; k < N; ++k) {
            const double *x = Z + xind(k);
            const double d1x = x[0] - x[3], d1y = x[1] - x[4], d2x = x[0] - x[5], d2y = x[1] - x[6];
            c[d.c_kin + 2 * k] = sqrt(d1x * d1x + d1y * d1y);
            c[d.c_kin + 2 * k + 1] = sqrt(d2x * d2x + d2y * d2y);
        }
}

void qlo_eval_c(const qlo_problem *p, const double *Z, double *c)
{
    eval_c_impl(p, p->x0, p->xf, Z, c);
}

void qlo_constraint_bounds(const qlo_problem *p, double *lb, double *ub)
{
    const ql_dims d = make_dims(p);
    int64_t i;
    for (i = 0; i < d.m_nlp; ++i) { lb[i] = 0.0; ub[i] = 0.0; }   /* nlp.jl:66-67 */
    for (i = 0; i < p->N; ++i) ub[d.c_body + i] = INFINITY;       /* nlp.jl:69 */
    if (p->kinematics)                                            /* nlp.jl:70 (commented out upstream) */
        for (i = 0; i < 2 * p->N; ++i) ub[d.c_kin + i] = p->model.l1 + p->model.l2 + p->model.lb / 2;
}

/* moi.jl:51-67, including the "lower bound of F" indices 22/24 + 20(k-1) as written */
void qlo_variable_bounds(const qlo_problem *p, double *xl, double *xu)
{
    const ql_dims d = make_dims(p);
    int64_t i, k;
    for (i = 0; i < d.n_nlp; ++i) { xl[i] = -INFINITY; xu[i] = INFINITY; }
    for (k = 1; k <= p->N; ++k) {
        xl[3 + 20 * (k - 1) - 1] = -M_PI / 2;
        xu[3 + 20 * (k - 1) - 1] = M_PI / 2;
        if (k < p->N) {
            xl[20 + 20 * (k - 1) - 1] = 0.001;
            xu[20 + 20 * (k - 1) - 1] = 0.02;
            xl[22 + 20 * (k - 1) - 1] = 0.0;
            xl[24 + 20 * (k - 1) - 1] = 0.0;
        }
    }
}

/* The one routine that performs jac_c!'s assignments; `put` abstracts the destination
 * (dense matrix, assignment mask, or sparse values through a plan). */
typedef void (*put_fn)(void *ctx, int64_t row, int64_t col, double v); /* 0-based */

static void jac_c_assign(const qlo_problem *p, const double *Z, put_fn put, void *ctx)
{
    const ql_dims d = make_dims(p);
    const int64_t N = p->N;
    int64_t k;
    int i, j;

    /* constraints.jl:228  jac_init .= I(n) */
    for (j = 0; j < NX; ++j)
        for (i = 0; i < NX; ++i) put(ctx, d.c_init + i, xind(0) + j, (i == j) ? 1.0 : 0.0);
    /* constraints.jl:229  jac_term .= I(n)[1:n-1, :] */
    for (j = 0; j < NX; ++j)
        for (i = 0; i < NX - 1; ++i) put(ctx, d.c_term + i, xind(N - 1) + j, (i == j) ? 1.0 : 0.0);

    /* dynamics_jacobian!, constraints.jl:168-205 */
    for (k = 0; k < N - 1; ++k) {
        const double *x = Z + xind(k), *u = Z + uind(k);
        double J[NX * NZK];
        int jump;
        const int mode = knot_mode(p, k + 1, &jump);
        qlo_rk4_jacobian(&p->model, mode, x, u, NULL, J);
        for (j = 0; j < NZK; ++j)
            for (i = 0; i < NX; ++i) {
                double v = J[i + NX * j];
                if (jump) v = (double)JUMP_DIAG[i] * v;     /* Diagonal{Int} * Matrix, :192-194 */
                put(ctx, d.c_dyn + k * NX + i, xind(k) + j, v);
            }
        /* :200  D[ci, xi[k+1]] .= -I(n) */
        for (j = 0; j < NX; ++j)
            for (i = 0; i < NX; ++i) put(ctx, d.c_dyn + k * NX + i, xind(k + 1) + j, (i == j) ? -1.0 : 0.0);
    }
    /* constraints.jl:235-243 */
    for (k = 0; k < N; ++k) put(ctx, d.c_cfirst + k, xind(k) + ((p->init_mode == 1) ? 4 : 6), 1.0);
    /* constraints.jl:246-256 */
    for (k = p->k_trans; k <= N; ++k)
        put(ctx, d.c_cother + (k - p->k_trans), xind(k - 1) + ((p->init_mode == 1) ? 6 : 4), 1.0);
    /* constraints.jl:259-260 */
    put(ctx, d.c_fctrl, uind(N - 2) + 1, 1.0);
    put(ctx, d.c_fctrl, uind(N - 2) + 3, 1.0);
    /* constraints.jl:263-274 (quirk Q4: branch on theta > 0; Q3: `lb` global == model.lb) */
    for (k = 0; k < N; ++k) {
        const double theta = Z[xind(k) + 2];
        const double lb = p->model.lb;
        put(ctx, d.c_body + k, xind(k) + 1, 1.0);
        if (theta > 0)
            put(ctx, d.c_body + k, xind(k) + 2, -lb / 2 * cos(theta));
        else
            put(ctx, d.c_body + k, xind(k) + 2, lb / 2 * cos(theta));
    }
    /* jac_kinematics, constraints.jl:276-288 (commented out upstream, and written there with the state indices of an
     * older 10-state model, x[7:8] / x[9:10]; restated with pb = x[1:2], p1 = x[4:5], p2 = x[6:7] as the constraint
     * itself uses them, :128-134):  d/dpb = d / norm(d),  d/dp_i = -d / norm(d) */
    if (p->kinematics)
        for (k = 0; k < N; ++k) {
            const double *x = Z + xind(k);
            const double d1x = x[0] - x[3], d1y = x[1] - x[4], d2x = x[0] - x[5], d2y = x[1] - x[6];
            const double n1 = sqrt(d1x * d1x + d1y * d1y), n2 = sqrt(d2x * d2x + d2y * d2y);
            put(ctx, d.c_kin + 2 * k, xind(k) + 0, d1x / n1);
            put(ctx, d.c_kin + 2 * k, xind(k) + 1, d1y / n1);
            put(ctx, d.c_kin + 2 * k, xind(k) + 3, -d1x / n1);
            put(ctx, d.c_kin + 2 * k, xind(k) + 4, -d1y / n1);
            put(ctx, d.c_kin + 2 * k + 1, xind(k) + 0, d2x / n2);
            put(ctx, d.c_kin + 2 * k + 1, xind(k) + 1, d2y / n2);
            put(ctx, d.c_kin + 2 * k + 1, xind(k) + 5, -d2x / n2);
            put(ctx, d.c_kin + 2 * k + 1, xind(k) + 6, -d2y / n2);
        }
}

typedef struct { double *jac; int64_t m; } dense_ctx;
static void put_dense(void *c, int64_t row, int64_t col, double v)
{
    dense_ctx *d = (dense_ctx *)c;
    d->jac[row + d->m * col] = v;
}
void qlo_jac_c_dense(const qlo_problem *p, const double *Z, double *jac)
{
    dense_ctx c = {jac, make_dims(p).m_nlp};
    jac_c_assign(p, Z, put_dense, &c);
}

typedef struct { unsigned char *mask; int64_t m; } mask_ctx;
static void put_mask(void *c, int64_t row, int64_t col, double v)
{
    mask_ctx *d = (mask_ctx *)c;
    (void)v;
    d->mask[row + d->m * col] = 1;
}
/* assignment mask of jac_c! (values are irrelevant to which entries get assigned) */
static unsigned char *assigned_mask(const qlo_problem *p)
{
    const ql_dims d = make_dims(p);
    double *Z = (double *)calloc((size_t)d.n_nlp, sizeof(double));
    mask_ctx c;
    int64_t k;
    c.m = d.m_nlp;
    c.mask = (unsigned char *)calloc((size_t)(d.m_nlp * d.n_nlp), 1);
    for (k = 0; k < p->N - 1; ++k) Z[uind(k) + 4] = 0.01;
    jac_c_assign(p, Z, put_mask, &c);
    free(Z);
    return c.mask;
}

int64_t qlo_nnz_block(const qlo_problem *p)
{
    const ql_dims d = make_dims(p);
    unsigned char *mask = assigned_mask(p);
    int64_t i, n = 0;
    for (i = 0; i < d.m_nlp * d.n_nlp; ++i) n += mask[i];
    free(mask);
    return n;
}

/* moi.jl:31-33 orders pairs as vec(CartesianIndices(m x n)): column-major, row fastest. */
void qlo_jacobian_structure(const qlo_problem *p, int64_t *rows, int64_t *cols)
{
    const ql_dims d = make_dims(p);
    unsigned char *mask = assigned_mask(p);
    int64_t r, c, n = 0;
    for (c = 0; c < d.n_nlp; ++c)
        for (r = 0; r < d.m_nlp; ++r)
            if (mask[r + d.m_nlp * c]) {
                rows[n] = r + 1;
                cols[n] = c + 1;
                ++n;
            }
    free(mask);
}

struct qlo_plan {
    int64_t m, n, nnz;
    int32_t *lin2pos; /* dense linear index -> position in the value array, -1 if unassigned */
};

qlo_plan *qlo_plan_create(const qlo_problem *p)
{
    const ql_dims d = make_dims(p);
    unsigned char *mask = assigned_mask(p);
    qlo_plan *pl = (qlo_plan *)malloc(sizeof *pl);
    int64_t i, n = 0;
    pl->m = d.m_nlp;
    pl->n = d.n_nlp;
    pl->lin2pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(d.m_nlp * d.n_nlp));
    for (i = 0; i < d.m_nlp * d.n_nlp; ++i) pl->lin2pos[i] = mask[i] ? (int32_t)n++ : -1;
    pl->nnz = n;
    free(mask);
    return pl;
}
void qlo_plan_destroy(qlo_plan *pl)
{
    if (!pl) return;
    free(pl->lin2pos);
    free(pl);
}

typedef struct { const qlo_plan *pl; double *vals; } sparse_ctx;
static void put_sparse(void *c, int64_t row, int64_t col, double v)
{
    sparse_ctx *s = (sparse_ctx *)c;
    const int32_t pos = s->pl->lin2pos[row + s->pl->m * col];
    if (pos >= 0) s->vals[pos] = v;     /* SPARSE_TRUE plans drop the structural zeros jac_c! assigns */
}
void qlo_jac_c_sparse(const qlo_plan *pl, const qlo_problem *p, const double *Z, double *vals)
{
    sparse_ctx c = {pl, vals};
    jac_c_assign(p, Z, put_sparse, &c);
}

/* ---- SPARSE_TRUE: the assigned entries that are not structurally zero.  Determined numerically: an entry
 * belongs to the pattern iff it is non-zero at one of a few generic pseudo-random points (identity blocks
 * contribute their diagonals, the RK4 blocks their mode-specific pattern, the jump knot its unmasked rows). */
typedef struct { unsigned char *mask; int64_t m; } nzmask_ctx;
static void put_nzmask(void *c, int64_t row, int64_t col, double v)
{
    nzmask_ctx *d = (nzmask_ctx *)c;
    if (v != 0.0) d->mask[row + d->m * col] = 1;
}
static unsigned char *nonzero_mask(const qlo_problem *p)
{
    const ql_dims d = make_dims(p);
    double *Z = (double *)malloc(sizeof(double) * (size_t)d.n_nlp);
    nzmask_ctx c;
    uint64_t s = 0x9E3779B97F4A7C15ull;
    int rep;
    int64_t i, k;
    c.m = d.m_nlp;
    c.mask = (unsigned char *)calloc((size_t)(d.m_nlp * d.n_nlp), 1);
    for (rep = 0; rep < 4; ++rep) {
        for (i = 0; i < d.n_nlp; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            Z[i] = 0.25 + 1.5 * (double)(s >> 11) * (1.0 / 9007199254740992.0);   /* in (0.25, 1.75): generic, theta > 0 */
            if (s & 1) Z[i] = -Z[i];
        }
        for (k = 0; k < p->N - 1; ++k) Z[uind(k) + 4] = 0.003 + 0.001 * (double)rep + 1e-4 * (double)(k % 7);
        jac_c_assign(p, Z, put_nzmask, &c);
    }
    free(Z);
    return c.mask;
}

static qlo_plan *plan_from_mask(const qlo_problem *p, unsigned char *mask)
{
    const ql_dims d = make_dims(p);
    qlo_plan *pl = (qlo_plan *)malloc(sizeof *pl);
    int64_t i, n = 0;
    pl->m = d.m_nlp;
    pl->n = d.n_nlp;
    pl->lin2pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(d.m_nlp * d.n_nlp));
    for (i = 0; i < d.m_nlp * d.n_nlp; ++i) pl->lin2pos[i] = mask[i] ? (int32_t)n++ : -1;
    pl->nnz = n;
    free(mask);
    return pl;
}
qlo_plan *qlo_plan_create_true(const qlo_problem *p) { return plan_from_mask(p, nonzero_mask(p)); }
int64_t qlo_plan_nnz(const qlo_plan *pl) { return pl->nnz; }
/* rows/cols (1-based, column-major) of the plan's pattern */
void qlo_plan_structure(const qlo_plan *pl, int64_t *rows, int64_t *cols)
{
    int64_t r, c;
    for (c = 0; c < pl->n; ++c)
        for (r = 0; r < pl->m; ++r) {
            const int32_t pos = pl->lin2pos[r + pl->m * c];
            if (pos >= 0) { rows[pos] = r + 1; cols[pos] = c + 1; }
        }
}

int qlo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void qlo_eval_batch(const qlo_plan *pl, const qlo_problem *p, int64_t B,
                    const double *Z, int64_t ldz, const double *x0, const double *xf,
                    double *f, double *grad, int64_t ldgrad, double *g, int64_t ldg,
                    double *jac, int64_t ldjac, int nthreads)
{
    int64_t b;
    (void)nthreads;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (b = 0; b < B; ++b) {
        const double *Zb = Z + b * ldz;
        if (f) f[b] = qlo_eval_f(p, Zb);
        if (grad) qlo_grad_f(p, Zb, grad + b * ldgrad);
        if (g) eval_c_impl(p, x0 ? x0 + b * NX : p->x0, xf ? xf + b * NX : p->xf, Zb, g + b * ldg);
        if (jac) qlo_jac_c_sparse(pl, p, Zb, jac + b * ldjac);
    }
}
