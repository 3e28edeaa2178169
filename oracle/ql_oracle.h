/*
 * ql_oracle.h -- CPU oracle for the planar-quadruped landing NLP evaluator.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * Julia evaluator (zixinz990/quadruped_landing, src/{costs,constraints,
 * planar_quadruped,quadratic_cost,nlp,moi}.jl).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it,
 * and only as the checker.  The product (quadruped_landing_b200/) never links
 * or calls anything in oracle/.
 *
 * PARITY STATUS: eval_f and eval_c are pinned by the reference's recorded
 * notebook outputs (src/main.ipynb:710,712 on src/data_6.csv, reproduced to the
 * last printed digit -- see tests/test_oracle_kat.py).  grad_f and jac_c are
 * "parity unpinned": nothing in the reference pins their values numerically
 * and Julia is not installed, so they are pinned only by the source text and by
 * an independent extended-precision derivative (tests/test_oracle_hp.py).
 *
 * Arithmetic rules reproduced (all IEEE fp64, no FMA contraction; compile with
 * -ffp-contract=off):
 *   - Julia n-ary + and * fold left; `-a*b` is `(-a)*b`; `0.5*h*f` is `(0.5*h)*f`.
 *   - ForwardDiff 0.10.25 (Manifest.toml:265-269, not vendored) dual rules:
 *       (a*b).p_i = (b.v * a.p_i) + (a.v * b.p_i)          (dual.jl `*`, partials.jl mul_tuples)
 *       (a/c).p_i = a.p_i / c, (a*c).p_i = a.p_i * c        (c a plain real)
 *       (a+-b).p_i = a.p_i +- b.p_i,  (-a).p_i = -a.p_i
 *     One 20-wide chunk (input is an SVector{20}).
 */
#ifndef QL_ORACLE_H
#define QL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QLO_NX 15
#define QLO_NU 5

/* planar_quadruped.jl:11-20 */
typedef struct {
    double g, mb, mf, lb, l1, l2;
} qlo_model;

/* nlp.jl:13-33 (fields the evaluators read) + quadratic_cost.jl:16-22 per knot */
typedef struct {
    int64_t N, k_trans, init_mode;
    qlo_model model;
    double x0[QLO_NX];
    double xf[QLO_NX];
    /* per-knot cost, knot-major: Q[k*15+i], R[k*5+i], q[k*15+i], r[k*5+i], c[k]; k = 0..N-1 */
    const double *Q, *R, *q, *r, *c;
    /* non-zero: the kinematic (leg-length) rows the reference carries COMMENTED OUT are switched on: cinds[8] = 2N rows
     * d[2k-1] = norm(pb - p1), d[2k] = norm(pb - p2) with bounds [0, l1 + l2 + lb/2] (nlp.jl:60,70; constraints.jl:115-138,
     * 276-288).  Default 0 = the reference as it runs. */
    int64_t kinematics;
} qlo_problem;

int64_t qlo_num_primals(const qlo_problem *p);                 /* nlp.jl:86 */
int64_t qlo_num_duals(const qlo_problem *p);                   /* nlp.jl:87 */
int64_t qlo_nnz_block(const qlo_problem *p);                   /* count of entries jac_c! assigns */

/* quadratic_cost.jl:33-42: q = -Q*xf, r = -R*uf, c = 0.5 xf'Q xf + 0.5 uf'R uf (Q,R diagonals) */
void qlo_lqr_cost(const double *Qd, const double *Rd, const double *xf, const double *uf,
                  double *q, double *r, double *c);

double qlo_eval_f(const qlo_problem *p, const double *Z);                  /* costs.jl:6-16 */
void qlo_grad_f(const qlo_problem *p, const double *Z, double *grad);     /* costs.jl:23-34 */
void qlo_eval_c(const qlo_problem *p, const double *Z, double *c);        /* constraints.jl:145-158 */
/* constraints.jl:212-291 on a dense column-major m_nlp x n_nlp matrix.  Like the
 * reference it only ASSIGNS the entries it touches; the caller zeroes `jac`. */
void qlo_jac_c_dense(const qlo_problem *p, const double *Z, double *jac);
/* bounds, nlp.jl:66-69 */
void qlo_constraint_bounds(const qlo_problem *p, double *lb, double *ub);
/* moi.jl:51-67 (variable bounds set by solve()) */
void qlo_variable_bounds(const qlo_problem *p, double *xl, double *xu);

/* SPARSE_BLOCK structure = column-major filter (moi.jl:31-33 ordering) of the
 * entries jac_c! assigns.  rows/cols are 1-based, length qlo_nnz_block(). */
void qlo_jacobian_structure(const qlo_problem *p, int64_t *rows, int64_t *cols);

/* A "plan" caches dense-linear-index -> sparse-position so batches can be
 * evaluated without a 10 MB dense temp per evaluation. */
typedef struct qlo_plan qlo_plan;
qlo_plan *qlo_plan_create(const qlo_problem *p);
void qlo_plan_destroy(qlo_plan *pl);
/* SPARSE_TRUE plan: the assigned entries that are not structurally zero (identity blocks as diagonals, RK4
 * blocks as their mode pattern).  qlo_jac_c_sparse / qlo_eval_batch work with either kind of plan. */
qlo_plan *qlo_plan_create_true(const qlo_problem *p);
int64_t qlo_plan_nnz(const qlo_plan *pl);
void qlo_plan_structure(const qlo_plan *pl, int64_t *rows, int64_t *cols);
/* same assignments as qlo_jac_c_dense, written into vals[nnz] in structure order */
void qlo_jac_c_sparse(const qlo_plan *pl, const qlo_problem *p, const double *Z, double *vals);

/* Batched driver (OpenMP over the batch when compiled with -fopenmp): rows of
 * Z/grad/g/jac have leading dimensions ldz/ldgrad/ldg/ldjac; x0/xf may be NULL
 * (use p->x0/xf) or [B][15] per-evaluation overrides; any output may be NULL. */
void qlo_eval_batch(const qlo_plan *pl, const qlo_problem *p, int64_t B,
                    const double *Z, int64_t ldz, const double *x0, const double *xf,
                    double *f, double *grad, int64_t ldgrad, double *g, int64_t ldg,
                    double *jac, int64_t ldjac, int nthreads);
int qlo_max_threads(void);

/* one RK4 step and its 15x20 Jacobian (column-major J[i + 15*j]); mode 1,2,3.
 * planar_quadruped.jl:189-248 */
void qlo_rk4(const qlo_model *m, int mode, const double *x, const double *u, double *xn);
void qlo_rk4_jacobian(const qlo_model *m, int mode, const double *x, const double *u,
                      double *xn, double *J);

/* ---- Lagrangian Hessian (ql_oracle_hess.c; NO reference target: src/moi.jl:26-28 offers [:Grad, :Jac] only) ----
 * H[i + 20*j] = sum_r lam[r] * d^2 rk4_r / dz_i dz_j by dense second-order forward mode. */
void qlo_rk4_hessian(const qlo_model *m, int mode, const double *x, const double *u, const double *lam, double *H);
/* sigma * Hess f + sum_r lambda_r * Hess g_r: dense n_nlp x n_nlp, column-major, full symmetric matrix */
void qlo_hess_lagrangian_dense(const qlo_problem *p, const double *Z, double sigma, const double *lambda, double *H);

#ifdef __cplusplus
}
#endif
#endif
