"""Solver glue: the caller of the hot path (reference src/moi.jl:46-103).

``solve(x0, prob; tol, c_tol, max_iter)`` in the reference installs the variable bounds (moi.jl:51-67), wraps the
evaluator in an ``MOI.NLPBlockData`` with the constraint bounds (moi.jl:69-74) and hands it to Ipopt with three
options (moi.jl:77-80).  Ipopt is not available in this image, so the same set-up is handed to the interior-point
NLP solver that IS available, ``scipy.optimize.minimize(method="trust-constr")``; when ``cyipopt`` can be imported
it is used instead with exactly the reference's options.  Either way the solver only ever touches the problem
through the four MOI callbacks of ``HybridNLP`` -- this module contains no model arithmetic.  The fallback is a
functional stand-in (it proves the callback boundary inside a solver loop); it is NOT a good solver for this
degenerate NLP -- the contact rows are redundant with the pinned-foot dynamics, so the Jacobian is rank deficient,
which Ipopt regularises and trust-constr does not (the reference's own recorded run ends in "Restoration Failed",
main.ipynb:727).

The point of the sparse structure: the reference reports a dense 1,327,995-entry Jacobian and spends 575 s of a
594 s solve inside Ipopt (src/main.ipynb:217-218,724); here the solver receives the 32,161-entry SPARSE_BLOCK (or
the 4,840-entry SPARSE_TRUE) structure.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from .evaluator import HybridNLP


@dataclass
class SolveResult:
    x: np.ndarray
    objective: float
    constr_violation: float
    iterations: int
    status: str
    evals: Dict[str, int] = field(default_factory=dict)
    backend: str = ""


class _Callbacks:
    """Counts and forwards the four callbacks (the counters are what Ipopt prints at main.ipynb:717-722)."""

    def __init__(self, nlp: HybridNLP):
        import scipy.sparse as sp

        self.nlp = nlp
        self.n = {"f": 0, "grad": 0, "g": 0, "jac": 0}
        rows, cols = nlp.jacobian_structure_arrays()
        self.rows, self.cols = rows - 1, cols - 1
        self._sp = sp
        self._grad = np.empty(nlp.n_nlp)
        self._g = np.empty(nlp.m_nlp)
        self._vals = np.empty(nlp.nnz)
        # exact Lagrangian Hessian when the evaluator offers :Hess (not in the reference, moi.jl:26-28)
        self.has_hess = "Hess" in (nlp.features_available() if hasattr(nlp, "features_available") else ())
        if self.has_hess:
            self.n["hess"] = 0
            hr, hc = nlp.hessian_structure_arrays()
            self.hrows, self.hcols = hr - 1, hc - 1
            self._h = np.empty(len(hr))

    def f(self, x):
        self.n["f"] += 1
        return self.nlp.eval_objective(x)

    def grad(self, x):
        self.n["grad"] += 1
        self.nlp.eval_objective_gradient(self._grad, x)
        return self._grad.copy()

    def g(self, x):
        self.n["g"] += 1
        self.nlp.eval_constraint(self._g, x)
        return self._g.copy()

    def jac(self, x):
        self.n["jac"] += 1
        self.nlp.eval_constraint_jacobian(self._vals, x)
        return self._sp.csr_matrix((self._vals, (self.rows, self.cols)), shape=(self.nlp.m_nlp, self.nlp.n_nlp))


    def hess_values(self, x, sigma, lam):
        self.n["hess"] += 1
        self.nlp.eval_hessian_lagrangian(self._h, x, sigma, lam)
        return self._h

    def hess_matrix(self, x, sigma, lam):
        """sigma Hess f + sum lam_r Hess g_r as a symmetric sparse matrix (the evaluator reports the lower triangle)"""
        v = self.hess_values(x, sigma, lam)
        L = self._sp.csr_matrix((v, (self.hrows, self.hcols)), shape=(self.nlp.n_nlp, self.nlp.n_nlp))
        return L + self._sp.tril(L, -1).T


def solve(x0, nlp: HybridNLP, *, tol: float = 1.0e-6, c_tol: float = 1.0e-6, max_iter: int = 2000,
          backend: Optional[str] = None, verbose: int = 0) -> SolveResult:
    """Same signature and defaults as the reference's ``solve`` (moi.jl:46-47); returns the primal solution and the
    solver statistics.  ``nlp`` must report a sparse structure (``use_sparse_jacobian=True``)."""
    if not nlp.use_sparse_jacobian:
        raise ValueError("solve() needs a sparse structure: HybridNLP(..., use_sparse_jacobian=True)")
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    xl, xu = nlp.variable_bounds()                    # moi.jl:51-67
    cl, cu = nlp.constraint_bounds()                  # prob.lb / prob.ub, moi.jl:69
    cb = _Callbacks(nlp)
    if backend is None:
        try:
            import cyipopt  # noqa: F401
            backend = "ipopt"
        except Exception:
            backend = "trust-constr"

    if backend == "ipopt":
        import cyipopt

        class _P:
            objective = staticmethod(cb.f)
            gradient = staticmethod(cb.grad)
            constraints = staticmethod(cb.g)

            @staticmethod
            def jacobianstructure():
                return cb.rows, cb.cols

            @staticmethod
            def jacobian(x):
                cb.n["jac"] += 1
                nlp.eval_constraint_jacobian(cb._vals, x)
                return cb._vals.copy()

        if cb.has_hess:
            _P.hessianstructure = staticmethod(lambda: (cb.hrows, cb.hcols))
            _P.hessian = staticmethod(lambda x, lagrange, obj_factor: cb.hess_values(x, obj_factor, lagrange).copy())
        prob = cyipopt.Problem(n=nlp.n_nlp, m=nlp.m_nlp, problem_obj=_P(), lb=xl, ub=xu, cl=cl, cu=cu)
        prob.add_option("max_iter", int(max_iter))            # moi.jl:78-80
        prob.add_option("tol", float(tol))
        prob.add_option("constr_viol_tol", float(c_tol))
        if not cb.has_hess:
            prob.add_option("hessian_approximation", "limited-memory")   # no :Hess feature (moi.jl:26-28)
        prob.add_option("print_level", 5 if verbose else 0)
        x, info = prob.solve(x0)
        g = cb.g(x)
        viol = float(np.max(np.maximum(np.maximum(cl - g, g - cu), 0.0)))
        return SolveResult(x, float(info["obj_val"]), viol, -1, str(info["status_msg"]), dict(cb.n), "ipopt")

    if backend != "trust-constr":
        raise ValueError(f"unknown backend {backend!r}")
    from scipy.optimize import Bounds, NonlinearConstraint, minimize, BFGS

    if cb.has_hess:       # exact second derivatives: sigma = 1, lambda = 0 for the objective; sigma = 0, lambda = v for g
        zero = np.zeros(nlp.m_nlp)
        con = NonlinearConstraint(cb.g, cl, cu, jac=cb.jac, hess=lambda x, v: cb.hess_matrix(x, 0.0, v))
        fhess = lambda x: cb.hess_matrix(x, 1.0, zero)
    else:
        con = NonlinearConstraint(cb.g, cl, cu, jac=cb.jac, hess=BFGS())     # quasi-Newton: the evaluator has no Hessian
        fhess = BFGS()
    res = minimize(cb.f, x0, jac=cb.grad, hess=fhess, method="trust-constr", bounds=Bounds(xl, xu, keep_feasible=False),
                   constraints=[con],
                   options={"maxiter": int(max_iter), "gtol": float(tol), "xtol": 1e-12, "verbose": int(verbose),
                            "sparse_jacobian": True, "initial_constr_penalty": 1.0})
    g = cb.g(res.x)
    viol = float(np.max(np.maximum(np.maximum(cl - g, g - cu), 0.0)))
    status = "converged" if (res.status in (1, 2) and viol <= c_tol) else f"status {res.status}: {res.message}"
    return SolveResult(np.asarray(res.x), float(res.fun), viol, int(res.nit), status, dict(cb.n), "trust-constr")
