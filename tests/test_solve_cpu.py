"""solve() glue (reference src/moi.jl:46-103) exercised on the CPU box with the oracle standing in for the GPU
evaluator: the solver must only need the four callbacks + bounds + sparse structure."""
import numpy as np
import pytest

import quadruped_landing_b200 as ql
from quadruped_landing_b200.solve import solve
from oracle_nlp import OracleNLP

pytest.importorskip("scipy")


def test_solve_drives_the_four_callbacks_through_a_solver_loop():
    """The fallback backend (scipy trust-constr) is a functional stand-in for Ipopt, not a good solver for this
    degenerate NLP (redundant contact rows make the Jacobian rank deficient; the reference's own Ipopt run ends in
    "Restoration Failed", main.ipynb:727).  What is checked is the glue: set-up, callbacks, statistics, bounds."""
    p = ql.build_problem(N=9, k_trans=4)
    nlp = OracleNLP(p)
    z0 = ql.initial_guess(p)
    cl, cu = nlp.constraint_bounds()
    g0 = nlp.o.eval_c(z0)
    viol0 = np.max(np.maximum(np.maximum(cl - g0, g0 - cu), 0.0))
    with pytest.warns(UserWarning):
        res = solve(z0, nlp, tol=1e-3, c_tol=1e-3, max_iter=15, backend="trust-constr")
    assert res.backend == "trust-constr" and res.x.shape == (p.n_nlp,) and res.iterations == 15
    assert res.evals["f"] >= 15 and res.evals["grad"] >= 15 and res.evals["g"] >= 15 and res.evals["jac"] >= 15
    assert np.isfinite(res.objective) and res.constr_violation <= viol0 * (1 + 1e-9)
    xl, xu = nlp.variable_bounds()
    assert np.all(res.x >= xl - 1e-6) and np.all(res.x <= xu + 1e-6)      # bounds of moi.jl:51-67 are honoured


def test_solve_uses_the_exact_hessian_when_offered():
    """With :Hess on the glue hands the solver exact second derivatives instead of BFGS updates:
    sigma = 1, lambda = 0 for the objective, sigma = 0, lambda = v for the constraints."""
    import warnings
    p = ql.build_problem(N=7, k_trans=3)
    nlp = OracleNLP(p, hessian=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = solve(ql.initial_guess(p), nlp, tol=1e-3, c_tol=1e-3, max_iter=8, backend="trust-constr")
    assert res.evals["hess"] >= 2 and 1 <= res.iterations <= 8 and np.isfinite(res.objective)


def test_solve_requires_a_sparse_structure():
    class Dense(OracleNLP):
        use_sparse_jacobian = False
    with pytest.raises(ValueError):
        solve(np.zeros(5), Dense(ql.build_problem(N=3, k_trans=2)))
