"""Test double: the oracle behind HybridNLP's callback interface, so host-side callers (solve()) can be tested on
the CPU box.  TEST ONLY -- the product's HybridNLP has no CPU path."""
import numpy as np

from oracle.oracle import Oracle


class OracleNLP:
    use_sparse_jacobian = True

    def __init__(self, prob, hessian=False):
        self.prob, self.o = prob, Oracle(prob)
        self.n_nlp, self.m_nlp, self.nnz = self.o.n_nlp, self.o.m_nlp, self.o.nnz
        self.hessian = hessian
        if hessian:      # block-diagonal lower triangle, dense within each knot's block (a superset of the product's pattern)
            r, c = np.nonzero(np.tril(np.kron(np.eye(prob.N), np.ones((20, 20)))[:self.n_nlp, :self.n_nlp]))
            order = np.lexsort((r, c))
            self._hr, self._hc = r[order] + 1, c[order] + 1

    def features_available(self):
        return ["Grad", "Jac", "Hess"] if self.hessian else ["Grad", "Jac"]

    def hessian_structure_arrays(self):
        return self._hr, self._hc

    def eval_hessian_lagrangian(self, H, x, sigma, mu):
        H[:] = self.o.hess_lagrangian_dense(x, sigma, mu)[self._hr - 1, self._hc - 1]

    def jacobian_structure_arrays(self):
        return self.o.jacobian_structure()

    def variable_bounds(self):
        return self.o.variable_bounds()

    def constraint_bounds(self):
        return self.o.constraint_bounds()

    def eval_objective(self, x):
        return self.o.eval_f(x)

    def eval_objective_gradient(self, grad, x):
        grad[:] = self.o.grad_f(x)

    def eval_constraint(self, g, x):
        g[:] = self.o.eval_c(x)

    def eval_constraint_jacobian(self, vec, x):
        vec[:] = self.o.jac_c_sparse(x)
