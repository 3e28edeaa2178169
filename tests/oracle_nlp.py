"""Test double: the oracle behind HybridNLP's callback interface, so host-side callers (solve()) can be tested on
the CPU box.  TEST ONLY -- the product's HybridNLP has no CPU path."""
import numpy as np

from oracle.oracle import Oracle


class OracleNLP:
    use_sparse_jacobian = True

    def __init__(self, prob):
        self.prob, self.o = prob, Oracle(prob)
        self.n_nlp, self.m_nlp, self.nnz = self.o.n_nlp, self.o.m_nlp, self.o.nnz

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
