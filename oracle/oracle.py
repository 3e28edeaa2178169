"""ctypes front end of the CPU oracle (oracle/ql_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the status header in oracle/ql_oracle.h.  Only
tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke()
import this module, and only to CHECK the CUDA path or to time the CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libql_oracle.so")
_lib = None


class _Model(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("g", "mb", "mf", "lb", "l1", "l2")]


class _Problem(C.Structure):
    _fields_ = [("N", C.c_int64), ("k_trans", C.c_int64), ("init_mode", C.c_int64),
                ("model", _Model), ("x0", C.c_double * 15), ("xf", C.c_double * 15),
                ("Q", C.c_void_p), ("R", C.c_void_p), ("q", C.c_void_p), ("r", C.c_void_p), ("c", C.c_void_p),
                ("kinematics", C.c_int64)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O3 -ffp-contract=off -fopenmp)."""
    srcs = [os.path.join(_HERE, f) for f in ("ql_oracle.c", "ql_oracle_hess.c", "ql_oracle.h", "ql_dyn_impl.inc", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P = C.POINTER(_Problem)
        dp = C.c_void_p
        L.qlo_num_primals.restype = C.c_int64
        L.qlo_num_primals.argtypes = [P]
        L.qlo_num_duals.restype = C.c_int64
        L.qlo_num_duals.argtypes = [P]
        L.qlo_nnz_block.restype = C.c_int64
        L.qlo_nnz_block.argtypes = [P]
        L.qlo_lqr_cost.restype = None
        L.qlo_lqr_cost.argtypes = [dp] * 7
        L.qlo_eval_f.restype = C.c_double
        L.qlo_eval_f.argtypes = [P, dp]
        for name in ("qlo_grad_f", "qlo_eval_c", "qlo_jac_c_dense"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [P, dp, dp]
        for name in ("qlo_constraint_bounds", "qlo_variable_bounds", "qlo_jacobian_structure"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [P, dp, dp]
        L.qlo_plan_create.restype = C.c_void_p
        L.qlo_plan_create.argtypes = [P]
        L.qlo_plan_create_true.restype = C.c_void_p
        L.qlo_plan_create_true.argtypes = [P]
        L.qlo_plan_nnz.restype = C.c_int64
        L.qlo_plan_nnz.argtypes = [C.c_void_p]
        L.qlo_plan_structure.restype = None
        L.qlo_plan_structure.argtypes = [C.c_void_p, dp, dp]
        L.qlo_plan_destroy.restype = None
        L.qlo_plan_destroy.argtypes = [C.c_void_p]
        L.qlo_jac_c_sparse.restype = None
        L.qlo_jac_c_sparse.argtypes = [C.c_void_p, P, dp, dp]
        L.qlo_eval_batch.restype = None
        L.qlo_eval_batch.argtypes = [C.c_void_p, P, C.c_int64, dp, C.c_int64, dp, dp,
                                     dp, dp, C.c_int64, dp, C.c_int64, dp, C.c_int64, C.c_int]
        L.qlo_max_threads.restype = C.c_int
        L.qlo_rk4.restype = None
        L.qlo_rk4.argtypes = [C.POINTER(_Model), C.c_int, dp, dp, dp]
        L.qlo_rk4_jacobian.restype = None
        L.qlo_rk4_jacobian.argtypes = [C.POINTER(_Model), C.c_int, dp, dp, dp, dp]
        L.qlo_rk4_hessian.restype = None
        L.qlo_rk4_hessian.argtypes = [C.POINTER(_Model), C.c_int, dp, dp, dp, dp]
        L.qlo_hess_lagrangian_dense.restype = None
        L.qlo_hess_lagrangian_dense.argtypes = [P, dp, C.c_double, dp, dp]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _model_struct(model) -> _Model:
    return _Model(model.g, model.mb, model.mf, model.lb, model.l1, model.l2)


def lqr_cost(Qd, Rd, xf, uf):
    Qd, Rd, xf, uf = _f64(Qd), _f64(Rd), _f64(xf), _f64(uf)
    q, r, c = np.empty(15), np.empty(5), np.empty(1)
    lib().qlo_lqr_cost(_ptr(Qd), _ptr(Rd), _ptr(xf), _ptr(uf), _ptr(q), _ptr(r), _ptr(c))
    return q, r, float(c[0])


def rk4(model, mode: int, x, u) -> np.ndarray:
    x, u = _f64(x), _f64(u)
    xn = np.empty(15)
    m = _model_struct(model)
    lib().qlo_rk4(C.byref(m), mode, _ptr(x), _ptr(u), _ptr(xn))
    return xn


def rk4_jacobian(model, mode: int, x, u):
    """Returns (xn[15], J[15,20]) with J[i,j] = d xn_i / d [x;u]_j."""
    x, u = _f64(x), _f64(u)
    xn, J = np.empty(15), np.empty(300)
    m = _model_struct(model)
    lib().qlo_rk4_jacobian(C.byref(m), mode, _ptr(x), _ptr(u), _ptr(xn), _ptr(J))
    return xn, J.reshape(20, 15).T.copy()


def rk4_hessian(model, mode: int, x, u, lam) -> np.ndarray:
    """H[20,20] = sum_r lam[r] * Hess(rk4_r) w.r.t. [x;u] (dense second-order forward mode)."""
    x, u, lam = _f64(x), _f64(u), _f64(lam)
    H = np.empty(400)
    m = _model_struct(model)
    lib().qlo_rk4_hessian(C.byref(m), mode, _ptr(x), _ptr(u), _ptr(lam), _ptr(H))
    return H.reshape(20, 20).T.copy()


class Oracle:
    """CPU evaluator of one problem (same seven entry points as src/moi.jl:1-33)."""

    def __init__(self, prob, kinematics: bool = False):
        """``kinematics=True`` switches on the leg-length rows the reference keeps commented out (nlp.jl:60,70;
        constraints.jl:115-138,276-288): 2N more constraints, 8N more Jacobian entries."""
        self.prob = prob
        self.kinematics = bool(kinematics)
        self._keep = [_f64(prob.Q), _f64(prob.R), _f64(prob.q), _f64(prob.r), _f64(prob.c)]
        s = _Problem()
        s.N, s.k_trans, s.init_mode = prob.N, prob.k_trans, prob.init_mode
        s.model = _model_struct(prob.model)
        for i in range(15):
            s.x0[i] = float(prob.x0[i])
            s.xf[i] = float(prob.xf[i])
        s.Q, s.R, s.q, s.r, s.c = [a.ctypes.data for a in self._keep]
        s.kinematics = 1 if kinematics else 0
        self._s = s
        self._p = C.byref(s)
        L = lib()
        self.n_nlp = int(L.qlo_num_primals(self._p))
        self.m_nlp = int(L.qlo_num_duals(self._p))
        self._plan = None
        self._plan_true = None
        self._nnz = None

    def __del__(self):
        for name in ("_plan", "_plan_true"):
            if getattr(self, name, None):
                lib().qlo_plan_destroy(getattr(self, name))
                setattr(self, name, None)

    # --- SPARSE_TRUE pattern (structural non-zeros only)
    @property
    def plan_true(self):
        if self._plan_true is None:
            self._plan_true = lib().qlo_plan_create_true(self._p)
        return self._plan_true

    @property
    def nnz_true(self) -> int:
        return int(lib().qlo_plan_nnz(self.plan_true))

    def jacobian_structure_true(self):
        rows = np.empty(self.nnz_true, dtype=np.int64)
        cols = np.empty(self.nnz_true, dtype=np.int64)
        lib().qlo_plan_structure(self.plan_true, _ptr(rows), _ptr(cols))
        return rows, cols

    def jac_c_sparse_true(self, Z) -> np.ndarray:
        Z = _f64(Z)
        out = np.full(self.nnz_true, np.nan)
        lib().qlo_jac_c_sparse(self.plan_true, self._p, _ptr(Z), _ptr(out))
        return out

    @property
    def plan(self):
        if self._plan is None:
            self._plan = lib().qlo_plan_create(self._p)
        return self._plan

    @property
    def nnz(self) -> int:
        if self._nnz is None:
            self._nnz = int(lib().qlo_nnz_block(self._p))
        return self._nnz

    # --- single evaluations
    def eval_f(self, Z) -> float:
        Z = _f64(Z)
        return float(lib().qlo_eval_f(self._p, _ptr(Z)))

    def grad_f(self, Z) -> np.ndarray:
        Z = _f64(Z)
        out = np.empty(self.n_nlp)
        lib().qlo_grad_f(self._p, _ptr(Z), _ptr(out))
        return out

    def eval_c(self, Z) -> np.ndarray:
        Z = _f64(Z)
        out = np.empty(self.m_nlp)
        lib().qlo_eval_c(self._p, _ptr(Z), _ptr(out))
        return out

    def jac_c_dense(self, Z) -> np.ndarray:
        """m_nlp x n_nlp matrix (Fortran order in memory, like the reference's reshape(vec, m, n))."""
        Z = _f64(Z)
        out = np.zeros((self.n_nlp, self.m_nlp))  # column-major m x n == C-order n x m
        lib().qlo_jac_c_dense(self._p, _ptr(Z), _ptr(out))
        return out.T

    def jac_c_sparse(self, Z) -> np.ndarray:
        Z = _f64(Z)
        out = np.full(self.nnz, np.nan)
        lib().qlo_jac_c_sparse(self.plan, self._p, _ptr(Z), _ptr(out))
        return out

    def hess_lagrangian_dense(self, Z, sigma: float, lam) -> np.ndarray:
        """sigma * Hess f + sum_r lam_r Hess g_r as a dense symmetric [n_nlp, n_nlp] matrix (no reference target)."""
        Z, lam = _f64(Z), _f64(lam)
        n = self.n_nlp
        H = np.empty(n * n)
        lib().qlo_hess_lagrangian_dense(self._p, _ptr(Z), float(sigma), _ptr(lam), _ptr(H))
        return H.reshape(n, n).T.copy()

    def jacobian_structure(self):
        rows = np.empty(self.nnz, dtype=np.int64)
        cols = np.empty(self.nnz, dtype=np.int64)
        lib().qlo_jacobian_structure(self._p, _ptr(rows), _ptr(cols))
        return rows, cols

    def constraint_bounds(self):
        lb, ub = np.empty(self.m_nlp), np.empty(self.m_nlp)
        lib().qlo_constraint_bounds(self._p, _ptr(lb), _ptr(ub))
        return lb, ub

    def variable_bounds(self):
        xl, xu = np.empty(self.n_nlp), np.empty(self.n_nlp)
        lib().qlo_variable_bounds(self._p, _ptr(xl), _ptr(xu))
        return xl, xu

    # --- batches
    def eval_batch(self, Z, x0=None, xf=None, want=("f", "grad", "g", "jac"), nthreads: int = 0, pattern="block"):
        Z = _f64(Z)
        B = Z.shape[0]
        assert Z.shape[1] == self.n_nlp
        x0 = None if x0 is None else _f64(x0)
        xf = None if xf is None else _f64(xf)
        f = np.empty(B) if "f" in want else None
        grad = np.empty((B, self.n_nlp)) if "grad" in want else None
        g = np.empty((B, self.m_nlp)) if "g" in want else None
        plan, nnz = (self.plan, self.nnz) if pattern == "block" else (self.plan_true, self.nnz_true)
        jac = np.empty((B, nnz)) if "jac" in want else None
        lib().qlo_eval_batch(plan, self._p, B, _ptr(Z), self.n_nlp, _ptr(x0), _ptr(xf),
                             _ptr(f), _ptr(grad), self.n_nlp, _ptr(g), self.m_nlp,
                             _ptr(jac), nnz, int(nthreads))
        return {"f": f, "grad": grad, "g": g, "jac": jac}

    @staticmethod
    def max_threads() -> int:
        return int(lib().qlo_max_threads())
