"""Host-side problem data for the planar-quadruped landing NLP.

Mirrors the reference's problem-definition layer (names, argument meaning and
1-based conventions where they are user visible):

* ``PlanarQuadruped``        -- src/planar_quadruped.jl:11-20
* ``QuadraticCost/LQRCost``  -- src/quadratic_cost.jl:16-42
* ``reference_trajectory``   -- src/ref_traj.jl:6-39
* ``ProblemData``            -- the fields of ``HybridNLP`` the evaluators read, src/nlp.jl:13-84
* ``default_problem`` / ``initial_guess`` -- notebook cells 2-7, src/main.ipynb:92-196

Everything here is plain numpy fp64 on the host, evaluated in the same operation
order as the Julia source so the tables handed to the device are bit-identical
to what the reference would build.  No arithmetic of the hot path lives here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

NX = 15  # state dim,   planar_quadruped.jl:25
NU = 5   # control dim, planar_quadruped.jl:26
NZK = NX + NU


@dataclass(frozen=True)
class PlanarQuadruped:
    """src/planar_quadruped.jl:11-20 (``Base.@kwdef struct PlanarQuadruped``)."""

    g: float = -9.81
    mb: float = 10.0
    mf: float = 0.1
    lb: float = 0.5
    l1: float = 0.25
    l2: float = 0.25

    def as_array(self) -> np.ndarray:
        return np.array([self.g, self.mb, self.mf, self.lb, self.l1, self.l2], dtype=np.float64)


def _half_quad(x: np.ndarray, d: np.ndarray) -> float:
    """``0.5 * x'Q * x`` for diagonal Q: dot(0.5 .* (x .* d), x), folded left (quadratic_cost.jl:38,46)."""
    t = 0.5 * (x * d)
    ret = t[0] * x[0]
    for j in range(1, len(x)):
        ret = ret + t[j] * x[j]
    return float(ret)


@dataclass(frozen=True)
class QuadraticCost:
    """src/quadratic_cost.jl:16-22; Q and R are stored as their diagonals."""

    Q: np.ndarray
    R: np.ndarray
    q: np.ndarray
    r: np.ndarray
    c: float


def LQRCost(Q, R, xf, uf=None) -> QuadraticCost:
    """src/quadratic_cost.jl:33-42.  ``Q``/``R`` may be matrices or diagonals."""
    Q = np.asarray(Q, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    Qd = np.diag(Q).copy() if Q.ndim == 2 else Q.copy()
    Rd = np.diag(R).copy() if R.ndim == 2 else R.copy()
    xf = np.asarray(xf, dtype=np.float64)
    uf = np.zeros(len(Rd)) if uf is None else np.asarray(uf, dtype=np.float64)
    q = (-Qd) * xf
    r = (-Rd) * uf
    c = _half_quad(xf, Qd) + _half_quad(uf, Rd)
    return QuadraticCost(Qd, Rd, q, r, c)


def reference_trajectory(model: PlanarQuadruped, N: int, k_trans: int, xterm, init_mode: int, dt: float
                         ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """src/ref_traj.jl:6-39.  Returns (Xref[N] of 15-vectors, Uref[N-1] of 5-vectors)."""
    g, mb = model.g, model.mb
    xterm = np.asarray(xterm, dtype=np.float64)
    Xref = np.tile(xterm.reshape(NX, 1), (1, N))
    # range(0, dt*(N-1), length=N): Julia's range is twice-precision; linspace matches to the ulp
    Xref[NX - 1, :] = np.linspace(0.0, dt * (N - 1), N)
    Uref = np.zeros((NU, N - 1))
    kt = k_trans - 1  # columns 1:k_trans-1  ->  [:kt]
    if init_mode == 1:
        Uref[1, :kt] = -mb * g
        Uref[1, kt:] = -mb * g / 2
        Uref[3, kt:] = -mb * g / 2
    else:
        Uref[3, :kt] = -mb * g
        Uref[3, kt:] = -mb * g / 2
        Uref[1, kt:] = -mb * g / 2
    Uref[4, :kt] = 0.001
    Uref[4, kt:] = 0.02
    return [Xref[:, k].copy() for k in range(N)], [Uref[:, k].copy() for k in range(N - 1)]


@dataclass
class ProblemData:
    """What ``HybridNLP`` (src/nlp.jl:13-84) carries into the evaluators.

    ``k_trans`` and ``init_mode`` keep the reference's 1-based meaning: knots
    ``k < k_trans`` are in mode ``init_mode`` (1 or 2), knots ``k >= k_trans`` in mode 3.
    Cost tables are knot-major: ``Q[k], R[k], q[k], r[k], c[k]`` for knot ``k`` (0-based row),
    the last row being the terminal cost.
    """

    model: PlanarQuadruped
    N: int
    k_trans: int
    init_mode: int
    x0: np.ndarray
    xf: np.ndarray
    Q: np.ndarray = field(repr=False)
    R: np.ndarray = field(repr=False)
    q: np.ndarray = field(repr=False)
    r: np.ndarray = field(repr=False)
    c: np.ndarray = field(repr=False)

    def __post_init__(self):
        self.x0 = np.ascontiguousarray(self.x0, dtype=np.float64)
        self.xf = np.ascontiguousarray(self.xf, dtype=np.float64)
        for name, w in (("Q", NX), ("R", NU), ("q", NX), ("r", NU)):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float64)
            if a.shape != (self.N, w):
                raise ValueError(f"{name} must have shape ({self.N}, {w}), got {a.shape}")
            setattr(self, name, a)
        self.c = np.ascontiguousarray(self.c, dtype=np.float64)
        if self.c.shape != (self.N,):
            raise ValueError("c must have shape (N,)")
        if self.x0.shape != (NX,) or self.xf.shape != (NX,):
            raise ValueError("x0/xf must have 15 entries")
        if self.N < 2:
            raise ValueError("N must be >= 2")
        if not (1 <= self.k_trans <= self.N):
            raise ValueError("k_trans must satisfy 1 <= k_trans <= N")
        if self.init_mode not in (1, 2):
            raise ValueError("init_mode must be 1 or 2")

    @classmethod
    def from_costs(cls, model, obj: Sequence[QuadraticCost], init_mode, k_trans, N, x0, xf) -> "ProblemData":
        """Same argument order as ``HybridNLP(model, obj, init_mode, k_trans, N, x0, xf)`` (nlp.jl:33-34)."""
        if len(obj) != N:
            raise ValueError("obj must hold N costs (N-1 stage costs + terminal)")
        return cls(model, int(N), int(k_trans), int(init_mode), x0, xf,
                   np.stack([o.Q for o in obj]), np.stack([o.R for o in obj]),
                   np.stack([o.q for o in obj]), np.stack([o.r for o in obj]),
                   np.array([o.c for o in obj]))

    # nlp.jl:86-87 and the closed forms derived in SURVEY.md section 8
    @property
    def n_nlp(self) -> int:
        return NX * self.N + NU * (self.N - 1)

    @property
    def m_nlp(self) -> int:
        return 18 * self.N - self.k_trans + 16

    @property
    def nnz_block(self) -> int:
        return 529 * self.N - self.k_trans - 87


def packZ(N: int, X, U) -> np.ndarray:
    """src/nlp.jl:94-102."""
    Z = np.zeros(NX * N + NU * (N - 1))
    for k in range(N - 1):
        Z[k * NZK:k * NZK + NX] = X[k]
        Z[k * NZK + NX:(k + 1) * NZK] = U[k]
    Z[(N - 1) * NZK:(N - 1) * NZK + NX] = X[N - 1]
    return Z


def unpackZ(N: int, Z):
    """src/nlp.jl:110-114."""
    Z = np.asarray(Z)
    X = [Z[k * NZK:k * NZK + NX].copy() for k in range(N)]
    U = [Z[k * NZK + NX:(k + 1) * NZK].copy() for k in range(N - 1)]
    return X, U


def default_states(model: PlanarQuadruped = PlanarQuadruped(), h_drop: float = 2.0,
                   theta0_deg: float = -30.0) -> Tuple[np.ndarray, np.ndarray]:
    """xinit / xterm of src/main.ipynb:92-93,113-132 (cells 2-3); drop height and pitch exposed for sweeps."""
    lb, l1, l2 = model.lb, model.l1, model.l2
    v_init_y = math.sqrt(2 * 9.81 * h_drop)
    xinit = np.zeros(NX)
    xinit[0] = -lb / 2.5
    xinit[1] = math.sqrt(l1 ** 2 + l2 ** 2) + 0.1
    xinit[2] = theta0_deg * math.pi / 180
    xinit[5] = -lb
    xinit[6] = 0.2
    xinit[8] = -v_init_y
    xinit[9] = -math.pi / 2
    xinit[13] = -1.0
    xterm = np.zeros(NX)
    xterm[0] = -lb / 2
    xterm[1] = math.sqrt(l1 ** 2 + l2 ** 2)
    xterm[5] = -lb
    return xinit, xterm


def build_problem(model: PlanarQuadruped = PlanarQuadruped(), N: int = 61, k_trans: int = 21,
                  init_mode: int = 1, dt: float = 0.009, xinit=None, xterm=None) -> ProblemData:
    """Cells 3-6 of the notebook (src/main.ipynb:107-171) for arbitrary (N, k_trans, init_mode)."""
    xi, xt = default_states(model)
    xinit = xi if xinit is None else np.asarray(xinit, dtype=np.float64)
    xterm = xt if xterm is None else np.asarray(xterm, dtype=np.float64)
    Xref, Uref = reference_trajectory(model, N, k_trans, xterm, init_mode, dt)
    Q = np.array([10.0] * 14 + [0.0])                      # main.ipynb:152
    R = np.array([1e-3, 1e-2, 1e-3, 1e-2, 0.0])            # main.ipynb:153
    obj = [LQRCost(Q, R, Xref[k], Uref[k]) for k in range(N - 1)]
    obj.append(LQRCost(Q, R * 0, Xref[N - 1], Uref[0]))    # main.ipynb:161
    return ProblemData.from_costs(model, obj, init_mode, k_trans, N, xinit, xterm)


def default_problem() -> ProblemData:
    """The reference instance: n=15, m=5, N=61, k_trans=21, init_mode=1, dt=0.009 (main.ipynb:107-112,126)."""
    return build_problem()


def initial_guess(prob: ProblemData, dt: float = 0.009) -> np.ndarray:
    """Z0 = packZ(nlp, Xguess, Uref), cells 7-8 (src/main.ipynb:181-196,742)."""
    N, kt = prob.N, prob.k_trans
    xinit, xterm = prob.x0, prob.xf
    X = [np.zeros(NX) for _ in range(N)]
    for k in range(1, N + 1):
        if k <= kt:
            X[k - 1] = xinit + (xterm - xinit) / (kt - 1) * (k - 1)
        else:
            X[k - 1][:14] = xterm[:14]
    for k in range(1, N):
        X[k][NX - 1] = X[k - 1][NX - 1] + (0.001 if k < kt else 0.02)
    _, Uref = reference_trajectory(prob.model, N, kt, xterm, prob.init_mode, dt)
    return packZ(N, X, Uref)


# ---- batches of problems / guesses and the reference's on-disk format -----------------
def sweep_initial_states(model: PlanarQuadruped, h_drops, theta0_degs) -> np.ndarray:
    """x0 of every problem of a drop-height x initial-pitch sweep (cf. main.ipynb:92-93,118,122)."""
    return np.stack([default_states(model, h_drop=float(h), theta0_deg=float(t))[0]
                     for h in h_drops for t in theta0_degs])


def initial_guess_batch(prob: ProblemData, x0_batch, dt: float = 0.009, xp=np):
    """Cell-7 guess (main.ipynb:181-196) for many initial states at once.  ``xp`` is numpy or torch: with torch
    tensors on the GPU the guesses are built on the device by eager tensor operations (equal to the host formula to
    1 ulp: torch divides by a scalar through its reciprocal).  ``HybridNLP.initial_guess_batch`` does the same with a
    CUDA kernel and reproduces the host formula bit for bit."""
    N, kt = prob.N, prob.k_trans
    base = initial_guess(prob, dt)                       # the class guess: everything that does not depend on x0
    if xp is np:
        x0_batch = np.asarray(x0_batch, dtype=np.float64)
        Z = np.tile(base, (x0_batch.shape[0], 1))
        xterm = prob.xf
        for k in range(1, kt + 1):                       # Xguess[k] = xinit + (xterm - xinit)/(k_trans-1)*(k-1)
            Z[:, (k - 1) * NZK:(k - 1) * NZK + 14] = (x0_batch + (xterm - x0_batch) / (kt - 1) * (k - 1))[:, :14]
        return Z
    import torch
    Z = torch.from_numpy(base).to(x0_batch.device).repeat(x0_batch.shape[0], 1)
    xterm = torch.from_numpy(prob.xf).to(x0_batch.device)
    for k in range(1, kt + 1):
        Z[:, (k - 1) * NZK:(k - 1) * NZK + 14] = (x0_batch + (xterm - x0_batch) / (kt - 1) * (k - 1))[:, :14]
    return Z


def save_solution_csv(path: str, Z) -> None:
    """`writedlm("data_6.csv", Z_sol, ',')` (main.ipynb:881): one value per line, full round-trip precision."""
    np.savetxt(path, np.asarray(Z, dtype=np.float64).reshape(-1), fmt="%.17g", delimiter=",")


def load_solution_csv(path: str) -> np.ndarray:
    """Reads a solution written by the reference or by ``save_solution_csv`` (plot_data.py:12 does the same)."""
    return np.loadtxt(path, delimiter=",").reshape(-1)


def solution_table(Z, N: int) -> np.ndarray:
    """The `reshape(-1, 20)` view the reference's plotting scripts use (plot_data.py:13-41,
    viz_temp_varying_horizon.py:22-25): row k = [x_k(15), u_k(5)], column 14 = time, 15-18 = forces, 19 = h;
    the missing last control is zero-padded."""
    Z = np.asarray(Z, dtype=np.float64).reshape(-1)
    return np.concatenate([Z, np.zeros(NU)]).reshape(N, NZK)
