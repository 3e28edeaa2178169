"""Pins the CPU oracle to every number the reference records for the hot path.

The reference has no test suite; what pins results is the recorded notebook run (src/main.ipynb)
on its shipped solution src/data_6.csv (committed here as tests/golden/ref_solutions.npz).
"""
import numpy as np

from oracle.oracle import Oracle
from quadruped_landing_b200 import default_problem, initial_guess


def test_dimensions_match_ipopt_log():
    o = Oracle(default_problem())
    assert o.n_nlp == 1215                      # main.ipynb:221 "Total number of variables"
    assert o.m_nlp == 1032 + 61                 # main.ipynb:225-226
    # main.ipynb:217-218: dense structure, 1253880 + 74115 nonzeros = m*n
    assert o.m_nlp * o.n_nlp == 1253880 + 74115
    assert 1032 * 1215 == 1253880 and 61 * 1215 == 74115
    assert o.nnz == 529 * 61 - 21 - 87 == 32161


def test_objective_on_recorded_solution(golden):
    o = Oracle(default_problem())
    # main.ipynb:710  "Objective...............:   1.1608112892558562e+02"
    assert o.eval_f(golden["data_6"]) == 1.1608112892558562e+02


def test_constraint_violation_on_recorded_solution(golden):
    o = Oracle(default_problem())
    c = o.eval_c(golden["data_6"])
    # main.ipynb:712  "Constraint violation....:   1.4928675395736724e-06" (max over equality rows)
    assert np.abs(c[:1032]).max() == 1.4928675395736724e-06
    assert np.abs(c[:1032]).argmax() + 1 == 329          # dynamics block k=20, component 15
    assert c[1032:].min() >= 0.0                         # body clearance rows are feasible


def test_recorded_boundary_residuals(golden):
    p = default_problem()
    z = golden["data_6"]
    # cells 9-10, main.ipynb:749-814: Z_sol[1:15]-xinit and Z_sol[end-14:end]-xterm
    d0 = z[:15] - p.x0
    assert d0[3] == -1.7424461934630155e-9 and d0[0] == -3.4916514124461173e-14
    d1 = z[-15:] - p.xf
    assert d1[14] == 0.848539898959304 and d1[0] == -7.227551890309769e-14
    # cell 11, main.ipynb:826-828
    assert z[-19] == 44.56221189408092 and z[-17] == 53.53778810591909
    assert z[-19] + z[-17] == 98.10000000000001
    c = Oracle(p).eval_c(z)
    assert np.array_equal(c[:15], d0) and np.array_equal(c[15:29], d1[:14])


def test_initial_guess_matches_iteration_zero():
    p = default_problem()
    o = Oracle(p)
    z0 = initial_guess(p)
    # main.ipynb:232 iter 0: inf_pr 3.13e-01 (the objective column is after Ipopt's bound push: soft pin only)
    assert abs(np.abs(o.eval_c(z0)[:1032]).max() - 0.313) < 5e-4
    assert abs(o.eval_f(z0) - 1.5438467869136336) < 1e-12


def test_structure_is_column_major_filter_of_assignments(golden):
    o = Oracle(default_problem())
    rows, cols = o.jacobian_structure()
    lin = (cols - 1) * o.m_nlp + (rows - 1)
    assert np.all(np.diff(lin) > 0)                      # column-major, row fastest, no duplicates (moi.jl:31-33)
    z = golden["data_6"]
    dense = o.jac_c_dense(z)
    vals = o.jac_c_sparse(z)
    assert not np.isnan(vals).any()
    assert np.array_equal(dense[rows - 1, cols - 1], vals)
    mask = np.zeros_like(dense, dtype=bool)
    mask[rows - 1, cols - 1] = True
    assert not dense[~mask].any()                        # nothing is assigned outside the pattern


def test_all_golden_solutions_satisfy_the_dynamics(golden):
    """data_1..5 are solutions of other initial conditions (different drop speed / pitch), so only the
    terminal rows and the dynamics defects are expected to vanish on them."""
    o = Oracle(default_problem())
    for i in range(1, 7):
        c = o.eval_c(golden[f"data_{i}"])
        assert np.abs(c[15:29]).max() < 1e-6
        assert np.abs(c[29:929]).max() < 1e-5
