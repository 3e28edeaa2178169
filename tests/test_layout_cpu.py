"""CPU checks of everything integer in the CUDA path: closed-form structure, run offsets, the segment
plan, and -- through tests/native/emul_host.cpp, which replays the kernel's staging-buffer logic with the
kernel's own headers -- the position of every Jacobian value the kernel will write."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from quadruped_landing_b200 import HybridNLP, build_problem, initial_guess

CLASSES = [(61, 21, 1), (61, 21, 2), (31, 11, 1), (41, 14, 2), (81, 27, 1), (101, 34, 2), (121, 41, 1),
           (2, 1, 1), (2, 2, 2), (3, 2, 1), (33, 33, 1), (33, 1, 2), (64, 32, 1), (65, 2, 2), (5, 5, 1), (32, 16, 2)]


def _z(prob, seed):
    rng = np.random.default_rng(seed)
    base = initial_guess(prob) if prob.k_trans > 1 else np.zeros(prob.n_nlp)
    Z = base + 1e-2 * rng.standard_normal(prob.n_nlp)
    Z[19::20] = np.clip(Z[19::20], 1e-3, 2e-2)
    return Z


@pytest.mark.parametrize("N,kt,im", CLASSES)
def test_structure_bit_exact(N, kt, im):
    prob = build_problem(N=N, k_trans=kt, init_mode=im)
    nlp = HybridNLP.from_problem(prob, use_sparse_jacobian=True)
    o = Oracle(prob)
    assert (nlp.n_nlp, nlp.m_nlp, nlp.nnz) == (o.n_nlp, o.m_nlp, o.nnz)
    assert nlp.n_nlp == 20 * N - 5 and nlp.m_nlp == 18 * N - kt + 16 and nlp.nnz == 529 * N - kt - 87
    r, c = nlp.jacobian_structure_arrays()
    r0, c0 = o.jacobian_structure()
    assert r.dtype == np.int64 and np.array_equal(r, r0) and np.array_equal(c, c0)


def test_dense_structure_is_the_references_grid():
    prob = build_problem(N=5, k_trans=3)
    nlp = HybridNLP.from_problem(prob, use_sparse_jacobian=False)
    st = nlp.jacobian_structure()
    m, n = nlp.m_nlp, nlp.n_nlp
    assert len(st) == m * n == nlp.nnz
    # vec(Tuple.(CartesianIndices(zeros(m, n)))), moi.jl:31-33
    assert st[0] == (1, 1) and st[1] == (2, 1) and st[m] == (1, 2) and st[-1] == (m, n)


@pytest.mark.parametrize("N,kt,im", CLASSES)
def test_segment_plan_and_value_positions(emul_lib, N, kt, im):
    prob = build_problem(N=N, k_trans=kt, init_mode=im)
    nlp = HybridNLP.from_problem(prob)
    segs = nlp._debug_segments()
    # the plan tiles the stream, never crosses a 32-knot pass and fits the staging buffers
    assert segs[0, 2] == 0 and segs[-1, 3] == nlp.nnz and np.all(segs[1:, 2] == segs[:-1, 3])
    assert np.all((segs[:, 0] - 1) // 32 == (segs[:, 0] + segs[:, 1] - 2) // 32)
    assert np.all(segs[:, 3] - (segs[:, 2] & ~1) <= 1064)
    assert set(np.unique(segs[:, 5])) <= {0, 1}
    for k in range(1, N + 1):
        assert emul_lib.emul_run_off(N, kt, im, k) == (np.nonzero(nlp.jacobian_structure_arrays()[1] > 20 * (k - 1))[0][0])
    Z = _z(prob, N + kt)
    out = np.empty(nlp.nnz)
    m = prob.model
    rc = emul_lib.emul_jac_stream(N, kt, im, m.g, m.mb, m.mf, m.lb, segs.ctypes.data, len(segs),
                                  Z.ctypes.data, out.ctypes.data, 1)
    assert rc == 0
    ref = Oracle(prob).jac_c_sparse(Z)
    assert np.array_equal(out, ref)          # bit-exact, including with persisted templates (2nd repetition)


@pytest.mark.parametrize("N,kt,im", CLASSES)
def test_sparse_true_structure_and_stream(emul_lib, N, kt, im):
    """SPARSE_TRUE: closed-form structure == the oracle's numerically determined non-zero pattern, and the run
    writer the kernel uses (csrc/true_run.h) reproduces the oracle's values bit for bit."""
    prob = build_problem(N=N, k_trans=kt, init_mode=im)
    nlp = HybridNLP.from_problem(prob, pattern="true")
    o = Oracle(prob)
    assert nlp.nnz == o.nnz_true == nlp.nnz_batch and nlp.nnz_block == o.nnz
    r, c = nlp.jacobian_structure_arrays()
    r0, c0 = o.jacobian_structure_true()
    assert np.array_equal(r, r0) and np.array_equal(c, c0)
    # it is a sub-sequence of SPARSE_BLOCK in the same (column-major) order
    rb, cb = o.jacobian_structure()
    lin, linb = (c - 1) * o.m_nlp + r, (cb - 1) * o.m_nlp + rb
    assert np.all(np.diff(lin) > 0) and np.isin(lin, linb).all()
    Z = _z(prob, 3 * N + kt)
    out = np.empty(o.nnz_true)
    m = prob.model
    assert emul_lib.emul_true_stream(N, kt, im, m.g, m.mb, m.mf, m.lb, Z.ctypes.data, out.ctypes.data, o.nnz_true) == 0
    assert np.array_equal(out, o.jac_c_sparse_true(Z))
    # and equals the SPARSE_BLOCK values at those positions; everything SPARSE_BLOCK adds is exactly zero
    vb = o.jac_c_sparse(Z)
    sel = np.isin(linb, lin)
    assert np.array_equal(vb[sel], out) and not vb[~sel].any()


def test_default_instance_true_pattern_count():
    """15 + 71 pattern entries + extras per initial-mode knot, 56 pattern entries at the jump knot, 57 in mode 3:
    4,840 structural non-zeros at the default instance (SURVEY.md 8a quotes 4,842 / 58; the jump mask removes
    15 of the 71 entries, not 13 -- checked against the oracle above)."""
    nlp = HybridNLP.from_problem(build_problem(), pattern="true")
    assert nlp.nnz == 4840
