"""CPU tests of the host side of the host-pointer path (csrc/hostrows.cpp): the VALS stream the kernel ships, its map
into the caller's rows, the constant image of the pattern and the row builder (full rows and registered rows, scalar
and AVX-512 writers, every row alignment, the worker pool).  The oracle supplies the expected rows; no GPU involved."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.oracle import Oracle
from quadruped_landing_b200 import HybridNLP, build_problem, initial_guess, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = [(61, 21, 1), (61, 21, 2), (31, 11, 1), (41, 14, 2), (121, 41, 1), (2, 1, 1), (2, 2, 2), (3, 2, 1),
           (33, 33, 1), (33, 1, 2), (65, 2, 2), (5, 5, 1)]


def _batch(prob, B, seed):
    rng = np.random.default_rng(seed)
    base = initial_guess(prob) if prob.k_trans > 1 else np.zeros(prob.n_nlp)
    Z = base[None, :] + 1e-2 * rng.standard_normal((B, prob.n_nlp))
    Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
    return np.ascontiguousarray(Z)


@pytest.mark.parametrize("N,kt,im", CLASSES)
@pytest.mark.parametrize("pattern", ["block", "true"])
def test_vals_stream_map_and_rows(emul_lib, N, kt, im, pattern):
    prob = build_problem(N=N, k_trans=kt, init_mode=im)
    nlp = HybridNLP.from_problem(prob, pattern=pattern)
    o = Oracle(prob)
    B = 5
    Z = _batch(prob, B, 11 * N + kt)
    rows = o.eval_batch(Z, want=("jac",), pattern=pattern)["jac"]
    pos = nlp._debug_vals_map()
    assert np.all(np.diff(pos) > 0) and pos[0] >= 0 and pos[-1] < nlp.nnz_batch
    # the kernel's VALS writer (emulated with the kernel's own headers) produces exactly the row entries at `pos`
    m = prob.model
    vals = np.empty((B, len(pos)))
    for b in range(B):
        assert emul_lib.emul_vals_stream(N, kt, im, m.g, m.mb, m.mf, m.lb, Z[b].ctypes.data, vals[b].ctypes.data, len(pos)) == 0
    assert np.array_equal(vals, rows[:, pos])
    # everything else in a row does not depend on Z
    rest = np.ones(nlp.nnz_batch, dtype=bool)
    rest[pos] = False
    assert np.all(rows[:, rest] == rows[0, rest])
    # full rows from VALS rows: bit-identical to the oracle's rows, for every row alignment (odd row stride)
    for ld in (nlp.nnz_batch, nlp.nnz_batch + 3, (nlp.nnz_batch + 8) & ~7):
        buf = np.full(B * ld + 8, np.nan)
        for shift in (0, 1, 5):
            out = buf[shift:shift + B * ld].reshape(B, ld)
            out[:] = np.nan
            nlp._debug_build_rows(vals, out[:, :nlp.nnz_batch], touched_only=False)
            assert np.array_equal(out[:, :nlp.nnz_batch], rows)
            assert np.isnan(out[:, nlp.nnz_batch:]).all()           # padding between rows untouched
    # registered rows: constant image once, then only the touched lines, three times with different Z
    ld = nlp.nnz_batch + 1
    out = np.full((B, ld), np.nan)
    nlp._debug_build_rows(np.zeros_like(vals), out[:, :nlp.nnz_batch], touched_only=False)
    for rep in range(3):
        Zr = _batch(prob, B, 1000 + rep)
        want = o.eval_batch(Zr, want=("jac",), pattern=pattern)["jac"]
        nlp._debug_build_rows(np.ascontiguousarray(want[:, pos]), out[:, :nlp.nnz_batch], touched_only=True)
        assert np.array_equal(out[:, :nlp.nnz_batch], want)
    assert np.isnan(out[:, nlp.nnz_batch:]).all()


def test_default_instance_counts():
    nlp = HybridNLP.from_problem(build_problem())
    info = nlp.host_path_info()
    assert info["pcie_jac_doubles_per_eval"] == 2794 == len(nlp._debug_vals_map())
    assert info["row_doubles"] == 32161 and info["lines_per_row"] == 4021
    assert info["touched_lines_per_row"] in (1701, 1702)       # + the partial line at the end of the row


def test_pool_builds_the_same_rows():
    prob = build_problem()
    nlp = HybridNLP.from_problem(prob)
    o = Oracle(prob)
    B = 97
    rows = o.eval_batch(_batch(prob, B, 3), want=("jac",))["jac"]
    vals = np.ascontiguousarray(rows[:, nlp._debug_vals_map()])
    for threads in (2, 5, -2, -4):            # negative: workers only, asynchronously (what the host pipeline does)
        out = np.full((B, nlp.nnz_batch), np.nan)
        for _ in range(3):                       # back-to-back jobs on the persistent pool
            nlp._debug_build_rows(vals, out, touched_only=False, threads=threads)
        assert np.array_equal(out, rows)


def test_scalar_writer_matches_the_avx512_one():
    """QLNLP_NO_AVX512=1 forces the portable line writer; run it in a fresh process and compare."""
    code = ("import numpy as np, sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from test_hostrows_cpu import _batch\n"
            "from oracle.oracle import Oracle\n"
            "from quadruped_landing_b200 import HybridNLP, build_problem\n"
            "p = build_problem(N=41, k_trans=14, init_mode=2); nlp = HybridNLP.from_problem(p)\n"
            "assert nlp.host_path_info()['avx512'] == 0\n"
            "rows = Oracle(p).eval_batch(_batch(p, 4, 1), want=('jac',))['jac']\n"
            "vals = np.ascontiguousarray(rows[:, nlp._debug_vals_map()])\n"
            "out = np.full((4, nlp.nnz_batch + 1), np.nan)\n"
            "nlp._debug_build_rows(vals, out[:, :-1], touched_only=False)\n"
            "assert np.array_equal(out[:, :-1], rows)\n"
            "nlp._debug_build_rows(vals, out[:, :-1], touched_only=True)\n"
            "assert np.array_equal(out[:, :-1], rows)\n" % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, QLNLP_NO_AVX512="1")
    subprocess.run([sys.executable, "-c", code], check=True, env=env)


def test_shard_bounds_match_the_python_host_logic():
    import ctypes as C
    from quadruped_landing_b200 import shard_bounds
    L = load_library()
    for B in (0, 1, 7, 4096, 65536, 1 << 20):
        for n in (1, 2, 3, 8):
            for s in range(n):
                lo, hi = C.c_int64(), C.c_int64()
                assert L.qlnlp_shard_bounds(B, n, s, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == shard_bounds(B, n, s)
