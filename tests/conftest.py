import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

# fp64 parity tolerance stated by BASELINE.json's north_star: 1e-12 relative / 1e-14 absolute
RTOL = 1e-12
ATOL = 1e-14


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def assert_parity(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want)
    tol = ATOL + RTOL * np.abs(want)
    bad = ~(err <= tol)
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} entries outside |a-b| <= 1e-14 + 1e-12|b|; "
                             f"first {idx.tolist()} got {got[bad][:5]} want {want[bad][:5]}")


def _compile_native(name):
    src = os.path.join(ROOT, "tests", "native", name + ".cpp")
    out_dir = os.path.join(ROOT, "tests", "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".so")
    csrc = os.path.join(ROOT, "quadruped_landing_b200", "csrc")
    deps = [src, os.path.join(ROOT, "tests", "native", "host_consts.h")] + \
           [os.path.join(csrc, h) for h in ("layout.h", "rk4_dual_gen.h", "true_run.h")]
    if (not os.path.exists(out)) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src], check=True)
    return out


@pytest.fixture(scope="session")
def rk4_host_lib():
    import ctypes as C
    L = C.CDLL(_compile_native("rk4_gen_host"))
    dp = C.c_void_p
    L.host_rk4_jac.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
    L.host_rk4.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, dp]
    return L


@pytest.fixture(scope="session")
def emul_lib():
    import ctypes as C
    L = C.CDLL(_compile_native("emul_host"))
    L.emul_jac_stream.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.emul_true_stream.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_void_p, C.c_void_p, C.c_int]
    L.emul_vals_stream.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_void_p, C.c_void_p, C.c_int]
    L.emul_hess_stream.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_void_p] * 5 + [C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    L.emul_run_off.argtypes = [C.c_int] * 4
    L.emul_rk4_pos.argtypes = [C.c_int] * 6
    return L


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_solutions.npz")))


def perturbed_batch(prob, bases, B, sigma, seed):
    """Z_b = base[b mod len] + sigma*xi_b, h entries clipped to [1e-3, 2e-2]."""
    rng = np.random.default_rng(seed)
    n = prob.n_nlp
    Z = np.stack([bases[b % len(bases)] for b in range(B)]) + sigma * rng.standard_normal((B, n))
    Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
    return np.ascontiguousarray(Z)
