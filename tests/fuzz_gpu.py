"""Randomised stress test of the CUDA path against the oracle: random problem classes, batch sizes, leading
dimensions / alignments, output subsets, both sparse patterns, device and host entry points (plain and registered
output rows), the MOI callbacks with their x cache, the opt-in kinematic rows, the Lagrangian Hessian, canary rows.
    python tests/fuzz_gpu.py [seconds] [seed]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql
from oracle.oracle import Oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
SENT = -123.456
t0 = time.time()
ncase = nfail = 0


def close(a, b):
    return np.all(np.abs(a - b) <= 1e-14 + 1e-12 * np.abs(b))


while time.time() - t0 < budget:
    N = int(rng.choice([2, 3, 5, 9, 31, 32, 33, 40, 61, 64, 65, 96, 97, 121]))
    kt = int(rng.integers(1, N + 1))
    im = int(rng.integers(1, 3))
    pattern = str(rng.choice(["block", "true"]))
    kin = bool(rng.random() < 0.2)
    p = ql.build_problem(N=N, k_trans=kt, init_mode=im)
    nlp = ql.HybridNLP.from_problem(p, pattern=pattern, kinematics=kin)
    o = Oracle(p, kinematics=kin)
    reg = None                       # a registered host output array, reused across the repetitions
    for rep in range(3):
        B = int(rng.choice([1, 2, 7, 31, 64, 65, 200, 513, 1500, 3000]))
        Z = rng.normal(size=(B, p.n_nlp))
        Z[:, 19::20] = rng.uniform(1e-3, 2e-2, size=(B, N - 1))
        want = tuple(k for k in ("f", "grad", "g", "jac") if rng.random() < 0.75) or ("g",)
        use_x0 = rng.random() < 0.3
        x0 = rng.normal(size=(B, 15)) if use_x0 else None
        ref = o.eval_batch(Z, x0=x0, want=want, pattern=pattern)
        widths = {"grad": nlp.n_nlp, "g": nlp.m_nlp, "jac": nlp.nnz_batch}
        mode = str(rng.choice(["device", "host", "host-registered", "callbacks", "hessian"]))
        if kin and mode in ("host-registered", "hessian"):
            mode = "host"
        ncase += 1
        try:
            if mode == "device":
                zpad = int(rng.integers(0, 3))
                Zd = torch.full((B, p.n_nlp + zpad), SENT, dtype=torch.float64, device="cuda")
                Zd[:, :p.n_nlp] = torch.from_numpy(Z)
                out, big = {}, {}
                for k in want:
                    if k == "f":
                        continue
                    pad = int(rng.integers(0, 4))
                    big[k] = torch.full((B + 2, widths[k] + pad), SENT, dtype=torch.float64, device="cuda")
                    out[k] = big[k][1:B + 1, :widths[k]]
                kw = {"x0": torch.from_numpy(x0).cuda()} if use_x0 else {}
                res = nlp.eval_batch(Zd[:, :p.n_nlp], want=want, out=out, **kw)
                torch.cuda.synchronize()
                for k, t in big.items():
                    assert bool((t[0] == SENT).all()) and bool((t[-1] == SENT).all()) and bool((t[:, widths[k]:] == SENT).all()), f"canary {k}"
                got = {k: res[k].cpu().numpy() for k in want}
            elif mode == "host":
                got = nlp.eval_batch_host(Z, x0=x0, want=want)
            elif mode == "host-registered":
                if reg is None:
                    reg = np.full((3000, nlp.nnz_batch), np.nan)
                    nlp.register_host_output(reg)
                lo = int(rng.integers(0, 3000 - B + 1))
                got = nlp.eval_batch_host(Z, x0=x0, want=want, out={"jac": reg[lo:lo + B]} if "jac" in want else None)
            elif mode == "callbacks":
                # the four MOI callbacks in random order at a few vectors, some of them repeated (x cache)
                want, got, refc = ("f", "grad", "g", "jac"), {k: [] for k in ("f", "grad", "g", "jac")}, {k: [] for k in ("f", "grad", "g", "jac")}
                for b in rng.integers(0, B, size=4):
                    z = Z[b].copy()
                    r1 = o.eval_batch(z[None, :], pattern=pattern)
                    for k in rng.permutation(["f", "grad", "g", "jac", "f", "jac"]):
                        if k == "f":
                            v = np.array(nlp.eval_objective(z))
                        else:
                            v = np.empty({"grad": nlp.n_nlp, "g": nlp.m_nlp, "jac": nlp.nnz}[k])
                            getattr(nlp, {"grad": "eval_objective_gradient", "g": "eval_constraint", "jac": "eval_constraint_jacobian"}[k])(v, z)
                        got[k].append(v)
                        refc[k].append(r1[k][0])
                got = {k: np.array(v) for k, v in got.items()}
                ref = {k: np.array(v) for k, v in refc.items()}
            else:
                want = ("hess",)
                nb = min(B, 8)
                mu = rng.normal(size=(nb, nlp.m_nlp))
                sg = rng.uniform(0.1, 2.0, size=nb)
                H = nlp.eval_hessian_batch(torch.from_numpy(Z[:nb].copy()).cuda(), torch.from_numpy(mu).cuda(), torch.from_numpy(sg).cuda())
                torch.cuda.synchronize()
                rows, cols = nlp.hessian_structure_arrays()
                Hd = np.stack([o.hess_lagrangian_dense(Z[b], sg[b], mu[b])[rows - 1, cols - 1] for b in range(nb)])
                scale = max(1.0, np.abs(Hd).max())
                got, ref = {"hess": H.cpu().numpy() / scale}, {"hess": Hd / scale}      # tolerance relative to the block's scale
                assert np.abs(got["hess"] - ref["hess"]).max() <= 1e-12, f"hess mismatch: {np.abs(got['hess'] - ref['hess']).max()}"
                got, ref = {"hess": ref["hess"]}, ref
            for k in want:
                assert got[k].shape == ref[k].shape, (k, got[k].shape, ref[k].shape)
                assert close(got[k], ref[k]), f"{k} mismatch: {np.abs(got[k] - ref[k]).max()}"
        except AssertionError as e:
            nfail += 1
            print(f"FAIL N={N} kt={kt} im={im} pattern={pattern} B={B} want={want} mode={mode} x0={use_x0}: {e}")
    del nlp
print(f"fuzz: {ncase} cases, {nfail} failures, {time.time() - t0:.0f} s")
sys.exit(1 if nfail else 0)
