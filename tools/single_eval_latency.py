"""Latency of the MOI callbacks at batch 1 (the drop-in use behind Ipopt: one decision vector per call, host pointers).
Every bundle uses a NEW x (so the first callback pays the launch) followed by the other three callbacks at the same x
(served from the x cache).  The reference's recorded run averages ~43 ms per f+grad+g+J bundle (src/main.ipynb:717-725)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quadruped_landing_b200 as ql

p = ql.default_problem()
z0 = ql.initial_guess(p)
rng = np.random.default_rng(0)
X = z0[None, :] + 1e-3 * rng.standard_normal((256, p.n_nlp))
for label, kw in (("SPARSE_BLOCK", dict(pattern="block")), ("SPARSE_TRUE", dict(pattern="true")),
                  ("DENSE (reference structure)", dict(use_sparse_jacobian=False))):
    nlp = ql.HybridNLP.from_problem(p, **kw)
    f1, grad, g, vals = np.empty(1), np.empty(nlp.n_nlp), np.empty(nlp.m_nlp), np.empty(nlp.nnz)
    for cache in (1, 0):
        nlp.set_option("x_cache", cache)
        t = {"eval_objective (new x: H2D + launch + D2H)": 0.0, "eval_objective_gradient": 0.0, "eval_constraint": 0.0,
             "eval_constraint_jacobian": 0.0}
        n = 200 if "DENSE" not in label else 30
        for it in range(-20, n):
            x = X[it % 256]
            c = []
            t0 = time.perf_counter(); nlp.eval_objective(x); c.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); nlp.eval_objective_gradient(grad, x); c.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); nlp.eval_constraint(g, x); c.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); nlp.eval_constraint_jacobian(vals, x); c.append(time.perf_counter() - t0)
            if it >= 0:
                for k, v in zip(t, c):
                    t[k] += v
        tot = sum(t.values()) / n * 1e6
        print(f"{label:28s} x_cache={cache}: " + ", ".join(f"{k} {v / n * 1e6:.1f} us" for k, v in t.items()) + f"  | bundle {tot:.1f} us")
    if "DENSE" not in label:
        t0 = time.perf_counter()
        for it in range(200):
            nlp.eval_all(X[it % 256], f1, grad, g, vals)
        print(f"{label:28s} qlnlp_eval_all (one call per iterate): {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
