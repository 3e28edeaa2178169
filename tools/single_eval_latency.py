"""Latency of the four MOI callbacks at batch 1 (the drop-in use behind Ipopt: one decision vector per call, host
pointers).  The reference's recorded run averages ~43 ms per f+grad+g+J bundle (src/main.ipynb:717-725)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quadruped_landing_b200 as ql

p = ql.default_problem()
z = ql.initial_guess(p)
for label, kw in (("SPARSE_BLOCK", dict(pattern="block")), ("SPARSE_TRUE", dict(pattern="true")),
                  ("DENSE (reference structure)", dict(use_sparse_jacobian=False))):
    nlp = ql.HybridNLP.from_problem(p, **kw)
    grad, g, vals = np.empty(nlp.n_nlp), np.empty(nlp.m_nlp), np.empty(nlp.nnz)
    calls = {"eval_objective": lambda: nlp.eval_objective(z),
             "eval_objective_gradient": lambda: nlp.eval_objective_gradient(grad, z),
             "eval_constraint": lambda: nlp.eval_constraint(g, z),
             "eval_constraint_jacobian": lambda: nlp.eval_constraint_jacobian(vals, z)}
    out = []
    for name, fn in calls.items():
        for _ in range(20):
            fn()
        n = 200 if "jacobian" not in name or label != "DENSE (reference structure)" else 30
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        out.append(f"{name} {(time.perf_counter() - t0) / n * 1e6:.0f} us")
    print(f"{label:28s}: " + ", ".join(out))
