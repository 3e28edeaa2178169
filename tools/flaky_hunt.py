import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql
p = ql.default_problem()
nlp = ql.HybridNLP.from_problem(p)
rng = np.random.default_rng(99)
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((1100, p.n_nlp))
Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
dev = {k: v.cpu().numpy() for k, v in nlp.eval_batch(torch.from_numpy(Z).cuda()).items()}
torch.cuda.synchronize()
bad = 0
for rep in range(40):
    for B in (1, 63, 64, 512, 513, 1100):
        out = {"jac": np.full((B, nlp.nnz_block), np.nan)}
        host = nlp.eval_batch_host(Z[:B], out=out)
        for k in ("f", "grad", "g", "jac"):
            if not np.array_equal(host[k], dev[k][:B]):
                d = host[k] != dev[k][:B]
                rows = np.unique(np.argwhere(d)[:, 0])
                print(f"rep {rep} B={B} {k}: {d.sum()} mismatches in rows {rows[:10]} (n rows {len(rows)}) nan {np.isnan(host[k]).sum()}",
                      "first cols", np.argwhere(d)[:5].tolist())
                bad += 1
print("bad", bad)
