"""Small case for compute-sanitizer: default class (bulk + plain store paths, per-evaluation x0) and an odd class."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql

for (N, kt, im, B) in [(61, 21, 1, 40), (33, 33, 2, 9), (5, 2, 1, 3)]:
    p = ql.build_problem(N=N, k_trans=kt, init_mode=im)
    nlp = ql.HybridNLP.from_problem(p)
    rng = np.random.default_rng(0)
    Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((B, p.n_nlp))
    Zd = torch.from_numpy(Z).cuda()
    x0 = torch.from_numpy(np.tile(p.x0, (B, 1))).cuda()
    out = nlp.eval_batch(Zd, x0=x0)
    jac = torch.empty((B, nlp.nnz_block), dtype=torch.float64, device="cuda")
    out2 = nlp.eval_batch(Zd, want=("jac",), out={"jac": jac})          # unaligned rows: plain store path
    torch.cuda.synchronize()
    assert torch.equal(out["jac"], out2["jac"])
    h = nlp.eval_batch_host(Z)
    assert np.array_equal(h["jac"], out["jac"].cpu().numpy())
print("sanitize case ok")
