"""One process, two handles on two GPUs (the C ABI binds each handle to its device)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql
from oracle.oracle import Oracle
p = ql.default_problem()
o = Oracle(p)
rng = np.random.default_rng(0)
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((300, p.n_nlp))
ref = o.eval_batch(Z)
nlps = [ql.HybridNLP.from_problem(p, device=d) for d in range(torch.cuda.device_count())]
outs = [nlp.eval_batch(torch.from_numpy(Z).to(f"cuda:{d}")) for d, nlp in enumerate(nlps)]
for d in range(len(nlps)):
    torch.cuda.synchronize(d)
for d, out in enumerate(outs):
    ok = all(np.all(np.abs(out[k].cpu().numpy() - ref[k]) <= 1e-14 + 1e-12 * np.abs(ref[k])) for k in ref)
    print(f"cuda:{d}", "ok" if ok else "MISMATCH", out["jac"].device)
h = [nlp.eval_batch_host(Z[:100]) for nlp in nlps]
print("host paths equal:", all(np.array_equal(h[0]["jac"], x["jac"]) for x in h))
