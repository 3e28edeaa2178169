"""A few launches of ONE kernel instantiation, for ncu (and plain timing without it).

    python tools/ncu_target.py --pattern block|true --want f,grad,g,jac --B 65536 [--launches 3]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql

ap = argparse.ArgumentParser()
ap.add_argument("--pattern", default="block")
ap.add_argument("--want", default="f,grad,g,jac")
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--launches", type=int, default=3)
a = ap.parse_args()
p = ql.default_problem()
nlp = ql.HybridNLP.from_problem(p, pattern=a.pattern)
rng = np.random.default_rng(0)
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((4096, p.n_nlp))
Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
Zt = torch.zeros((a.B, 1216), dtype=torch.float64, device="cuda")[:, :1215]
Zt.copy_(torch.from_numpy(Z).cuda().repeat((a.B + 4095) // 4096, 1)[:a.B])
want = tuple(a.want.split(","))
out = nlp.eval_batch(Zt, want=want)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.launches):
    nlp.eval_batch(Zt, want=want, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.launches
nb = 8 * (nlp.n_nlp + ("grad" in want) * nlp.n_nlp + ("g" in want) * nlp.m_nlp + ("f" in want) + ("jac" in want) * nlp.nnz_batch)
print(f"{a.pattern} {a.want} B={a.B}: {ms:.4f} ms  {a.B / ms / 1e3:.2f} M evals/s  {nb * a.B / ms / 1e6:.0f} GB/s  {nlp.launch_info()}")
