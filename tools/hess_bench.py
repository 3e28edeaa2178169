"""Throughput of the batched Lagrangian-Hessian kernel (default instance, 3,355 values per evaluation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql
p = ql.default_problem(); nlp = ql.HybridNLP.from_problem(p, hessian=True)
rng = np.random.default_rng(0)
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((4096, p.n_nlp))
Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
for B in (4096, 65536):
    Zd = torch.zeros((B, 1216), dtype=torch.float64, device="cuda")[:, :1215]
    Zd.copy_(torch.from_numpy(Z).cuda().repeat(B // 4096, 1))
    mu = torch.randn((B, nlp.m_nlp + 1), dtype=torch.float64, device="cuda")[:, :nlp.m_nlp]
    sig = torch.rand(B, dtype=torch.float64, device="cuda")
    H = nlp.eval_hessian_batch(Zd, mu, sig)
    for _ in range(5): nlp.eval_hessian_batch(Zd, mu, sig, out=H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): nlp.eval_hessian_batch(Zd, mu, sig, out=H)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nbytes = 8 * (nlp.n_nlp + nlp.m_nlp + 1 + nlp.nnz_hess)
    print(f"Hessian B={B}: {ms:.4f} ms  {B / ms / 1e3:.2f} M evals/s  {nbytes * B / ms / 1e6:.0f} GB/s ({nbytes} B/eval)")
