"""Timing of the SPARSE_TRUE pattern (device-resident and host-pointer paths) next to SPARSE_BLOCK."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql

p = ql.default_problem()
rng = np.random.default_rng(0)
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((4096, p.n_nlp))
Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
for pattern in ("block", "true"):
    nlp = ql.HybridNLP.from_problem(p, pattern=pattern)
    bytes_eval = 8 * (2 * nlp.n_nlp + nlp.m_nlp + 1 + nlp.nnz_batch)
    for Bt in (4096, 65536):
        Zt = torch.from_numpy(Z).cuda().repeat(Bt // 4096, 1).contiguous()
        out = nlp.eval_batch(Zt)
        for _ in range(10):
            nlp.eval_batch(Zt, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 40 if Bt == 4096 else 8
        e0.record()
        for _ in range(n):
            nlp.eval_batch(Zt, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{pattern:6s} device B={Bt:6d}: {ms:.3f} ms  {Bt / ms * 1e3 / 1e6:.2f} M evals/s  {bytes_eval * Bt / ms / 1e6:.0f} GB/s  {nlp.launch_info()}")
    Zp = torch.from_numpy(Z).pin_memory()
    hout = {"f": torch.empty(4096, dtype=torch.float64).pin_memory().numpy(),
            "grad": torch.empty((4096, nlp.n_nlp), dtype=torch.float64).pin_memory().numpy(),
            "g": torch.empty((4096, nlp.m_nlp), dtype=torch.float64).pin_memory().numpy(),
            "jac": torch.empty((4096, nlp.nnz_batch), dtype=torch.float64).pin_memory().numpy()}
    for _ in range(2):
        nlp.eval_batch_host(Zp.numpy(), out=hout)
    t0 = time.perf_counter()
    for _ in range(10):
        nlp.eval_batch_host(Zp.numpy(), out=hout)
    dt = (time.perf_counter() - t0) / 10
    print(f"{pattern:6s} host   B=  4096: {dt * 1e3:.2f} ms  {4096 / dt / 1e3:.1f} k evals/s  D2H {bytes_eval * 4096 / dt / 1e9:.1f} GB/s")
