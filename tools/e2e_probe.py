"""Host-pointer (e2e) path timings: pattern x registered/unregistered x chunk size, pinned host buffers, B=4096."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadruped_landing_b200 as ql

B = int(os.environ.get("B", 4096))
devices = [int(x) for x in os.environ["DEVICES"].split(",")] if os.environ.get("DEVICES") else None
p = ql.default_problem()
rng = np.random.default_rng(0)
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((B, p.n_nlp))
Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
Zp = torch.from_numpy(Z).pin_memory().numpy()
ALLOC = os.environ.get("ALLOC", "torch")
def pinned(shape):
    return ql.host_alloc(shape) if ALLOC == "huge" else torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
print("output arrays:", ALLOC)
for pattern in ("block", "true"):
    nlp = ql.HybridNLP.from_problem(p, pattern=pattern, devices=devices)
    hout = {"f": pinned((B,)), "grad": pinned((B, nlp.n_nlp)), "g": pinned((B, nlp.m_nlp)), "jac": pinned((B, nlp.nnz_batch))}
    for reg in (False, True):
        if reg:
            t0 = time.perf_counter(); nlp.register_host_output(hout["jac"]); treg = time.perf_counter() - t0
        for chunk in [int(x) for x in os.environ.get('CHUNKS', '128,256,512').split(',')]:
            nlp.set_option("host_chunk", chunk)
            for _ in range(3):
                nlp.eval_batch_host(Zp, out=hout)
            n = 15
            per = []
            for _ in range(n):
                t0 = time.perf_counter()
                nlp.eval_batch_host(Zp, out=hout)
                per.append(time.perf_counter() - t0)
            dt = sum(per) / n
            if os.environ.get("PER_CALL"):
                print("      calls ms: " + " ".join(f"{x * 1e3:.1f}" for x in per))
            info = nlp.host_path_info()
            tt = nlp._debug_host_times()
            if not hasattr(nlp, "_tt"): nlp._tt = {k: 0.0 for k in tt}
            dtt = {k: (tt[k] - nlp._tt[k]) / (n + 3) * 1e3 for k in tt}; nlp._tt = tt
            print("      per call ms: " + ", ".join(f"{k} {v:.2f}" for k, v in dtt.items()))
            print(f"{pattern:5s} registered={int(reg)} chunk={chunk:4d}: {dt * 1e3:6.2f} ms  {B / dt / 1e3:7.1f} k evals/s  "
                  f"threads/dev {info['threads_per_device']} avx512 {info['avx512']}" + (f"  (register: {treg * 1e3:.1f} ms)" if reg else ""), flush=True)
    # what the pieces cost alone: no Jacobian (PCIe of f/grad/g only) and Jacobian only
    for want in (("f", "grad", "g"), ("jac",)):
        nlp.set_option("host_chunk", 256)
        for _ in range(2):
            nlp.eval_batch_host(Zp, out=hout, want=want)
        t0 = time.perf_counter()
        for _ in range(10):
            nlp.eval_batch_host(Zp, out=hout, want=want)
        dt = (time.perf_counter() - t0) / 10
        print(f"{pattern:5s} registered=1 want={'+'.join(want):10s}: {dt * 1e3:6.2f} ms  {B / dt / 1e3:7.1f} k evals/s", flush=True)
    if os.environ.get("THREADS_SWEEP") and pattern == "block":
        for T in (2, 4, 8, 12, 16):
            nlp.set_option("host_threads", T)
            for _ in range(2):
                nlp.eval_batch_host(Zp, out=hout)
            t0 = time.perf_counter()
            for _ in range(10):
                nlp.eval_batch_host(Zp, out=hout)
            dt = (time.perf_counter() - t0) / 10
            print(f"block registered=1 threads={T:2d}: {dt * 1e3:6.2f} ms  {B / dt / 1e3:7.1f} k evals/s", flush=True)
