"""The host row builder alone (no GPU work): rows/s and GB/s written vs threads, for output buffers in ordinary memory,
torch-pinned memory and (when available) qlnlp_host_alloc memory."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quadruped_landing_b200 as ql

B = 4096
p = ql.default_problem()
nlp = ql.HybridNLP.from_problem(p)
nv = nlp.host_path_info()["pcie_jac_doubles_per_eval"]
vals = np.random.default_rng(0).standard_normal((B, nv))
bufs = {"numpy": np.empty((B, nlp.nnz_block))}
try:
    import torch
    bufs["torch-pinned"] = torch.empty((B, nlp.nnz_block), dtype=torch.float64).pin_memory().numpy()
    vals_p = torch.from_numpy(vals).pin_memory().numpy()
except Exception as e:
    print("no torch pinned memory:", e); vals_p = vals
if hasattr(ql, "host_alloc"):
    try:
        bufs["qlnlp_host_alloc"] = ql.host_alloc((B, nlp.nnz_block))
    except Exception as e:
        print("host_alloc failed:", e)
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| lines/row touched", nlp.host_path_info()["touched_lines_per_row"])
for name, out in bufs.items():
    out[:] = 0
    for touched in (False, True):
        for T in (1, 2, 4, 8, 16):
            if T > len(os.sched_getaffinity(0)): continue
            nlp._debug_build_rows(vals_p, out, touched, threads=T)
            t0 = time.perf_counter()
            for _ in range(5):
                nlp._debug_build_rows(vals_p, out, touched, threads=T if T > 1 else 1)
            dt = (time.perf_counter() - t0) / 5
            lines = nlp.host_path_info()["touched_lines_per_row"] if touched else nlp.host_path_info()["lines_per_row"]
            print(f"{name:18s} touched={int(touched)} threads={T:2d}: {dt * 1e3:7.2f} ms  {B / dt / 1e3:8.1f} k rows/s  {B * lines * 64 / dt / 1e9:6.1f} GB/s written", flush=True)
