"""Quick on-GPU diagnostic: parity summary per output + rough kernel timing.  Not a benchmark."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import quadruped_landing_b200 as ql
from oracle.oracle import Oracle

p = ql.default_problem()
nlp = ql.HybridNLP.from_problem(p)
o = Oracle(p)
rng = np.random.default_rng(0)
B = 512
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((B, p.n_nlp))
Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
Zd = torch.from_numpy(Z).cuda()
ref = o.eval_batch(Z)
rows, cols = nlp.jacobian_structure_arrays()
for label, outspec in (("bulk", None), ("plain", "odd")):
    out = {}
    if outspec == "odd":
        out["jac"] = torch.full((B, nlp.nnz_block), float("nan"), dtype=torch.float64, device="cuda")
    res = nlp.eval_batch(Zd, out=out)
    torch.cuda.synchronize()
    print(label, nlp.launch_info())
    for k in ("f", "grad", "g", "jac"):
        got = res[k].cpu().numpy()
        bad = ~(np.abs(got - ref[k]) <= 1e-14 + 1e-12 * np.abs(ref[k]))
        exact = np.array_equal(got, ref[k])
        print(f"  {k}: mismatches {bad.sum()} / {bad.size}  bit-exact {exact}  nan {np.isnan(got).sum()}")
        if bad.any():
            idx = np.argwhere(bad)[:8]
            for ii in idx:
                ii = tuple(ii)
                extra = (rows[ii[1]], cols[ii[1]]) if k == "jac" else ""
                print("     ", ii, got[ii], ref[k][ii], extra)
for Bt in (4096, 65536):
    Zt = Zd.repeat(Bt // B, 1).contiguous()
    out = nlp.eval_batch(Zt)
    torch.cuda.synchronize()
    for want in (("f", "grad", "g", "jac"), ("g", "jac"), ("f", "grad", "g")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            nlp.eval_batch(Zt, out=out, want=want)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            nlp.eval_batch(Zt, out=out, want=want)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = Bt * (9720 + (8744 if "g" in want else 0) + (9728 if "grad" in want else 0) + (257288 if "jac" in want else 0))
        print(f"B={Bt} want={want}: {ms:.3f} ms  {Bt / ms * 1e3:.3e} eval/s  {nbytes / ms / 1e6:.1f} GB/s")
