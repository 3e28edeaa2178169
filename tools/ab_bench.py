"""A/B timing of kernel variants on the GPU: each variant .so is timed in its own subprocess.

    python tools/ab_bench.py build   name=DEF1,DEF2 ...     (CPU box: compiles libqlnlp_<name>.so)
    python tools/ab_bench.py run     name[@ENV=VALUE,...] ...   (GPU box: times each in its own process, prints a table)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, os, json
sys.path.insert(0, %r)
import numpy as np, torch
import quadruped_landing_b200 as ql
p = ql.default_problem()
rng = np.random.default_rng(0)
Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((4096, p.n_nlp))
Z[:, 19::20] = np.clip(Z[:, 19::20], 1e-3, 2e-2)
res = {}
CASES = [("block", ("f", "grad", "g", "jac"), "full"), ("block", ("g", "jac"), "g+J"), ("true", ("f", "grad", "g", "jac"), "TRUE"),
         ("block", ("f", "grad", "g"), "f+grad+g"), ("block", ("g",), "g")]
nlps = {pat: ql.HybridNLP.from_problem(p, pattern=pat) for pat in ("block", "true")}
for Bt in (4096, 65536):
    Zt = torch.zeros((Bt, 1216), dtype=torch.float64, device="cuda")[:, :1215]     # padded rows: TMA load path
    Zt.copy_(torch.from_numpy(Z).cuda().repeat(Bt // 4096, 1))
    # like bench.py: the inputs of consecutive launches are distinct and together larger than L2 (4 x 40 MB at B = 4,096)
    Zs = [Zt] if Bt > 4096 else [Zt] + [(Zt + 1e-6 * (j + 1)) for j in range(3)]
    Zs = [z if z.stride(0) == 1216 else torch.zeros((Bt, 1216), dtype=torch.float64, device="cuda")[:, :1215].copy_(z) for z in Zs]
    for pat, want, label in CASES:
        nlp = nlps[pat]
        out = nlp.eval_batch(Zt, want=want)
        torch.cuda.synchronize()
        for _ in range(20):
            nlp.eval_batch(Zt, out=out, want=want)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 50 if Bt == 4096 else 8
            e0.record()
            for i in range(n):
                nlp.eval_batch(Zs[i %% len(Zs)], out=out, want=want)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / n)
        res[f"{Bt // 1024}k:{label}"] = Bt / best * 1e3 / 1e6
        del out
    if Bt > 4096:
        # bench.py's C3: per-problem x0, the decision vectors rewritten by another kernel between two evaluations;
        # only the evaluator's launches are timed
        nlp = nlps["block"]
        x0 = torch.from_numpy(np.tile(p.x0, (Bt, 1))).cuda()
        noise = 1e-9 * torch.randn((Bt, 1215), device="cuda", dtype=torch.float64)
        out = nlp.eval_batch(Zt, x0=x0)
        best = 1e9
        for rep in range(3):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
            for a, b in ev:
                a.record(); nlp.eval_batch(Zt, x0=x0, out=out); b.record()
                Zt.add_(noise)
            torch.cuda.synchronize()
            best = min(best, sum(a.elapsed_time(b) for a, b in ev) / len(ev))
        res[f"{Bt // 1024}k:C3"] = Bt / best * 1e3 / 1e6
        del out, noise
print(json.dumps(res))
''' % ROOT


def main():
    mode, args = sys.argv[1], sys.argv[2:]
    if mode == "build":
        from quadruped_landing_b200 import build
        for a in args:
            name, _, defs = a.partition("=")
            print(build.build_variant(name, [d for d in defs.split(",") if d]))
    else:
        rows = []
        for name in args:                  # name[@ENV=VALUE,...]: the variant's library, with extra environment
            var, _, envs = name.partition("@")
            lib = os.path.join(ROOT, "quadruped_landing_b200", f"libqlnlp_{var}.so" if var != "default" else "libqlnlp.so")
            env = dict(os.environ, QLNLP_LIB=lib)
            env.update(kv.split("=", 1) for kv in envs.split(",") if kv)
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
            if r.returncode != 0:
                print(name, "FAILED", r.stderr[-500:])
                continue
            rows.append((name, json.loads(r.stdout.strip().splitlines()[-1])))
        keys = list(rows[0][1]) if rows else []
        print(f"{'variant':32s}" + "".join(f"{k:>12s}" for k in keys) + "   (M evals/s)")
        for name, d in rows:
            print(f"{name:32s}" + "".join(f"{d[k]:12.2f}" for k in keys))


if __name__ == "__main__":
    main()
