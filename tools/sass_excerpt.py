#!/usr/bin/env python3
"""profiles/r02_sass_excerpt.md from the built library: per-kernel opcode counts and two excerpts (cuobjdump -sass, -res-usage).
    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.md"""
import os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quadruped_landing_b200", "libqlnlp.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = dict(re.findall(r"Function (\S+):\n\s*REG:(\d+)", res))
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?;)", line)
    if m and cur:
        funcs[cur].append((m.group(1), m.group(2).strip()))
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("void ", "").replace("(ql::Launch)", "").replace("(ql::HessLaunch)", "").replace("(int)", "").replace("(bool)", "")
OPS = ["UBLKCP", "SYNCS", "UBLKPF", "PREEXIT", "ACQBULK", "DADD", "DMUL", "DFMA", "STS", "LDS", "STG", "LDG"]
print("# SASS evidence (round 2, final build) -- `cuobjdump -sass quadruped_landing_b200/libqlnlp.so`, sm_100a cubin only\n")
print("Regenerate with `python tools/sass_excerpt.py`.  Demangled: `ql::eval_kernel<JM, FASTDIV, RAGGED>`; JM 0 = no Jacobian, 1 = SPARSE_BLOCK,")
print("2 = SPARSE_TRUE, 3 = VALS (host path); `ql::hess_kernel<FASTDIV>` = Lagrangian Hessian.  `UBLKCP` = TMA bulk copy (`cp.async.bulk`),")
print("`SYNCS` = mbarrier operations, `UBLKPF` = bulk L2 prefetch, `PREEXIT` / `ACQBULK` = `griddepcontrol.launch_dependents` / `.wait`")
print("(programmatic dependent launch); the RK4 / dual arithmetic is DADD / DMUL (un-fused by construction), DFMA")
print("appears only inside the exact reciprocal division and sincos.\n")
print("| kernel | instructions | KB | " + " | ".join(OPS) + " | registers |")
print("|---|---|---|" + "---|" * (len(OPS) + 1))
for name, ins in funcs.items():
    cnt = [sum(1 for _, t in ins if re.search(r"(^|\s)" + op + r"\b", t) or re.search(r"(^|\s)" + op + r"\.", t)) for op in OPS]
    print(f"| `{demangle(name)}` | {len(ins)} | {len(ins) * 16 / 1024:.1f} | " + " | ".join(map(str, cnt)) + f" | {regs.get(name, '?')} |")
key = next(n for n in funcs if "eval_kernelILi1ELb1ELb0" in n)
ins = funcs[key]
i = next(j for j, (_, t) in enumerate(ins) if t.startswith("UBLKCP.G.S"))
print("\n## The segment store of `ql::eval_kernel<1, 1, 0>` (SPARSE_BLOCK stream): TMA bulk store shared -> global, commit\n```")
for a, t in ins[max(0, i - 6):i + 3]:
    print(f"/*{a}*/  {t}")
print("```\n\n## The staged load of a decision vector: mbarrier init / expect-tx / TMA bulk load global -> shared, and the wait on it\n```")
for a, t in ins:
    if "SYNCS" in t or t.startswith("UBLKCP.S.G") or "UBLKCP.S" in t:
        print(f"/*{a}*/  {t}")
print("```\n\n## The row stores: streaming (SPARSE_BLOCK kernel) and L1::no_allocate (compact-output kernels)\n```")
for k in ("eval_kernelILi1ELb1ELb0", "eval_kernelILi0ELb1ELb0"):
    name = next(n for n in funcs if k in n)
    kinds = sorted({t.split()[0] if not t.startswith("@") else t.split()[1] for _, t in funcs[name] if re.search(r"\bSTG", t)})
    print(f"{demangle(name)}: {', '.join(kinds)}")
print("```")
