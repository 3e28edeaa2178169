#!/usr/bin/env python3
"""Aggregate the source page of an .ncu-rep by CUDA source line: samples, executed instructions, top stalls.
    python tools/ncu_lines.py rep [top=40]        (first profiled launch only)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows, hdr, path, seen, skip = [], None, None, set(), False
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]; skip = path in seen; seen.add(path); continue      # later launches repeat the files
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 10 and r[0] != "" and not skip: rows.append((path, r))
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
iconf = hdr.index("L1 Wavefronts Shared Excessive")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
num = lambda s: int(s) if s.strip().lstrip("-").isdigit() else 0
tot = sum(num(r[isamp]) for _, r in rows); totex = sum(num(r[iex]) for _, r in rows)
print(f"total samples {tot}  warp instructions executed {totex}")
byfile = {}
for p, r in rows:
    a = byfile.setdefault(p, [0, 0, 0]); a[0] += num(r[isamp]); a[1] += num(r[iex]); a[2] += num(r[iconf])
for p, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  file {p:28s} {a[0] / tot:6.3f} samp {a[1] / totex:6.3f} exec  excess smem wavefronts {a[2]}")
for p, r in sorted(rows, key=lambda pr: -num(pr[1][isamp]))[:top]:
    st = sorted(((num(r[i]), h[6:]) for i, h in stalls), reverse=True)[:3]
    print(f"{num(r[isamp]) / tot:6.3f} samp {num(r[iex]) / totex:6.3f} exec  xs_wf {r[iconf]:>9s}  {p}:{r[0]}  {r[1].strip()[:80]}   [{', '.join(f'{n}:{v}' for v, n in st if v)}]")
