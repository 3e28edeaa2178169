#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, on the CPU box) into profiles/<name>.md + traffic.json.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_eval_kernel_block [--traffic] [--what "command that was profiled"]
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_config_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * f[unit]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    what = sys.argv[sys.argv.index("--what") + 1] if "--what" in sys.argv else \
        "bench.py --steps 3 --warmup 3 --no-extras --no-cpu (B=4096 full evaluation per launch)"
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, launches = raw[0], raw[1], raw[2:]
    md = [f"# ncu summary of `{rep}`", "",
          "Captured with `ncu --set full --clock-control none --import-source on -k regex:eval_kernel` on a B200 "
          f"under `{what}`. "
          "Times under ncu are cold-cache and serialised; use the SHARES and per-launch counters.", ""]
    traffic = []
    for n, r in enumerate(launches):
        md.append(f"## launch {n}")
        md.append("| metric | value | unit |")
        md.append("|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                md.append(f"| {k} | {r[i]} | {units[i]} |")
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        t = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
        traffic.append(t)
        md.append(f"| dram read+write per launch | {t:.0f} | byte |")
        md.append("")
    # stall breakdown + opcode mix from the SASS source page of the first launch
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv"))))
    hdr2, rows = None, []
    nk = 0
    for r in src:
        if r and r[0] == "Kernel Name":
            nk += 1
        elif r and r[0] == "Address":
            hdr2 = r
        elif nk == 1 and hdr2 and len(r) > 10:
            rows.append(r)
    if rows:
        isamp, iex, isrc = hdr2.index("# Samples"), hdr2.index("Instructions Executed"), hdr2.index("Source")
        stalls = [h for h in hdr2 if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[isamp] or 0) for r in rows)
        totex = sum(int(r[iex] or 0) for r in rows)
        md += ["## warp-stall samples (first launch)", f"total samples {tot}, warp instructions executed {totex}", "",
               "| reason | samples | share |", "|---|---|---|"]
        agg = {h: sum(int(r[hdr2.index(h)] or 0) for r in rows) for h in stalls}
        for h, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            if v:
                md.append(f"| {h} | {v} | {v / tot:.3f} |")
        op = defaultdict(lambda: [0, 0])
        for r in rows:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
            o = m.group(2) if m else "?"
            op[o][0] += int(r[isamp] or 0)
            op[o][1] += int(r[iex] or 0)
        md += ["", "## opcode mix (first launch)", "| opcode | samples share | executed share |", "|---|---|---|"]
        for o, (a, b) in sorted(op.items(), key=lambda kv: -kv[1][1])[:22]:
            md.append(f"| {o} | {a / tot:.3f} | {b / totex:.3f} |")
        md.append("")
    open(out + ".md", "w").write("\n".join(md) + "\n")
    if "--traffic" in sys.argv:
        json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic),
                   "source": f"profiles/{out.split('/')[-1]}.md (dram__bytes_read.sum + dram__bytes_write.sum, "
                             "ncu --set full, B=4096 full evaluation)"},
                  open("profiles/traffic.json", "w"), indent=1)
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
