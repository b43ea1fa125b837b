#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (captured with --set full --import-source on): headline raw
metrics, the hottest source lines (instructions, stall samples, shared-memory wavefronts) and the
SASS opcode mix.  Usage: ncu_summary.py report.ncu-rep kernel_substring [title]"""
import collections
import os
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else pat
    print(f"# {title}")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    h, u = rows[0], rows[1]
    kcol = h.index("Kernel Name")
    for v in rows[2:]:
        if pat in v[kcol]:
            for k in RAW:
                if k in h:
                    print(f"{k:72s} {v[h.index(k)]:>18s} {u[h.index(k)]}")
            break
    out = ncu(rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + pat)
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, lines, ops = None, None, [], collections.Counter()
    stall = collections.Counter()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 10:
            d = dict(zip(hdr, r))
            try:
                n, s = int(d["Instructions Executed"]), int(d["# Samples"])
            except (KeyError, ValueError):
                continue
            if r[0] != "":
                lines.append((cur, int(r[0]), r[1].strip()[:90], n, s, int(d.get("L1 Wavefronts Shared", 0) or 0)))
            else:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
                if m:
                    op = m.group(2)
                    ops[".".join(op.split(".")[:2]) if op.startswith(("LD", "ST")) else op.split(".")[0]] += n
                for k, val in d.items():
                    if k.startswith("stall_") and "Not Issued" not in k:
                        try:
                            stall[k] += int(val)
                        except ValueError:
                            pass
    ti, ts, tw = sum(x[3] for x in lines) or 1, sum(x[4] for x in lines) or 1, sum(x[5] for x in lines) or 1
    print(f"\nsource lines: {ti:.3e} warp instructions, {ts} stall samples, {tw:.3e} shared-memory wavefronts")
    key = 4 if os.environ.get('NCU_SORT') == 'smp' else 3
    for x in sorted(lines, key=lambda x: -x[key])[:int(os.environ.get('NCU_TOP', 14))]:
        print(f" {100 * x[3] / ti:5.1f}% inst {100 * x[4] / ts:5.1f}% smp {100 * x[5] / tw:5.1f}% smem  {x[0]}:{x[1]}  {x[2]}")
    to = sum(ops.values()) or 1
    print("\nopcode mix: " + ", ".join(f"{k} {100 * v / to:.1f}%" for k, v in ops.most_common(14)))
    tt = sum(stall.values()) or 1
    print("stalls: " + ", ".join(f"{k[6:]} {100 * v / tt:.0f}%" for k, v in stall.most_common(6)))


if __name__ == "__main__":
    main()
