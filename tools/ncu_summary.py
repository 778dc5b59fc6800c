#!/usr/bin/env python
"""Summarise ncu output for profiles/: (1) a launch list CSV (--metrics gpu__time_duration.sum) into
per-kernel totals and shares, (2) a --set full .ncu-rep into the handful of metrics DESIGN.md cites.

    python tools/ncu_summary.py launches gpurun_out/launches.csv
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict


def launches(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v for _, v in tot.values())
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for name, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:8d} {v:12.1f} {v / n:10.2f} {100 * v / total:6.1f}%")
    print(f"{'TOTAL':70s} {sum(n for n, _ in tot.values()):8d} {total:12.1f}")


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_blocks", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in data:
        d = {"kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", ""),
             "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k in KEYS:
            if k in col:
                d[k] = f"{r[col[k]]} {units[col[k]]}".strip()
        res.append(d)
    for i, d in enumerate(res):
        print(f"--- launch {i}")
        for k, v in d.items():
            print(f"{k:80s} {v}")
    return res


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
