#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (.ncu-rep) or an `ncu --metrics gpu__time_duration.sum --csv` launch list into the
text kept under profiles/.   python tools/ncu_summary.py gpurun_out/x.ncu-rep | gpurun_out/launches.csv"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"ncu --set full --clock-control none --import-source on, {path.split('/')[-1]}: {len(rows) - 2} profiled launch(es)")
    for r in rows[2:]:
        print(f"\n{r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for m in METRICS:
            if m in col:
                print(f"  {m:95s} {r[col[m]]:>18s} {units[col[m]]}")


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 14 and r[0].isdigit()]
    by = defaultdict(list)
    for r in rows:
        by[r[4].split("(")[0].replace("void ", "")].append(float(r[14]))
    total = sum(sum(v) for v in by.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none, {path.split('/')[-1]}: {len(rows)} launches, {total / 1e6:.3f} ms of kernel time")
    for k, v in sorted(by.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k:40s} x{len(v):3d}  avg {sum(v) / len(v) / 1e3:10.1f} us   share {100 * sum(v) / total:6.2f} %")
    f = [x for k, v in by.items() if "force_kernel" in k for x in v]
    i = [x for k, v in by.items() if "integrate" in k for x in v]
    if f and i:
        print(f"  one step = force + integrate: force share {100 * (sum(f) / len(f)) / (sum(f) / len(f) + sum(i) / len(i)):.2f} %")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (report if p.endswith(".ncu-rep") else launches)(p)
