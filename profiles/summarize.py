#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launch-list CSVs into the small text summaries committed under profiles/.

    python profiles/summarize.py rep   gpurun_out/r1b_merge.ncu-rep   > profiles/r1b_merge_ncu.txt
    python profiles/summarize.py list  gpurun_out/r1b_launches.csv    > profiles/r1b_launches.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path} (one block per captured launch)")
    for vals in rows[2:]:
        print()
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:95s} {vals[i]:>22s} {units[i]}")


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        d[name][0] += 1
        d[name][1] += v
    tot = sum(v[1] for v in d.values())
    print(f"# launch list {path}: gpu__time_duration.sum per kernel (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:10.1f} {v[1] / tot:7.3f}")
    print(f"{'TOTAL':60s} {sum(v[0] for v in d.values()):8d} {tot / 1e6:10.3f}")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
