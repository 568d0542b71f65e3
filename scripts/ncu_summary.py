#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed): headline counters the north star names
(FP32/FP64 pipe and issue utilisation, warp execution efficiency, L1/L2 hit rates, DRAM bytes)
plus a per-SASS-region table of issued instructions / active threads / stall samples.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--regions 50] > profiles/xyz.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    step = int(sys.argv[sys.argv.index("--regions") + 1]) if "--regions" in sys.argv else 50
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        print(f"## {name}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {row[i]} | {units[i]} |")
        print()
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    h = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if not h:
        return
    hd = src[h[0]]
    data = [r for r in src[h[0] + 1:] if len(r) == len(hd)]
    ci, ct, cs, so = hd.index("Instructions Executed"), hd.index("Thread Instructions Executed"), hd.index("# Samples"), hd.index("Source")
    ti = sum(int(r[ci]) for r in data); tt = sum(int(r[ct]) for r in data); ts = sum(int(r[cs]) for r in data)
    print(f"SASS instructions: {len(data)}; warp-instructions executed {ti}; average active threads {tt / ti:.2f} / 32; stall samples {ts}\n")

    def op(s):
        p = s.split()
        o = p[1] if p[0].startswith("@") else p[0]
        return o.split(".")[0]
    print("| SASS range | % of issued | avg active threads | % of stall samples | dominant opcodes |\n|---|---|---|---|---|")
    for k in range(0, len(data), step):
        seg = data[k:k + step]
        i = sum(int(r[ci]) for r in seg); t = sum(int(r[ct]) for r in seg); s = sum(int(r[cs]) for r in seg)
        ops = ", ".join(f"{a}x{b}" for a, b in collections.Counter(op(r[so]) for r in seg).most_common(4))
        print(f"| {k}-{k + len(seg) - 1} | {i / ti * 100:.1f} | {t / max(i, 1):.1f} | {s / max(ts, 1) * 100:.1f} | {ops} |")


if __name__ == "__main__":
    main()
