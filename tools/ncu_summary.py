#!/usr/bin/env python3
"""Turn an `ncu --set full --import-source on` report into the markdown summary kept under
profiles/ (runs here, without a GPU: ncu only reads the report).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" samples_per_launch > profiles/xyz.md
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__block_size", "launch__grid_size",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def ncu(*args):
    # NCU_SUMMARY_KERNEL: regular expression that picks one kernel of a report with several
    k = os.environ.get("NCU_SUMMARY_KERNEL")
    extra = ["-k", "regex:" + k] if k else []
    return subprocess.run(["ncu", *extra, *args], capture_output=True, text=True).stdout


def main():
    rep, title, nsamp = sys.argv[1], sys.argv[2], float(sys.argv[3])
    src_file = sys.argv[4] if len(sys.argv) > 4 else "flacb200_kernels.cuh"
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    print(f"# {title}\n")
    print(f"Report: `{os.path.basename(rep)}` (ncu --set full --clock-control none --import-source on, one launch).\n")
    print("| metric | value |\n|---|---|")
    print(f"| kernel | `{vals[hdr.index('Kernel Name')]}` |")
    for k in RAW:
        if k in hdr:
            print(f"| {k} | {vals[hdr.index(k)]} {units[hdr.index(k)]} |")
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    src_path = os.environ.get("NCU_SUMMARY_SOURCE", os.path.join(ROOT, "flac-py_b200", "csrc", src_file))
    src = open(src_path).read().split("\n")  # must be the source the profiled binary was built from
    fn_at, cur = [None] * (len(src) + 2), None
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?(?:static )?(?:__device__|__global__|__host__)[^;{]*?\b(\w+)\s*\(", l)
        if m and not l.strip().endswith(";"):
            cur = m.group(1)
        fn_at[i] = cur
    h2, f, agg, aggi, lines = None, None, collections.Counter(), collections.Counter(), []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            f = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            h2 = r
            continue
        if h2 and len(r) > 8 and r[2] == "-" and r[0] not in ("", "Function Name"):
            try:
                i_s, i_i = h2.index("# Samples"), h2.index("Instructions Executed")
                ln = int(r[0])
                key = fn_at[ln] if f == src_file and ln < len(fn_at) else f
                agg[key] += int(r[i_s])
                aggi[key] += int(r[i_i])
                lines.append((f, ln, r[1].strip()[:100], int(r[i_s]), int(r[i_i])))
            except (ValueError, IndexError):
                pass
    ts, ti = sum(agg.values()) or 1, sum(aggi.values()) or 1
    print(f"\nThread instructions per channel-sample: {ti * 32 / nsamp:.0f}\n")
    print("| function | stall samples | instructions | thread-instr / sample |\n|---|---|---|---|")
    for k, v in agg.most_common(18):
        print(f"| {k} | {100 * v / ts:.1f} % | {100 * aggi[k] / ti:.1f} % | {aggi[k] * 32 / nsamp:.1f} |")
    print("\nHottest source lines (stall samples %, instructions %):\n\n```")
    for l in sorted(lines, key=lambda x: -x[3])[:16]:
        print(f"{l[0][:24]:24s} L{l[1]:<5d} {100 * l[3] / ts:5.1f} {100 * l[4] / ti:5.1f}  {l[2]}")
    print("```")


if __name__ == "__main__":
    main()
