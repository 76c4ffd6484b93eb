#!/usr/bin/env python
"""Read a .ncu-rep here (no GPU): key metrics per captured launch + top stall-sample instructions."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:70]
        print("==", name)
        for k in KEYS:
            if k in hdr:
                print(f"   {k:82s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")


def source(path, top=22, launch=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    b = blocks[launch]
    hdr = b["hdr"]
    ia, isamp, iex, ith = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    data = [(int(r[isamp] or 0), r[ia].strip(), int(r[iex] or 0), i, r[ith]) for i, r in enumerate(b["rows"]) if len(r) > isamp and r[isamp].isdigit()]
    tot = sum(d[0] for d in data) or 1
    print(f"-- source page, launch {launch}: {b['name'][:60]}: {tot} samples, {len(data)} SASS instructions, "
          f"{sum(d[2] for d in data)} warp-instructions executed")
    for s, src, ex, i, th in sorted(data, reverse=True)[:top]:
        print(f"{s / tot * 100:5.1f}%  #{i:4d} exec={ex:9d} thr={th:>5}  {src[:96]}")


if __name__ == "__main__":
    p = sys.argv[1]
    raw(p)
    source(p, launch=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
