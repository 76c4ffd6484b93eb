#!/usr/bin/env python
"""profiles/ncu_kernels.json from one `ncu --set full --import-source on` capture of a few bench steps (read here, no
GPU): per pipeline stage the DRAM bytes, warp-instructions and duration PER STEP, and for the compositor the numbers
of its issue-slot roofline (lanes per instruction, FMA-pipe and issue utilisation, share of the stall samples at the
batch barrier, evaluated (pixel, Gaussian) pairs).  bench.py reads the file into `roofline` and `stages[*].dram_*`.

    python tools/ncu_kernels.py gpurun_out/<tag>/prof_step.ncu-rep [profiles/ncu_kernels.json]

Launches are grouped by (kernel, grid size): every group occurs once per step (the background and the link span of
project / emit have different grids, the five sort classes different block sizes), so a stage's per-step figure is
the sum over its groups of the per-launch average."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

STAGE_OF = [("project_", "project"), ("scan_", "scan"), ("work_", "scan"), ("emit_kernel", "emit"), ("tile_sort_kernel", "sort"),
            ("merge_", "sort"), ("bg_reach_kernel", "sort"), ("composite_kernel", "composite"), ("copy_static_tiles", "composite"), ("ns_", "post"),
            ("snapshot_kernel", "post"), ("link_pose_kernel", "pose"), ("fk_kernel", "pose")]


def stage_of(name):
    for key, st in STAGE_OF:
        if key in name:
            return st
    return "other"


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return 0.0


def main(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, k):
        return num(r[col[k]]) if k in col else 0.0

    def scaled(r, k, target):  # value of a byte / time metric in `target` units
        v, u = get(r, k), units[col[k]] if k in col else ""
        mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        return v * mul

    groups = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        key = (name, r[col["Grid Size"]] if "Grid Size" in col else "", r[col["Block Size"]] if "Block Size" in col else "")
        g = groups.setdefault(key, collections.defaultdict(float))
        g["n"] += 1
        g["dram"] += scaled(r, "dram__bytes_read.sum", "byte") + scaled(r, "dram__bytes_write.sum", "byte")
        g["inst"] += get(r, "sm__inst_executed.sum") or get(r, "smsp__inst_executed.sum")
        g["us"] += scaled(r, "gpu__time_duration.sum", "us")
        g["lanes"] += get(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
        g["fma"] += get(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")
        g["issue"] += get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        g["warps"] += get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
        g["regs"] = get(r, "launch__registers_per_thread")
    stages = collections.OrderedDict()
    kernels = []
    for (name, grid, blk), g in groups.items():
        n = g["n"]
        st = stage_of(name)
        s = stages.setdefault(st, collections.defaultdict(float))
        s["dram_bytes"] += g["dram"] / n
        s["inst_executed"] += g["inst"] / n
        s["time_us_ncu"] += g["us"] / n
        kernels.append({"kernel": name, "grid": grid, "block": blk, "stage": st, "launches_captured": int(n), "time_us": round(g["us"] / n, 2),
                        "dram_bytes": int(g["dram"] / n), "inst_executed": int(g["inst"] / n), "lanes_per_inst": round(g["lanes"] / n, 2),
                        "fma_pipe_pct": round(g["fma"] / n, 1), "issue_active_pct": round(g["issue"] / n, 1),
                        "warps_active_pct": round(g["warps"] / n, 1), "registers": int(g["regs"])})
    res = {k: {kk: (int(vv) if kk != "time_us_ncu" else round(vv, 2)) for kk, vv in v.items()} for k, v in stages.items()}
    # compositor details: the dominant composite_kernel group + its source page
    comp = max((k for k in kernels if "composite_kernel" in k["kernel"]), key=lambda k: k["time_us"], default=None)
    if comp:
        c = res["composite"]
        c.update({"lanes_per_inst": comp["lanes_per_inst"], "fma_pipe_pct": comp["fma_pipe_pct"], "issue_active_pct": comp["issue_active_pct"],
                  "registers": comp["registers"], "kernel": comp["kernel"]})
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        blocks, cur = [], None
        for r in srows:
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1], "hdr": None, "rows": []}
                blocks.append(cur)
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = r
            elif cur is not None:
                cur["rows"].append(r)
        for b in blocks:
            if "composite_kernel" not in b["name"] or not b["hdr"]:
                continue
            h = b["hdr"]
            ia, isamp, iex, ith = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
            tot = bar = 0
            pairs = 0.0
            for r in b["rows"]:
                if len(r) <= isamp or not r[isamp].isdigit():
                    continue
                tot += int(r[isamp])
                if "B2R" in r[ia] or "BAR." in r[ia]:
                    bar += int(r[isamp])
                if "0x40b1999a" in r[ia]:      # the sigma <= 5.55 range test: one per evaluated (pixel, Gaussian) pair
                    pairs += int(r[iex] or 0) * num(r[ith])
            c["barrier_stall_share_pct"] = round(100.0 * bar / max(tot, 1), 1)
            c["evaluated_pairs"] = int(pairs)
            break
    res["_kernels"] = kernels
    res["_source"] = f"tools/ncu_kernels.py {path}: ncu --set full --clock-control none -s 60 -c 26, python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs; per-step sums of per-launch averages by (kernel, grid)"
    json.dump(res, open(out, "w"), indent=1)
    for k, v in res.items():
        if not k.startswith("_"):
            print(k, v)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "profiles/ncu_kernels.json")
