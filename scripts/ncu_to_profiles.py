#!/usr/bin/env python
"""Condense an ncu --set full report into tracked evidence: one small CSV per kernel launch under profiles/ and the
per-kernel DRAM traffic table bench.py reads (profiles/traffic.json).
  python scripts/ncu_to_profiles.py <report.ncu-rep> <tag> [<workload>]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
]

rep, tag = sys.argv[1], sys.argv[2]
workload = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen = {}
traffic = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("gx::", "").strip()
    short = re.sub(r"[^A-Za-z0-9_]+", "_", name).strip("_")
    n = seen.get(short, 0)
    seen[short] = n + 1
    out = os.path.join(root, "profiles", f"{tag}_ncu_full_{short}" + (f"_{n}" if n else "") + ".csv")
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", r[idx["Kernel Name"]], ""])
        for m in KEEP:
            if m in idx:
                w.writerow([m, r[idx[m]], units[idx[m]]])
    def num(m):
        v = float(r[idx[m]].replace(",", ""))
        u = units[idx[m]].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "tbyte": 1e12}.get(u, 1)
    t = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    traffic.setdefault(short, []).append({"dram_bytes": t, "ms": float(r[idx["gpu__time_duration.sum"]].replace(",", "")) *
                                          {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1),
                                          "file": os.path.relpath(out, root)})
    print(f"{short:40s} dram {t / 1e9:7.3f} GB  -> {os.path.relpath(out, root)}")
tpath = os.path.join(root, "profiles", "traffic.json")
tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
tj[f"{workload}_kernels"] = {"source": f"ncu --set full --clock-control none, report tag {tag} (per launch, cold cache)", "launches": traffic}
json.dump(tj, open(tpath, "w"), indent=1)
