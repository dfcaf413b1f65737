#!/usr/bin/env python
"""Reduce `ncu --page raw --csv` exports (gpurun_out/r2_ncu_raw_<name>.csv) to the metrics the roofline discussion uses:
profiles/r2_ncu_full_<name>.json, and profiles/traffic.json (DRAM bytes per launch of each configuration's dominant kernel,
read by bench.py for `roofline.traffic`)."""
import csv, json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio")
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
tpath = os.path.join(root, "profiles", "traffic.json")
if os.path.exists(tpath):
    try:
        traffic = json.load(open(tpath))
        if not all(isinstance(v, dict) for v in traffic.values()):
            traffic = {}
    except Exception:
        traffic = {}
for name in sys.argv[1:]:
    src = os.path.join(root, "gpurun_out", f"r2_ncu_raw_{name}.csv")
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2]
    d = {"kernel": data[hdr.index("Kernel Name")], "_note": "one `ncu --set full --clock-control none` capture (cold caches, serialised launch): durations are NOT bench values"}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            try:
                d[k] = {"value": float(data[i].replace(",", "")), "unit": units[i]}
            except ValueError:
                d[k] = {"value": data[i], "unit": units[i]}
    json.dump(d, open(os.path.join(root, "profiles", f"r2_ncu_full_{name}.json"), "w"), indent=1)
    by = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k in d:
            by += d[k]["value"] * UNIT.get(d[k]["unit"], 1.0)
    traffic[name.lower()] = {"dram_bytes_per_launch": by, "kernel": d["kernel"], "source": f"profiles/r2_ncu_full_{name}.json (ncu --set full, one launch)"}
    print(name, d["kernel"][:70], "dram bytes", by, "ms", d.get("gpu__time_duration.sum"))
json.dump(traffic, open(tpath, "w"), indent=1)
