#!/usr/bin/env python
"""gpurun_out/r1_ncu_raw_<cfg>.csv (ncu --page raw) -> profiles/r1_ncu_full_<cfg>.json (the metrics the roofline argument
uses), plus profiles/traffic.json (DRAM bytes per launch of the bench kernel) and the launch-list summary."""
import collections, csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
def raw(cfg):
    rows = list(csv.reader(open(os.path.join(SRC, f"r1_ncu_raw_{cfg}.csv"))))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h, u, v = rows[hi], rows[hi + 1], rows[hi + 2]
    return {n: {"value": v[i], "unit": u[i]} for i, n in enumerate(h) if n in KEEP}
def to_bytes(m):
    x, unit = float(m["value"].replace(",", "")), m["unit"].lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
for cfg in ("C2", "C1x", "C3", "C4", "C5"):
    try:
        d = raw(cfg)
    except Exception as e:
        print(cfg, "missing:", e); continue
    d["_note"] = "one `ncu --set full --clock-control none` capture (cold caches, serialised launch): durations are NOT bench values"
    json.dump(d, open(os.path.join(DST, f"r1_ncu_full_{cfg}.json"), "w"), indent=1)
    shutil.copy(os.path.join(SRC, f"r1_ncu_hot_{cfg}.txt"), os.path.join(DST, f"r1_ncu_hot_{cfg}.txt"))
    if cfg == "C2":
        t = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
        json.dump({"k_stream_dram_bytes_per_launch": t, "source": "profiles/r1_ncu_full_C2.json (dram__bytes_read.sum + dram__bytes_write.sum, one launch = the whole 100M-doc step)",
                   "algorithmic_bytes_per_launch": 875000256}, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
# launch list summary
rows = list(csv.reader(open(os.path.join(SRC, "r1_launches_bench.csv"))))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < 10: continue
    name, ns = r[4].split("(")[0][:70], float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
with open(os.path.join(DST, "r1_launches_bench_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n")
    f.write("(whole process: index synthesis, correctness gate, warm-up, 2 resident steps, 2 host-docset steps)\n")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t/1e3:10.1f} us {100*t/tot:5.1f}%  x{c:<4d} {n}\n")
shutil.copy(os.path.join(SRC, "r1_launches_bench.csv"), os.path.join(DST, "r1_launches_bench.csv"))
for f in ("r1_bench_n1.json", "r1_bench_reference_arm.json", "r1_configs.txt"):
    shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
if os.path.exists(os.path.join(SRC, "scaling.jsonl")):
    shutil.copy(os.path.join(SRC, "scaling.jsonl"), os.path.join(DST, "r1_scaling.jsonl"))
print(open(os.path.join(DST, "r1_launches_bench_summary.txt")).read())
