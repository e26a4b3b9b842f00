#!/usr/bin/env python3
"""Turn the raw ncu exports under gpurun_out/ into the small tracked summaries under profiles/.

usage: make_profile_summaries.py <tag> <launches.csv> <full.ncu-rep> [<kernel name substring>]
Writes profiles/<tag>_launches.csv (per-kernel totals + shares), profiles/<tag>_<kernel>_metrics.txt
(selected raw metrics of the full capture) and profiles/<tag>_sweep_dram_bytes.json (DRAM traffic per
read, used by bench.py's roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1:4]
kern = sys.argv[4] if len(sys.argv) > 4 else "k_sweep"
n_reads = int(sys.argv[5]) if len(sys.argv) > 5 else 500000
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

rows = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) > vi:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python tools/profile_step.py (2 M reads, 3 launches of the step)\n")
    f.write("kernel,launches,total_ms,share_pct,avg_ms\n")
    for k, v in agg.items():
        f.write(f"{k},{len(v)},{sum(v)/1e6:.3f},{100*sum(v)/tot:.2f},{sum(v)/len(v)/1e6:.3f}\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
vals = dict(zip(h, zip(v, u)))
with open(os.path.join(out, f"{tag}_{kern}_metrics.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none -k regex:{kern} -c 1 python tools/profile_step.py --reads {n_reads} --steps 1\n")
    for k in want:
        if k in vals:
            f.write(f"{k:75s} {vals[k][0]:>20s} {vals[k][1]}\n")
    stalls = sorted(((float(vals[k][0]), k) for k in vals if k.startswith("smsp__average_warp") and "issue_stalled" in k and vals[k][0]), reverse=True)
    for x, k in stalls[:8]:
        f.write(f"{k:75s} {x:20.3f}\n")
if kern == "k_sweep":
    def num(k):
        x, unit = vals[k]
        x = float(x.replace(",", ""))
        return x * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
    d = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    json.dump({"kernel": "k_sweep", "reads_in_capture": n_reads, "dram_bytes": d, "dram_bytes_per_read": d / n_reads,
               "source": os.path.basename(rep)}, open(os.path.join(out, f"{tag}_sweep_dram_bytes.json"), "w"), indent=1)
print("ok")
