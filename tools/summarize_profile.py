#!/usr/bin/env python
"""Turn gpurun_out/<tag>_prof.ncu-rep (ncu --set full of one fwd + one bwd launch) and
gpurun_out/<tag>_launches.csv (ncu --metrics gpu__time_duration.sum launch list of the same bench command) into
the tracked summaries under profiles/: <tag>_ncu_raw_selected.csv, <tag>_launches.csv (kernel, duration),
traffic.json (DRAM bytes per launch, read by bench.py) and a share-of-step table printed to stdout.
usage: python tools/summarize_profile.py r1b"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1b"
rep = os.path.join(ROOT, "gpurun_out", tag + "_prof.ncu-rep")
KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__registers_per_thread
launch__grid_size launch__block_size launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__occupancy_limit_warps smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
smsp__warps_active.avg.per_cycle_active smsp__average_warp_latency_per_inst_issued.ratio
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed dram__throughput.avg.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__cycles_elapsed.avg smsp__cycles_active.avg
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
lts__t_sector_hit_rate.pct sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active""".split()

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
names = [r[hdr.index("Kernel Name")] for r in data]
out = os.path.join(ROOT, "profiles", tag + "_ncu_raw_selected.csv")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + names)
    for i, h in enumerate(hdr):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            w.writerow([h, units[i]] + [r[i] for r in data])
print("wrote", out)

def col(name):
    i = hdr.index(name)
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
    return [float(r[i]) * scale for r in data]

rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
traffic = {"source": "ncu --set full --clock-control none, 1 launch each, N=1048576 poses (profiles/%s_ncu_raw_selected.csv: "
                     "dram__bytes_read.sum + dram__bytes_write.sum)" % tag,
           "algorithmic": {"dhfk_fwd_kernel": 562036736, "dhfk_bwd_kernel": 725614592}}
for n, r, w_ in zip(names, rd, wr):
    key = "dhfk_fwd_kernel" if "fwd" in n else "dhfk_bwd_kernel"
    traffic[key + "_bytes_per_launch"] = r + w_
    traffic[key + "_read_write"] = [r, w_]
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("wrote profiles/traffic.json", {k: v for k, v in traffic.items() if k.endswith("per_launch")})

# launch list -> kernel, duration; share of the fwd+bwd step
ll = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
lines = [l for l in open(ll) if l.startswith('"')]
rd_ = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.OrderedDict()
with open(os.path.join(ROOT, "profiles", tag + "_launches.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum (%s)" % (rd_[0]["Metric Unit"] if rd_ else "")])
    for r in rd_:
        if "dhfk" not in r["Kernel Name"]:
            continue
        w.writerow([r["ID"], r["Kernel Name"], r["Grid Size"], r["Block Size"], r["Metric Value"]])
        agg.setdefault(r["Kernel Name"], []).append(float(r["Metric Value"].replace(",", "")))
print("dhfk launches in the list:", sum(len(v) for v in agg.values()))
for k, v in agg.items():
    print("  %-70s n=%3d  avg %.1f  min %.1f" % (k[:70], len(v), sum(v) / len(v), min(v)))
