#!/usr/bin/env python
"""profiles/r2_roofline_inputs.json from an `ncu --set full` capture of the headline kernel:

    python profiles/make_roofline_inputs.py gpurun_out/prof_r2_full.ncu-rep k_grid_bwd 536870912 \
        "ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 (4096 models: the launch the bench times)"

bench.py reads the file for the counters that do not depend on the FLOP contract (executed FP64 instructions and
FLOP per sample, FP64-pipe active %, DRAM bytes per sample of one launch)."""
import collections
import csv
import json
import re
import subprocess
import sys

rep, kname, nsamp, how = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + kname, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


m = {h: num(v) for h, v in zip(hdr, vals)}
u = dict(zip(hdr, units))


def to_bytes(key):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]
    return m[key] * scale


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kname,
                      "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]
ia, isrc = h2.index("Instructions Executed"), h2.index("Source")
tot = collections.Counter()
for r in rows[2:]:
    try:
        n = int(r[ia])
    except (ValueError, IndexError):
        continue
    tot[re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]] += n
scale = m["smsp__inst_executed.sum"] / sum(tot.values())  # the source page may aggregate several launches
per = {k: v * scale * 32 / nsamp for k, v in tot.items()}
fp64 = sum(per.get(k, 0.0) for k in ("DFMA", "DMUL", "DADD", "DSETP"))
dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
out = {
    "source": f"{rep.split('/')[-1]}: {how}",
    "kernel": m["Kernel Name"].split("(")[0],
    "samples_per_launch": int(nsamp),
    "gpu_time_ms": m["gpu__time_duration.sum"] * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]],
    "registers_per_thread": m["launch__registers_per_thread"],
    "warps_active_per_sm": m["sm__warps_active.avg.per_cycle_active"],
    "pipe_fp64_active": m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"] / 100.0,
    "issue_active": m["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100.0,
    "inst_per_sample": sum(per.values()),
    "fp64_inst_per_sample": fp64,
    "executed_flop_per_sample": 2.0 * per.get("DFMA", 0.0) + per.get("DMUL", 0.0) + per.get("DADD", 0.0),
    "dfma_dmul_dadd_per_sample": [per.get("DFMA", 0.0), per.get("DMUL", 0.0), per.get("DADD", 0.0)],
    "dram_bytes_per_launch": dram,
    "dram_bytes_per_sample": dram / nsamp,
    "l2_hit_rate": m["lts__t_sector_hit_rate.pct"] / 100.0,
}
print(json.dumps(out, indent=1))
