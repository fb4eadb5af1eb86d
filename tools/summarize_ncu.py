#!/usr/bin/env python
"""Summarise gpurun_out/*.ncu-rep + launches csv into small text/JSON files under profiles/ (tracked)."""
import csv
import collections
import json
import subprocess
import sys

rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]
keep += [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued")]
out = []
for r in rows[2:]:
    out.append({k: (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip() for k in keep if k in hdr})
with open("profiles/%s_ncu_full_summary.json" % tag, "w") as fh:
    json.dump(out, fh, indent=1)
# launch list: per-kernel totals and shares
tot = collections.defaultdict(lambda: [0, 0.0])
with open(launches) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    name = r["Kernel Name"]
    tot[name][0] += 1
    tot[name][1] += ms
total = sum(v[1] for v in tot.values())
with open("profiles/%s_launches_summary.txt" % tag, "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    fh.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    fh.write("%-90s %6s %10s %7s\n" % ("kernel", "count", "total ms", "share"))
    for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        fh.write("%-90s %6d %10.3f %6.1f%%\n" % (name[:90], n, ms, 100 * ms / total))
print(open("profiles/%s_launches_summary.txt" % tag).read())
