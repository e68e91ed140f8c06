#!/usr/bin/env python
"""Reduce `ncu --page raw --csv` of one forward (tools/gpu_evidence.sh: step_full_raw.csv) to the headline metrics per
launch, named from the bench's launch table.  Usage: ncu_step_summary.py <step_full_raw.csv> <launch_table.json> <out.csv>"""
import csv, json, sys

raw, table, out = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max"]
idx = [hdr.index(w) for w in want if w in hdr]
launches = json.load(open(table))["launches"]
assert len(launches) == len(data), f"{len(launches)} launches in the table, {len(data)} in the capture"
it = hdr.index("gpu__time_duration.sum")
def us(r):
    v = float(r[it].replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(units[it], v)
tot = sum(us(r) for r in data)
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "share_of_step_ncu", "share_of_step_events"] + [hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
    for l, r in zip(launches, data):
        w.writerow([l["launch"], f"{us(r) / tot:.4f}", f"{l['share']:.4f}"] + [r[i][:60] for i in idx])
print(f"{len(data)} launches, {tot:.1f} us under ncu -> {out}")
