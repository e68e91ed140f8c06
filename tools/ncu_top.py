#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel and the top stall lines (needs ncu on PATH)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"]
idx = [i for i, h in enumerate(hdr) if h in want]
print("==== kernels")
for r in rows[2:]:
    print(" | ".join(f"{hdr[i].split('.')[0][-34:]}={r[i][:60]}{units[i]}" for i in idx))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for s in sections:
    key = s["name"][:70]
    if key in seen:
        continue
    seen.add(key)
    h = s["hdr"]
    ia, isrc = h.index("# Samples"), h.index("Source")
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[ia] or 0) for r in s["rows"]) or 1
    print(f"==== {s['name'][:110]}  samples {tot}, {len(s['rows'])} SASS lines")
    agg = {}
    for r in s["rows"]:
        for i in stall:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print("   stall mix:", ", ".join(f"{k[6:]} {100*v/tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    for r in sorted(s["rows"], key=lambda r: -int(r[ia] or 0))[:ntop]:
        st = sorted([(int(r[i] or 0), h[i][6:]) for i in stall], reverse=True)[:2]
        print("   %5.1f%%  %-70s %s" % (100 * int(r[ia]) / tot, r[isrc][:70], st))
