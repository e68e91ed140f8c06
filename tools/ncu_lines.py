#!/usr/bin/env python
"""Per-source-line summary of one .ncu-rep (needs -lineinfo and --import-source on): share of the warp-stall samples
and of the executed warp instructions per CUDA source line.  Usage: tools/ncu_lines.py <report> [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hdr, out = None, None, []
num = lambda v: int(v) if v.lstrip("-").isdigit() else 0
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        out.append((num(r[hdr.index("# Samples")]), num(r[hdr.index("Instructions Executed")]), cur, int(r[0]), r[1].strip()[:105]))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print("total samples", tot, "total warp instructions", toti)
print("--- by samples")
for o in sorted(out, reverse=True)[:top]:
    print(f"{100 * o[0] / tot:5.1f}% samp {100 * o[1] / toti:5.1f}% inst  {o[2]}:{o[3]}  {o[4]}")
print("--- by instructions")
for o in sorted(out, key=lambda o: -o[1])[:top]:
    print(f"{100 * o[0] / tot:5.1f}% samp {100 * o[1] / toti:5.1f}% inst  {o[2]}:{o[3]}  {o[4]}")
