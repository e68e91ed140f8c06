#!/usr/bin/env python
"""Prints the timeline of stem_umma_kernel's CTA 0 from a HGR_STEM_TRACE file (development tool).
Usage: HGR_STEM_TRACE=trace.txt python tools/run_op.py stem_fused 1024 1; tools/stem_trace.py trace.txt [first_item]"""
import collections, sys
ev = collections.defaultdict(list)
for l in open(sys.argv[1]):
    r, e, i, t = map(int, l.split())
    ev[r].append((e, i, t))
it0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10
t0 = min(t for r in ev for _, _, t in ev[r])
print("--- MMA warp issues")
prev = None
for e, i, t in sorted(ev[1], key=lambda x: x[2]):
    name = {0: f"G0(B{i})", 1: f"G1(i{i // 4},p{i % 4})", 2: f"G2({i})"}[e]
    item = {0: i // 5, 1: i // 4, 2: i}[e]
    if it0 <= item < it0 + 3:
        print(f"{t - t0:8d} (+{t - prev if prev else 0:5d}) {name}")
        prev = t
for r, name in ((4, "G0 group 0"), (5, "G0 group 1")):
    print("---", name, "(per block: wait c1_full | tcgen05.ld | release | plane waits | SiLU + stores + arrivals)")
    by = collections.defaultdict(dict)
    for e, i, t in ev[r]:
        by[i][e] = t - t0
    for B in sorted(by):
        d = by[B]
        if it0 * 5 <= B < (it0 + 3) * 5 and all(k in d for k in (0, 1, 2, 4, 5)):
            print(f"B{B:4d} start {d[0]:7d}: wait {d[1] - d[0]:5d} ld {d[2] - d[1]:4d} planes {d[4] - d[2]:5d} math {d[5] - d[4]:5d}  total {d[5] - d[0]:5d}")
for r, name in ((2, "E group 0"), (3, "E group 1")):
    print("---", name, "(per item: wait acc_full | E1 | wait acc2_full | E2)")
    by = collections.defaultdict(dict)
    for e, i, t in ev[r]:
        by[i][e] = t - t0
    for it in sorted(by):
        d = by[it]
        if it0 - 2 <= it < it0 + 4 and len(d) == 5:
            print(f"item {it:3d} start {d[0]:7d}: wait {d[1] - d[0]:5d} E1 {d[2] - d[1]:5d} wait {d[3] - d[2]:5d} E2 {d[4] - d[3]:5d} total {d[4] - d[0]:5d}")
if 6 in ev:
    print("--- builder warp 2: block announced at")
    prev = None
    for e, i, t in sorted(ev[6], key=lambda x: x[2]):
        if it0 * 5 <= i < (it0 + 2) * 5:
            print(f"{t - t0:8d} (+{t - prev if prev else 0:5d}) A(B{i})")
            prev = t
