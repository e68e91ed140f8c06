#!/usr/bin/env python
"""Pretty-print a bench.py --profile-out launch table, optionally next to an older one."""
import json, sys
t = json.load(open(sys.argv[1]))
old = {l["launch"]: l for l in json.load(open(sys.argv[2]))["launches"]} if len(sys.argv) > 2 else {}
print("step_ms_sum %.3f" % t["step_ms_sum"])
groups = {}
for l in t["launches"]:
    o = old.get(l["launch"])
    print("%-52s %-12s %7.3f ms %5.1f%% %7.1f TF/s %7.1f GB/s %s" % (
        l["launch"][-52:], l["kind"], l["ms"], 100 * l["share"], l["tflops"] or 0, l["gbs"] or 0,
        ("(was %.3f)" % o["ms"]) if o else ""))
    groups[l["kind"]] = groups.get(l["kind"], 0) + l["ms"]
print({k: round(v, 3) for k, v in groups.items()})
