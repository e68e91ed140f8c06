#!/bin/bash
# Launch list (device time per kernel) of the training step.  Usage: tools/gpu_train_prof.sh <tag> [batch]
tag=${1:-trainprof}; batch=${2:-32}
out=gpurun_out/$tag
mkdir -p $out
CMD="python bench.py --workload train --batch $batch --steps 1 --warmup 3"
timeout 600 $CMD > $out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file $out/train_launches.csv $CMD > $out/ncu.log 2>&1
echo "ncu exit $?" | tee $out/summary.txt
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$out/train_launches.csv")) if len(r)>10 and r[0].isdigit()]
names=[r[4] for r in rows]; t=[float(r[-1]) for r in rows]
# last step = everything after the last but one adamw_kernel
idx=[i for i,n in enumerate(names) if "adamw_kernel" in n]
a,b=idx[-2]+1, idx[-1]+1  # the last complete step
agg=collections.OrderedDict()
for n,v in zip(names[a:b],t[a:b]):
    k=n.split("(")[0].replace("void hgr::<unnamed>::","").replace("void ","")[:60]
    c=agg.setdefault(k,[0,0.0]); c[0]+=1; c[1]+=v
tot=sum(v[1] for v in agg.values())
print("launches in one step:", b-a, "sum of kernel time: %.1f us" % (tot/1e3))
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-62s %4d launches %9.1f us %5.1f%%" % (k,c,v/1e3,100*v/tot))
PY
