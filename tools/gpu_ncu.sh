#!/bin/bash
# One ncu --set full capture of selected kernels.  Usage: tools/gpu_ncu.sh <tag> <kernel regex> <skip> <count>
tag=$1; regex=$2; skip=$3; count=$4
out=gpurun_out/$tag
mkdir -p $out
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 $SHORT > $out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c $count -o $out/prof $SHORT > $out/ncu.log 2>&1
echo "ncu exit $?" | tee $out/summary.txt
tail -3 $out/ncu.log
