#!/bin/bash
# Source-level ncu captures of single launches picked by kernel-name regex and launch index.
# Usage: tools/gpu_ncu2.sh <tag> "<regex>:<skip>:<count>" ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 $SHORT > $out/plain.log 2>&1 || exit 1
i=0
for spec in "$@"; do
  IFS=: read regex skip count <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c $count -o $out/prof_$i $SHORT > $out/ncu_$i.log 2>&1
  echo "ncu $i ($regex) exit $?" | tee -a $out/summary.txt
  i=$((i + 1))
done
