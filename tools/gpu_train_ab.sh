#!/bin/bash
# Training step on one GPU: parity tests, then the graph-replayed step under each value of one env switch.
# Usage: tools/gpu_train_ab.sh <tag> [switch, default HGR_TRAIN_FORK] [values, default "0 1 2 4 7 0"]
tag=${1:-trainab}; sw=${2:-HGR_TRAIN_FORK}; vals=${3:-0 1 2 4 7 0}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_train.py -q -s -m gpu -x > $out/tests.log 2>&1
echo "tests exit $?" | tee $out/summary.txt
i=0
for v in $vals; do
  i=$((i+1))
  env $sw=$v timeout 600 python bench.py --workload train --train-graph --steps 200 --warmup 20 > $out/train_${i}_$v.json 2> $out/train_${i}_$v.err
  echo "$sw=$v exit $? $(python -c "import json;d=json.loads(open('$out/train_${i}_$v.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['clocks'])")" | tee -a $out/summary.txt
done
grep -h "\[parity\]" $out/tests.log | cut -c1-240 > $out/parity.txt
grep -hE "passed|failed|error|Error" $out/tests.log | tail
