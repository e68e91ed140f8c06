#!/bin/bash
# Training-step parity on the GPU, test by test so that one trapped kernel does not hide the others.
# Usage: tools/gpu_train.sh <tag> [pytest -k expression]
tag=${1:-train}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests/test_gpu_train.py -q -s -m gpu -k "$2" > $out/train.log 2>&1
  echo "train exit $?" | tee $out/summary.txt
else
  for grp in wgrad dgrad_stride2 attention_bwd loss_kernel adamw_kernel forward_backward drop_in trainer; do
    timeout 900 python -m pytest tests/test_gpu_train.py -q -s -m gpu -k "$grp" > $out/train_$grp.log 2>&1
    echo "train $grp exit $?" | tee -a $out/summary.txt
  done
fi
grep -h "\[parity\]" $out/*.log | cut -c1-260 > $out/parity.txt
grep -hE "passed|failed|error|Error" $out/*.log | tail -30
