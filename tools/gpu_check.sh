#!/bin/bash
# Runs the GPU parity suites group by group (a trapped kernel poisons only its own process)
# and leaves the logs under gpurun_out/.  Usage: tools/gpu_check.sh [tag]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > $out/gpu.txt 2>&1
ls /root/reference > $out/reference_present.txt 2>&1
for grp in conv_bn_act linear conv1 layernorm attention cls_head pose_head; do
  timeout 600 python -m pytest tests/test_gpu_ops.py -q -s -m gpu -k "test_$grp" > $out/ops_$grp.log 2>&1
  echo "ops $grp exit $?" | tee -a $out/summary.txt
done
timeout 600 python -m pytest tests/test_gpu_tail.py -q -s -m gpu > $out/tail.log 2>&1
echo "tail exit $?" | tee -a $out/summary.txt
timeout 900 python -m pytest tests/test_gpu_forward.py -q -s -m gpu > $out/forward.log 2>&1
echo "forward exit $?" | tee -a $out/summary.txt
grep -h "\[parity\]" $out/*.log > $out/parity.txt
grep -hE "passed|failed|error" $out/*.log | tail -20
