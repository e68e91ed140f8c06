#!/bin/bash
# 2-GPU check of the bucketed all-reduce under the backward: parity tests, then the training bench with the
# overlapped exchange (default) and with the single all-reduce after the backward (HGR_TRAIN_OVERLAP=0).
# Usage (gpurun --gpus 2): tools/gpu_overlap.sh <tag>
tag=${1:-overlap}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_train.py -q -s -m gpu -k "three_parts or two_gpus or cuda_graph or trainer" \
  > $out/tests.log 2>&1
echo "tests exit $?" | tee $out/summary.txt
N=$(nvidia-smi -L | wc -l)
for ov in 1 0 1 0; do
  HGR_TRAIN_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus $N --workload train --train-graph --steps 200 --warmup 20 \
    > $out/train_ov$ov.json 2> $out/train_ov$ov.err
  echo "bench overlap=$ov exit $?" | tee -a $out/summary.txt
  tail -n 1 $out/train_ov$ov.json | cut -c1-600 | tee -a $out/summary.txt
done
grep -hE "passed|failed|error|Error" $out/tests.log | tail
