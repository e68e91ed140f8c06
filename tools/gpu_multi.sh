#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): the 2-GPU gradient-parity test, then the forward and training lines at N ranks.
# Usage: tools/gpu_multi.sh <tag>
tag=${1:-multi}
out=gpurun_out/$tag
mkdir -p $out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_train.py -q -s -m gpu -k "two_gpus" > $out/tests.log 2>&1
echo "two-GPU test exit $?" | tee $out/summary.txt
grep -h "\[parity\]" $out/tests.log | cut -c1-240 | tee -a $out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $TR bench.py --gpus $N --workload train --train-graph --steps 200 --warmup 20 > $out/train_${N}gpu.json 2> $out/train.err
echo "train exit $? $(python -c "import json;d=json.loads(open('$out/train_${N}gpu.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'])")" | tee -a $out/summary.txt
timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > $out/forward_${N}gpu.json 2> $out/forward.err
echo "forward exit $? $(python -c "import json;d=json.loads(open('$out/forward_${N}gpu.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('train',{}).get('ms_per_step'))")" | tee -a $out/summary.txt
