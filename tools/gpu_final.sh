#!/bin/bash
# Final pass on one GPU without profilers: the whole GPU suite, the default bench line, the reference arm, the training line.
# Usage: tools/gpu_final.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu -s > $out/pytest.log 2>&1
echo "pytest exit $?" | tee $out/summary.txt
tail -3 $out/pytest.log | tee -a $out/summary.txt
grep "\[parity\]" $out/pytest.log > $out/parity_lines.txt
timeout 900 python bench.py --profile-out $out/launch_table.json > $out/bench.json 2> $out/bench.err
echo "bench exit $?" | tee -a $out/summary.txt
cut -c1-700 $out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
echo "reference arm exit $?" | tee -a $out/summary.txt
timeout 600 python bench.py --workload train --train-graph --steps 200 --warmup 20 > $out/bench_train.json 2>> $out/bench.err
echo "train exit $?" | tee -a $out/summary.txt
cut -c1-300 $out/bench_train.json
