#!/bin/bash
# Full GPU suite + bench + ncu launch list (+ optional full capture of the GEMM kernel).
# Usage: tools/gpu_bench.sh <tag> [full]
tag=${1:-bench}
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests -q -m gpu > $out/pytest.log 2>&1
echo "pytest exit $?" | tee $out/summary.txt
tail -3 $out/pytest.log
timeout 900 python bench.py --profile-out $out/launch_table.json > $out/bench.json 2> $out/bench.err
echo "bench exit $?" | tee -a $out/summary.txt
cat $out/bench.json
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 $SHORT > $out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 130 --csv --log-file $out/launches.csv $SHORT > $out/ncu_list.log 2>&1
echo "ncu list exit $?" | tee -a $out/summary.txt
if [ "$2" = "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 44 -c 6 -o $out/gemm_full $SHORT > $out/ncu_full.log 2>&1
  echo "ncu full exit $?" | tee -a $out/summary.txt
fi
