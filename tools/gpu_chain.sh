#!/bin/bash
# conv2 -> cspelan1.cv1 chained kernel: unit parity, whole-forward parity, then A/B against the two launches.
tag=${1:-chain}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -s -m gpu -k "conv_chain" > $out/unit.log 2>&1
echo "unit exit $?" | tee $out/summary.txt
grep -h "\[parity\]" $out/unit.log | cut -c1-150 | head -12
grep -E "passed|failed|error|hgr:" $out/unit.log | tail -5
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_ops.py -q -m gpu -x > $out/pytest.log 2>&1
echo "pytest exit $?" | tee -a $out/summary.txt
tail -3 $out/pytest.log
for v in 0 1 0 1; do
  env HGR_CONV_CHAIN=$v timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 40 --profile-out $out/table_$v.json > $out/bench_$v.json 2>> $out/bench.err
  python - <<PY
import json
d=json.load(open("$out/bench_$v.json"))
print("HGR_CONV_CHAIN=$v: value %.0f ms/step %.3f gemm frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
tail -3 $out/bench.err
python tools/show_table.py $out/table_0.json | grep -E "conv2|cspelan1.cv1 |step_ms"
python tools/show_table.py $out/table_1.json | grep -E "conv2|step_ms"
