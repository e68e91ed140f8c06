#!/bin/bash
# A/B of one environment switch on the forward bench.  Usage: tools/gpu_ab2.sh <tag> <ENVVAR>
tag=${1:-ab}; var=${2:-HGR_CLUSTER}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -q -m gpu -x > $out/pytest.log 2>&1
echo "pytest exit $?" | tee $out/summary.txt
tail -3 $out/pytest.log
for v in ${3:-0 1 0 1}; do
  env $var=$v timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 40 --profile-out $out/table_$v.json > $out/bench_$v.json 2>> $out/bench.err
  python - <<PY
import json
d=json.load(open("$out/bench_$v.json"))
print("$var=$v: value %.0f ms/step %.3f gemm frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
tail -3 $out/bench.err
