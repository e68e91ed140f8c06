#!/bin/bash
# Fused ViT layer tail: unit parity, whole-forward parity, then A/B against the three separate launches.
tag=${1:-vit}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -s -m gpu -k "vit_block" > $out/unit.log 2>&1
echo "unit exit $?" | tee $out/summary.txt
grep -h "\[parity\]" $out/unit.log | head -20
tail -5 $out/unit.log
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_ops.py -q -m gpu -x > $out/pytest.log 2>&1
echo "pytest exit $?" | tee -a $out/summary.txt
tail -3 $out/pytest.log
for v in 0 1 0 1; do
  env HGR_VIT_FUSED=$v timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 40 --profile-out $out/table_$v.json > $out/bench_$v.json 2>> $out/bench.err
  python - <<PY
import json
d=json.load(open("$out/bench_$v.json"))
print("HGR_VIT_FUSED=$v: value %.0f ms/step %.3f gemm frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
tail -3 $out/bench.err
python tools/show_table.py $out/table_1.json | grep -i "layers.1\|step_ms"
