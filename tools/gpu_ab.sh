#!/bin/bash
# A/B of the launch-level switches.  Usage: tools/gpu_ab.sh <tag>
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests -q -m gpu -x > $out/pytest.log 2>&1
echo "pytest exit $?" | tee $out/summary.txt
tail -3 $out/pytest.log
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  HGR_PDL=$1 HGR_ZIGZAG=$2 timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 40 > $out/bench_pdl$1_zz$2.json 2>> $out/bench.err
  python - <<PY
import json
d=json.load(open("$out/bench_pdl$1_zz$2.json"))
print("pdl $1 zigzag $2: value %.0f ms/step %.3f gemm frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]))
PY
done
