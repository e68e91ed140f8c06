#!/bin/bash
# Parity suites + bench with the per-launch table, no profiler.  Usage: tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-quick}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -k "$2" > $out/pytest.log 2>&1
else
  timeout 1200 python -m pytest tests -q -m gpu > $out/pytest.log 2>&1
fi
echo "pytest exit $?" | tee $out/summary.txt
tail -15 $out/pytest.log
timeout 900 python bench.py --profile-out $out/launch_table.json --no-cpu-baseline > $out/bench.json 2> $out/bench.err
echo "bench exit $?" | tee -a $out/summary.txt
tail -5 $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "roofline", d["roofline"]["frac"], d["clocks"])
PY
