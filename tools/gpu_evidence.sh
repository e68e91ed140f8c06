#!/bin/bash
# Round evidence in one gpurun call: GPU parity suites, the default bench line (with cpu_baseline),
# the ncu launch list of one short bench run, and one `ncu --set full` capture of every launch of one
# forward pass (exported to CSV on the box; the .ncu-rep of the whole step is too large to bring back).
# Usage: tools/gpu_evidence.sh <tag> [launches per forward]
tag=${1:-evidence}
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests -q -m gpu > $out/pytest.log 2>&1
echo "pytest exit $?" | tee $out/summary.txt
tail -3 $out/pytest.log
timeout 900 python bench.py --profile-out $out/launch_table.json > $out/bench.json 2> $out/bench.err
echo "bench exit $?" | tee -a $out/summary.txt
cat $out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
echo "reference arm exit $?" | tee -a $out/summary.txt
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
OURS="regex:gemm_kernel|halo_kernel|vit_block_kernel|conv_chain_kernel|conv1_kernel|attention_kernel|pose_head_kernel|cls_head_kernel|fill_cls_kernel"
NL=${2:-37}  # launches per forward (46 with HGR_VIT_FUSED=0 HGR_CONV_CHAIN=0)
timeout 600 $SHORT > $out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -s $((3 * NL)) -c $((2 * NL)) --csv \
    --log-file $out/launches.csv $SHORT > $out/ncu_list.log 2>&1
echo "ncu list exit $?" | tee -a $out/summary.txt
timeout 1500 ncu --set full --clock-control none -k "$OURS" -s $((3 * NL)) -c $NL -o $out/step_full $SHORT > $out/ncu_full.log 2>&1
echo "ncu full exit $?" | tee -a $out/summary.txt
ncu -i $out/step_full.ncu-rep --page raw --csv > $out/step_full_raw.csv 2>> $out/ncu_full.log
ls -la $out
rm -f $out/step_full.ncu-rep
