#!/bin/bash
# Round evidence in one gpurun call: GPU parity suites, the default bench line (with cpu_baseline),
# the ncu launch list of one short bench run, and one `ncu --set full` capture of every launch of one
# forward pass (exported to CSV on the box; the .ncu-rep of the whole step is too large to bring back).
# Usage: tools/gpu_evidence.sh <tag> [launches per forward]
tag=${1:-evidence}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu -s > $out/pytest.log 2>&1
echo "pytest exit $?" | tee $out/summary.txt
tail -3 $out/pytest.log
grep "\[parity\]" $out/pytest.log > $out/parity_lines.txt
timeout 900 python bench.py --profile-out $out/launch_table.json > $out/bench.json 2> $out/bench.err
echo "bench exit $?" | tee -a $out/summary.txt
cat $out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
echo "reference arm exit $?" | tee -a $out/summary.txt
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
OURS="regex:gemm_kernel|halo_kernel|vit_block_kernel|conv_chain_kernel|stem_chain_kernel|stem_umma_kernel|gelan_tail_kernel|conv1_kernel|conv1_tc_kernel|attention_kernel|attention_tc_kernel|pose_head_kernel|pose_head_tc_kernel|cls_head_kernel|fill_cls_kernel"
NL=${2:-35}  # launches per forward
timeout 600 $SHORT > $out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -s $((3 * NL)) -c $((2 * NL)) --csv \
    --log-file $out/launches.csv $SHORT > $out/ncu_list.log 2>&1
echo "ncu list exit $?" | tee -a $out/summary.txt
timeout 1500 ncu --set full --clock-control none -k "$OURS" -s $((3 * NL)) -c $NL -o $out/step_full $SHORT > $out/ncu_full.log 2>&1
echo "ncu full exit $?" | tee -a $out/summary.txt
ncu -i $out/step_full.ncu-rep --page raw --csv > $out/step_full_raw.csv 2>> $out/ncu_full.log
# the memory-bound tail (get_max_preds, crop normalise, fused crop warp): one full capture of one warm launch each,
# taken from the bench's roofline_memory pass
TAIL="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 $TAIL > $out/plain_tail.log 2>&1 || echo "plain tail run failed" | tee -a $out/summary.txt
: > $out/tail_full_raw.csv
for kn in max_preds_kernel crop_normalize_kernel crop_warp_normalize_kernel; do
  timeout 600 ncu --set full --clock-control none -k "regex:$kn" -s 4 -c 1 -o $out/tail_$kn $TAIL > $out/ncu_tail_$kn.log 2>&1
  echo "ncu tail $kn exit $?" | tee -a $out/summary.txt
  ncu -i $out/tail_$kn.ncu-rep --page raw --csv >> $out/tail_full_raw.csv 2>> $out/ncu_tail_$kn.log
  rm -f $out/tail_$kn.ncu-rep
done
rm -f $out/step_full.ncu-rep
ls -la $out
