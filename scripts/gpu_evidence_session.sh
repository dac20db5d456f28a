#!/bin/bash
# Evidence session (1 GPU, run under gpurun; publish with scripts/publish_profiles.py): GPU suite, smoke, all configs, bench, ncu launch
# list of the bench command, one --set full capture per kernel of the path and of the rows around it.
TAG=${1:-r2s6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-400
timeout 200 python __graft_entry__.py --smoke-only 2>&1 | tail -1 | tee $OUT/smoke.log
for c in lego_1m lego_100k dtu_500k owlii_2m; do timeout 150 python scripts/quick_perf.py --config $c >> $OUT/quick_perf.jsonl 2>>$OUT/quick_perf.err; done; cut -c1-900 $OUT/quick_perf.jsonl
timeout 400 python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "launch list rows: $(wc -l < $OUT/launches.csv)"
KERNELS='preprocess_kernel|radix_hist_all_kernel|onesweep_pass_kernel|instance_block_sums_kernel|scan_exclusive_kernel|duplicate_kernel|render_forward_kernel|render_backward_mma_kernel|geom_backward_kernel'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -s 45 -c 15 -f -o $OUT/prof \
    python scripts/quick_perf.py --config lego_1m --iters 1 --warmup 3 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/ncu_raw.csv > $OUT/ncu_full_summary.txt 2>&1; grep -c "^==" $OUT/ncu_full_summary.txt
K='activate_forward_kernel|activate_backward_kernel|knn_query_kernel|knn_boxes_kernel|knn_morton_kernel|sh_grad_combine_kernel|ssim_stats_kernel|ssim_grad_kernel|densify_stats_kernel|densify_masks_kernel'
timeout 400 ncu --set full --clock-control none -k regex:"$K" -c 20 -f -o $OUT/prof_next_rows \
    python scripts/quick_perf_next_rows.py --iters 1 --warmup 0 > $OUT/ncu_next_rows.log 2>&1
ncu -i $OUT/prof_next_rows.ncu-rep --page raw --csv > $OUT/ncu_next_rows_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/ncu_next_rows_raw.csv > $OUT/ncu_next_rows_summary.txt 2>&1; grep -c "^==" $OUT/ncu_next_rows_summary.txt
timeout 200 python scripts/quick_perf_next_rows.py > $OUT/quick_perf_next_rows.jsonl 2>> $OUT/quick_perf.err; cut -c1-600 $OUT/quick_perf_next_rows.jsonl
rm -f $OUT/prof_next_rows.ncu-rep
ls -la $OUT | cut -c20-120
