#!/bin/bash
# First GPU session of round 2 (1 GPU): everything written after round 1's GPU budget was spent.
#   1. the GPU suite, then the experimental tests (SFB_EARLY_GATHER hand-off)
#   2. bench at N = 1 (must still read ~1234 Msplats/s)
#   3. ncu --set full of the kernels that have no capture yet (activate, knn, sh_grad_combine, loss, densify)
#   4. A/B of the opt-in tight tile rectangles (SFB_TIGHT_RECT=1)
#   5. config 4 in one process (view x time jobs, precomputed RGB)
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log | cut -c1-300
SFB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_sh_factored.py -m gpu -q > $OUT/pytest_experimental.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_experimental.log; tail -5 $OUT/pytest_experimental.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-300 $OUT/bench_n1.json
K='activate_forward_kernel|activate_backward_kernel|knn_query_kernel|knn_boxes_kernel|knn_morton_kernel|sh_grad_combine_kernel|ssim_stats_kernel|ssim_grad_kernel|densify_stats_kernel|densify_masks_kernel|extract_dcolor_kernel'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 24 -f -o $OUT/prof_next_rows \
    python scripts/quick_perf_next_rows.py --iters 1 --warmup 0 > $OUT/ncu_next_rows.log 2>&1
ncu -i $OUT/prof_next_rows.ncu-rep --page raw --csv > $OUT/ncu_next_rows_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/ncu_next_rows_raw.csv > $OUT/ncu_next_rows_summary.txt 2>&1; grep -c "^==" $OUT/ncu_next_rows_summary.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_new_rows.py > $OUT/memcheck_new_rows.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck_new_rows.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" $OUT/memcheck_new_rows.log | head -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_new_rows.py > $OUT/racecheck_new_rows.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/racecheck_new_rows.log; grep -E "RACECHECK SUMMARY|hazard" $OUT/racecheck_new_rows.log | head -5
for v in 0 1; do echo "== SFB_PRE_OCC4=$v"; SFB_PRE_OCC4=$v timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_pre_occ4_$v.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['fwd_ms'], d['fwd_stages']['preprocess'])"; done
for c in lego_100k lego_1m dtu_500k; do timeout 300 python scripts/ab_tight_rect.py --config $c >> $OUT/ab_tight_rect.jsonl 2>> $OUT/ab_tight_rect.err; done; cat $OUT/ab_tight_rect.jsonl
timeout 300 python scripts/run_view_time.py --rounds 24 > $OUT/view_time_n1.json 2> $OUT/view_time_n1.err; cat $OUT/view_time_n1.json; tail -2 $OUT/view_time_n1.err
