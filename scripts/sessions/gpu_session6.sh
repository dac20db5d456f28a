#!/bin/bash
TAG=${1:-s6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log | cut -c1-300
echo "== lego_1m"; python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf.jsonl | cut -c1-1400
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
ncu --set full --clock-control none --import-source on \
    -k regex:'onesweep_pass_kernel|duplicate_kernel|radix_hist_all_kernel|tile_ranges|preprocess_kernel|geom_backward' \
    -s 20 -c 10 -o $OUT/prof python scripts/quick_perf.py --config lego_1m --iters 1 --warmup 3 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
