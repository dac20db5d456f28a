#!/bin/bash
# Tests + A/B of kernel variants + ncu capture on the GPU box.  Usage: bash scripts/gpu_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
for c in lego_1m dtu_500k owlii_2m lego_100k; do
  echo "== $c default"; python scripts/quick_perf.py --config $c | tee -a $OUT/ab.jsonl | cut -c1-1500
done
echo "== lego_1m SFB_NO_CULL=1"; SFB_NO_CULL=1 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/ab.jsonl | cut -c1-300
echo "== lego_1m SFB_SORT=legacy"; SFB_SORT=legacy python scripts/quick_perf.py --config lego_1m | tee -a $OUT/ab.jsonl | cut -c1-300
python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-3000 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $OUT/launches.csv \
    python scripts/quick_perf.py --config lego_1m --iters 2 --warmup 4 > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'onesweep_pass_kernel|render_forward_kernel|render_backward_kernel|geom_backward_kernel|preprocess_kernel|duplicate_kernel|radix_hist_all_kernel|tile_ranges' \
    -s 30 -c 12 -o $OUT/prof python scripts/quick_perf.py --config lego_1m --iters 1 --warmup 3 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log | cut -c1-300
