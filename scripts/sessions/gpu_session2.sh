#!/bin/bash
# Full session: tests, quick_perf on all configs (+ sort tile A/B), bench, ncu launch list + full captures.
TAG=${1:-s2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log | cut -c1-300
for c in lego_1m lego_100k dtu_500k owlii_2m; do
  echo "== $c"; python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | cut -c1-1500
done
echo "== lego_1m SFB_SORT_IPT=8"; SFB_SORT_IPT=8 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_ipt8.jsonl | cut -c1-900
echo "== owlii_2m SFB_SORT_IPT=8"; SFB_SORT_IPT=8 python scripts/quick_perf.py --config owlii_2m | tee -a $OUT/quick_perf_ipt8.jsonl | cut -c1-900
python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-2200 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-500 $OUT/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 50 --csv --log-file $OUT/launches.csv \
    python scripts/quick_perf.py --config lego_1m --iters 2 --warmup 4 > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'onesweep_pass_kernel|render_forward_kernel|render_backward_kernel|geom_backward_kernel|preprocess_kernel|duplicate_kernel|radix_hist_all_kernel|tile_ranges' \
    -s 30 -c 12 -o $OUT/prof python scripts/quick_perf.py --config lego_1m --iters 1 --warmup 3 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
