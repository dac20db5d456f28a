#!/bin/bash
# Quick correctness + timing loop on the GPU box.  Usage: bash scripts/gpu_quick.sh <tag> [ncu-regex]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
for c in lego_1m dtu_500k owlii_2m; do
  echo "== $c default"; python scripts/quick_perf.py --config $c | tee -a $OUT/ab.jsonl | cut -c1-1500
  echo "== $c SFB_SORT=legacy"; SFB_SORT=legacy python scripts/quick_perf.py --config $c | tee -a $OUT/ab.jsonl | cut -c1-900
done
python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-3000 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
if [ -n "${2:-}" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s 20 -c 8 -o $OUT/prof \
      python scripts/quick_perf.py --config lego_1m --iters 1 --warmup 3 > $OUT/ncu_full.log 2>&1
  tail -2 $OUT/ncu_full.log | cut -c1-200
fi
