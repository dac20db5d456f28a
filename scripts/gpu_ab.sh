#!/bin/bash
# A/B of kernel variants on the GPU box.  Usage: bash scripts/gpu_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
for c in lego_1m dtu_500k; do
  echo "== $c default"; python scripts/quick_perf.py --config $c | tee -a $OUT/ab.jsonl | cut -c1-1200
  echo "== $c SFB_NO_CULL=1"; SFB_NO_CULL=1 python scripts/quick_perf.py --config $c | tee -a $OUT/ab.jsonl | cut -c1-300
  echo "== $c SFB_SORT=legacy"; SFB_SORT=legacy python scripts/quick_perf.py --config $c | tee -a $OUT/ab.jsonl | cut -c1-300
done
python bench.py --steps 100 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-2500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
