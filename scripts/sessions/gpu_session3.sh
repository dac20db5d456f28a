#!/bin/bash
# tests + timing + sanitizers
TAG=${1:-s3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log | cut -c1-400
for c in lego_1m dtu_500k owlii_2m; do
  echo "== $c"; python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | cut -c1-1500
done
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-900 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
echo "== memcheck (lego_100k, 2 steps)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/quick_perf.py --config lego_100k --iters 1 --warmup 1 > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" $OUT/memcheck.log | head -5
echo "== racecheck (plumbing_256)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/quick_perf.py --config plumbing_256 --iters 1 --warmup 1 > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/racecheck.log; grep -E "RACECHECK SUMMARY|hazard" $OUT/racecheck.log | head -8
