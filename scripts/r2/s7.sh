#!/bin/bash
# Round-2 session 7 (1 GPU): the GPU suite (depth cotangent, every A/B arm through the parity file), quick_perf, bench.
TAG=${1:-r2s7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -18 $OUT/pytest_gpu.log | cut -c1-600
timeout 150 python scripts/quick_perf.py --config lego_1m >> $OUT/quick_perf.jsonl 2>>$OUT/quick_perf.err; cut -c1-900 $OUT/quick_perf.jsonl
timeout 400 python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
