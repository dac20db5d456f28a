#!/bin/bash
# Round-2 session 4 (1 GPU): the fused backward + exchange kernel on emulated ranks, the GPU suite, quick_perf.
TAG=${1:-r2s4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tests/exchange_emulation.py > $OUT/emulation.json 2> $OUT/emulation.err; echo "emulation rc=$?"; cut -c1-2500 $OUT/emulation.json; tail -5 $OUT/emulation.err | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-400
for c in lego_1m dtu_500k owlii_2m; do timeout 150 python scripts/quick_perf.py --config $c >> $OUT/quick_perf.jsonl 2>>$OUT/quick_perf.err; done; cut -c1-900 $OUT/quick_perf.jsonl
timeout 400 python bench.py --steps 200 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-400 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
