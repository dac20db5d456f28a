#!/bin/bash
TAG=${1:-s5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log | cut -c1-400
echo "== microbench aos_copy"; ./scripts/microbench/aos_copy | tee $OUT/aos_copy.jsonl
echo "== lego_1m default"; python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf.jsonl | cut -c1-1400
echo "== lego_1m SFB_NO_LD256=1"; SFB_NO_LD256=1 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_nold256.jsonl | cut -c1-1400
for c in dtu_500k owlii_2m; do echo "== $c"; python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | cut -c1-1400; done
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-700 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
