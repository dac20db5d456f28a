#!/bin/bash
TAG=${1:-s9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-300
python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== lego_1m"; python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf.jsonl | cut -c1-1400
echo "== lego_1m SFB_BWD_SHFL=1"; SFB_BWD_SHFL=1 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_shfl.jsonl | cut -c1-1400
for c in lego_100k dtu_500k owlii_2m; do echo "== $c"; python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | cut -c1-420; done
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-400 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
