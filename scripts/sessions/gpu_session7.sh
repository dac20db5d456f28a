#!/bin/bash
TAG=${1:-s7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log | cut -c1-300
echo "== lego_1m"; python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf.jsonl | cut -c1-1400
echo "== lego_1m SFB_NO_HITS=1"; SFB_NO_HITS=1 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_nohits.jsonl | cut -c1-1400
for c in lego_100k dtu_500k owlii_2m; do echo "== $c"; python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | cut -c1-700; done
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
SFB_NO_HITS=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1_nohits.json 2> $OUT/bench_n1_nohits.err; cut -c1-300 $OUT/bench_n1_nohits.json
