#!/bin/bash
# Session 13: next rows (fused loss, densification) on the GPU + in-forward accumulator clearing + fast-exp A/B.
TAG=${1:-s13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 420 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-400
SFB_LIB_VARIANT=fastexp timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q > $OUT/pytest_gpu_fastexp.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_fastexp.log; tail -5 $OUT/pytest_gpu_fastexp.log | cut -c1-400
SUM='
import json,sys
d=json.loads(sys.stdin.read()); f=d["fwd_stages"]; b=d["bwd_stages"]
print("fwd_ms",round(d["fwd_ms"],4),"bwd_ms",round(d["bwd_ms"],4),{k:round(v*1e3,1) for k,v in list(f.items())+list(b.items())})'
echo "== lego_1m default"; timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf.jsonl | python -c "$SUM"
echo "== lego_1m memset in backward"; SFB_ZERO_IN_FWD=0 timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_zero_in_bwd.jsonl | python -c "$SUM"
echo "== lego_1m fastexp"; SFB_LIB_VARIANT=fastexp timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_fastexp.jsonl | python -c "$SUM"
for c in lego_100k dtu_500k owlii_2m; do echo "== $c"; timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | python -c "$SUM"; done
echo "== next rows"; timeout 200 python scripts/quick_perf_next_rows.py | tee $OUT/quick_perf_next_rows.jsonl | cut -c1-600
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-400 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
