#!/bin/bash
TAG=${1:-s12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log | cut -c1-300
SUM='
import json,sys
d=json.loads(sys.stdin.read()); f=d["fwd_stages"]; b=d["bwd_stages"]
print("fwd_ms",round(d["fwd_ms"],4),"bwd_ms",round(d["bwd_ms"],4),{k:round(v*1e3,1) for k,v in list(f.items())+list(b.items())})'
for c in lego_1m lego_100k dtu_500k; do for m in 1 0; do
  echo "== $c SMALL=$m"; SFB_SORT_SMALL=$m timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf_small$m.jsonl | python -c "$SUM"
done; done
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-400 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
