#!/bin/bash
# Session 14: full GPU tests on the trimmed render kernels, torch.norm rounding diagnostic, fast-exp margins + A/B.
TAG=${1:-s14}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-400
python scripts/diag_norm.py 2>&1 | tee $OUT/diag_norm.txt
python scripts/parity_margin.py 2>&1 | tee $OUT/parity_margin.jsonl | cut -c1-700
SFB_LIB_VARIANT=fastexp python scripts/parity_margin.py 2>&1 | tee -a $OUT/parity_margin.jsonl | cut -c1-700
SUM='
import json,sys
d=json.loads(sys.stdin.read()); f=d["fwd_stages"]; b=d["bwd_stages"]
print("fwd_ms",round(d["fwd_ms"],4),"bwd_ms",round(d["bwd_ms"],4),{k:round(v*1e3,1) for k,v in list(f.items())+list(b.items())})'
echo "== lego_1m default"; timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf.jsonl | python -c "$SUM"
echo "== lego_1m fastexp"; SFB_LIB_VARIANT=fastexp timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_fastexp.jsonl | python -c "$SUM"
echo "== dtu_500k default"; timeout 120 python scripts/quick_perf.py --config dtu_500k | tee -a $OUT/quick_perf.jsonl | python -c "$SUM"
echo "== dtu_500k fastexp"; SFB_LIB_VARIANT=fastexp timeout 120 python scripts/quick_perf.py --config dtu_500k | tee -a $OUT/quick_perf_fastexp.jsonl | python -c "$SUM"
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
SFB_LIB_VARIANT=fastexp timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1_fastexp.json 2> $OUT/bench_n1_fastexp.err; cut -c1-300 $OUT/bench_n1_fastexp.json; tail -3 $OUT/bench_n1_fastexp.err
