#!/bin/bash
# Session 15: fused tile ranges / folded memsets / host-visible num_rendered (A/B via knobs), fast exp as default.
TAG=${1:-s15}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-400
SFB_LIB_VARIANT=exactexp timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q > $OUT/pytest_gpu_exactexp.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_exactexp.log; tail -3 $OUT/pytest_gpu_exactexp.log | cut -c1-300
SUM='
import json,sys
d=json.loads(sys.stdin.read()); f=d["fwd_stages"]; b=d["bwd_stages"]
print("fwd_ms",round(d["fwd_ms"],4),"bwd_ms",round(d["bwd_ms"],4),"ksum_fwd",round(sum(f.values()),4),{k:round(v*1e3,1) for k,v in list(f.items())+list(b.items())})'
for c in lego_1m dtu_500k; do
echo "== $c default"; timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | python -c "$SUM"
echo "== $c SFB_FUSED_RANGES=0"; SFB_FUSED_RANGES=0 timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf_ranges_kernel.jsonl | python -c "$SUM"
echo "== $c SFB_FOLD_MEMSETS=0"; SFB_FOLD_MEMSETS=0 timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf_memset_nodes.jsonl | python -c "$SUM"
echo "== $c SFB_NR_MEMCPY=1"; SFB_NR_MEMCPY=1 timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf_nr_memcpy.jsonl | python -c "$SUM"
done
for c in lego_100k owlii_2m; do echo "== $c default"; timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf.jsonl | python -c "$SUM"; done
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-300 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
SFB_FUSED_RANGES=0 SFB_FOLD_MEMSETS=0 SFB_NR_MEMCPY=1 timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1_old_forward.json 2> $OUT/bench_n1_old_forward.err; cut -c1-300 $OUT/bench_n1_old_forward.json
