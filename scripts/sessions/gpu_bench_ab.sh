#!/bin/bash
# bench.py under a few environment variants (host-gap investigation).  Usage: bash scripts/gpu_bench_ab.sh <tag>
TAG=${1:-bab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
pick='import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step","host_enqueue_ms_per_step")}, d["e2e"]["ms_per_step"], round(sum(d["stages"]["ms_per_step_by_kernel"].values()),4), d["stages"]["ms_per_step_by_kernel"])'
echo "== default";      python bench.py --steps 200 --warmup 10 --no-cpu-baseline | tee $OUT/bench_default.json | python -c "$pick"
echo "== SFB_TMA=1"; SFB_TMA=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline | tee $OUT/bench_notma.json | python -c "$pick"
echo "== CUB reference point"; ./scripts/microbench/cub_sort_bench | tee $OUT/cub_sort.jsonl
echo "== SFB_GEOM_MINB3=1"; SFB_GEOM_MINB3=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline | tee $OUT/bench_minb3.json | python -c "$pick"
