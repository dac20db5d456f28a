#!/bin/bash
# Session 18: factored SH-gradient mode + sfb_sh_grad_combine, fused activations (8f-3), distCUDA2 (8f-5): GPU tests,
# quick perf of the main path (geom_backward must not regress) and of the new rows.
TAG=${1:-s18}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -25 $OUT/pytest_gpu.log | cut -c1-400
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 120 python scripts/quick_perf.py --config lego_1m > $OUT/quick_perf.jsonl; cut -c1-1500 $OUT/quick_perf.jsonl
timeout 400 python scripts/quick_perf_next_rows.py --iters 15 --warmup 3 > $OUT/quick_perf_next_rows.jsonl 2> $OUT/quick_perf_next_rows.err; cut -c1-600 $OUT/quick_perf_next_rows.jsonl; tail -5 $OUT/quick_perf_next_rows.err
