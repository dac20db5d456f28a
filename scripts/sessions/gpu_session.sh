#!/bin/bash
# One GPU-box session: tests, bench, launch list, full ncu capture of the hot kernels.
# Usage (under gpurun): bash scripts/gpu_session.sh <tag> [skip-tests]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ "${2:-}" != "skip-tests" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 3000 $OUT/bench_n1.json; tail -5 $OUT/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
for c in lego_100k dtu_500k owlii_2m; do python scripts/quick_perf.py --config $c >> $OUT/quick_perf.jsonl 2>> $OUT/quick_perf.err; done
tail -3 $OUT/quick_perf.jsonl
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 60 --csv --log-file $OUT/launches.csv \
    python scripts/quick_perf.py --config lego_1m --iters 2 --warmup 4 > $OUT/ncu_launches.log 2>&1
# full capture of the hot kernels (one launch each, after warm-up)
ncu --set full --clock-control none --import-source on \
    -k regex:'radix_scatter_kernel|render_forward_kernel|render_backward_kernel|geom_backward_kernel|preprocess_kernel|duplicate_kernel|radix_hist_kernel' \
    -s 60 -c 14 -o $OUT/prof python scripts/quick_perf.py --config lego_1m --iters 1 --warmup 3 > $OUT/ncu_full.log 2>&1
ls -la $OUT
