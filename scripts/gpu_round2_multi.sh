#!/bin/bash
# Multi-GPU session of round 2 (N GPUs, default 2): NCCL equivalence check, then bench.py with the factored exchange
# with and without the early all-gather (SFB_EARLY_GATHER=1), then config 4 sharded over the ranks.
TAG=${1:-r2m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_exchange_nccl.py > $OUT/check_exchange_n$N.json 2> $OUT/check_exchange_n$N.err; echo "check rc=$?"; grep '^{' $OUT/check_exchange_n$N.json
SFB_EARLY_GATHER=1 timeout 300 $TR scripts/check_exchange_nccl.py > $OUT/check_exchange_early_n$N.json 2> $OUT/check_exchange_early_n$N.err; echo "check(early) rc=$?"; grep '^{' $OUT/check_exchange_early_n$N.json
for early in 0 1; do
  SFB_EARLY_GATHER=$early BENCH_WATCHDOG_S=280 timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > $OUT/bench_n${N}_early$early.json 2> $OUT/bench_n${N}_early$early.err
  python -c "
import json,sys
d=json.loads(open('$OUT/bench_n${N}_early$early.json').read().strip().splitlines()[-1]); print('early=$early', d['value'], d['ms_per_step'], d['exchange'])"
done
timeout 300 $TR scripts/run_view_time.py --rounds 24 > $OUT/view_time_n$N.json 2> $OUT/view_time_n$N.err; grep '^{' $OUT/view_time_n$N.json
