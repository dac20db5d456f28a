#!/bin/bash
# Round-2 multi-GPU session (N GPUs of one box): exchange equivalence on real NVLink, bench with each exchange,
# BASELINE config 4 sharded over the ranks.   bash scripts/r2/multi.sh <tag> <N>
TAG=${1:-r2m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
NCCL_DEBUG=WARN timeout 400 $TR scripts/check_exchange.py > $OUT/check_exchange_n$N.json 2> $OUT/check_exchange_n$N.err; echo "check rc=$?"; grep '^{' $OUT/check_exchange_n$N.json | cut -c1-1500; tail -5 $OUT/check_exchange_n$N.err | cut -c1-300
run_bench() { # name, env...
  local name=$1; shift
  env "$@" BENCH_WATCHDOG_S=330 timeout 350 $TR bench.py --gpus $N --steps 100 --warmup 10 > $OUT/bench_n${N}_$name.json 2> $OUT/bench_n${N}_$name.err
  python - "$name" $OUT/bench_n${N}_$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d['value'],1), "ms", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), "exchange", d['exchange'])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -2 $OUT/bench_n${N}_$name.err | cut -c1-300
}
run_bench nvlink SFB_X=0
run_bench nvlink_p2p SFB_XCHG_NO_MULTICAST=1
run_bench factored SFB_EXCHANGE=factored
ROUNDS=$((1800 / N))
timeout 400 $TR scripts/run_view_time.py --rounds $ROUNDS > $OUT/view_time_n$N.json 2> $OUT/view_time_n$N.err; grep '^{' $OUT/view_time_n$N.json; tail -2 $OUT/view_time_n$N.err | cut -c1-300
timeout 300 $TR scripts/run_view_time.py --rounds $ROUNDS --exchange allreduce > $OUT/view_time_allreduce_n$N.json 2> $OUT/view_time_allreduce_n$N.err; grep '^{' $OUT/view_time_allreduce_n$N.json
ls $OUT
