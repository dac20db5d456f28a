#!/bin/bash
# Round-2 multi-GPU session, lean (GPU-minutes are charged N x):  bash scripts/r2/multi2.sh <tag> <N> [full|lean]
# full: exchange emulation test, exchange equivalence on real NVLink, bench with every exchange, configs[4] both ways.
# lean: exchange equivalence, bench with the default exchange, configs[4] with the default exchange.
TAG=${1:-r2m}
N=${2:-2}
MODE=${3:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$MODE" = full ]; then
  timeout 300 python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -12 | cut -c1-1200 | tee $OUT/pytest_exchange.log
fi
NCCL_DEBUG=WARN timeout 300 $TR scripts/check_exchange.py > $OUT/check_exchange_n$N.json 2> $OUT/check_exchange_n$N.err; echo "check rc=$?"; grep '^{' $OUT/check_exchange_n$N.json | cut -c1-1800; tail -5 $OUT/check_exchange_n$N.err | cut -c1-300
run_bench() { # name, env...
  local name=$1; shift
  env "$@" BENCH_WATCHDOG_S=280 timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > $OUT/bench_n${N}_$name.json 2> $OUT/bench_n${N}_$name.err
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
if [ "$MODE" = lean8 ]; then run_bench factored SFB_EXCHANGE=factored; fi
if [ "$MODE" = full ]; then
  run_bench nvlink_p2p SFB_XCHG_NO_MULTICAST=1
  run_bench factored SFB_EXCHANGE=factored
fi
ROUNDS=$((1800 / N))
timeout 300 $TR scripts/run_view_time.py --rounds $ROUNDS > $OUT/view_time_n$N.json 2> $OUT/view_time_n$N.err; grep '^{' $OUT/view_time_n$N.json; tail -2 $OUT/view_time_n$N.err | cut -c1-300
if [ "$MODE" = full ]; then
  timeout 300 $TR scripts/run_view_time.py --rounds $ROUNDS --exchange allreduce > $OUT/view_time_allreduce_n$N.json 2> $OUT/view_time_allreduce_n$N.err; grep '^{' $OUT/view_time_allreduce_n$N.json
fi
ls $OUT
