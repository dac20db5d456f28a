#!/bin/bash
# Multi-GPU session (run under gpurun --gpus N):  bash scripts/multi_gpu_session.sh <tag> <N> [full|mid|lean]
# exchange equivalence on real NVLink (scripts/check_exchange.py), bench with the default exchange (mid: + the fused
# one-kernel variant; full: + without multicast), BASELINE configs[4] sharded over the ranks.
TAG=${1:-r2f}
N=${2:-2}
MODE=${3:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
NCCL_DEBUG=WARN timeout 300 $TR scripts/check_exchange.py > $OUT/check_exchange_n$N.json 2> $OUT/check_exchange_n$N.err; echo "check rc=$?"; grep '^{' $OUT/check_exchange_n$N.json | cut -c1-2500; tail -5 $OUT/check_exchange_n$N.err | cut -c1-300
run_bench() { # name, env...
  local name=$1; shift
  env "$@" BENCH_WATCHDOG_S=280 timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > $OUT/bench_n${N}_$name.json 2> $OUT/bench_n${N}_$name.err
  python - "$name" $OUT/bench_n${N}_$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d['value'],1), "ms", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), "exchange", {k:v for k,v in d['exchange'].items() if k!='note'})
    print("   kernels", d['stages']['ms_per_step_by_kernel'])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -2 $OUT/bench_n${N}_$name.err | cut -c1-300
}
run_bench nvlink SFB_X=0
if [ "$MODE" = full ]; then run_bench nvlink_p2p SFB_XCHG_NO_MULTICAST=1; fi
if [ "$MODE" != lean ]; then run_bench fused SFB_XCHG_FUSED=1; fi
ROUNDS=$((1800 / N))
timeout 300 $TR scripts/run_view_time.py --rounds $ROUNDS > $OUT/view_time_n$N.json 2> $OUT/view_time_n$N.err; grep '^{' $OUT/view_time_n$N.json; tail -2 $OUT/view_time_n$N.err | cut -c1-300
ls $OUT
