#!/bin/bash
# Session 19 (N GPUs, default 2): NCCL correctness of the factored gradient exchange, then bench.py at N ranks with
# the factored exchange (default) and with the plain all-reduce of the whole slab (A/B).
TAG=${1:-s19}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_exchange_nccl.py > $OUT/check_exchange_n$N.json 2> $OUT/check_exchange_n$N.err; echo "check rc=$?"; cat $OUT/check_exchange_n$N.json; tail -3 $OUT/check_exchange_n$N.err | cut -c1-300
BENCH_WATCHDOG_S=280 timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; cut -c1-400 $OUT/bench_n$N.json; tail -2 $OUT/bench_n$N.err | cut -c1-300
SFB_EXCHANGE=allreduce BENCH_WATCHDOG_S=280 timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > $OUT/bench_n${N}_allreduce.json 2> $OUT/bench_n${N}_allreduce.err; cut -c1-400 $OUT/bench_n${N}_allreduce.json; tail -2 $OUT/bench_n${N}_allreduce.err | cut -c1-300
python - <<PY
import json
for f in ("bench_n$N.json", "bench_n${N}_allreduce.json"):
    try:
        d = json.loads(open("$OUT/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("exchange"))
    except Exception as e:
        print(f, "unreadable", e)
PY
