#!/bin/bash
TAG=${1:-s10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log | cut -c1-300
for lb in 8 16 32; do for ch in 0 1; do
  echo "== lego_1m LB=$lb CH=$ch"; SFB_SORT_LB=$lb SFB_SORT_CHAINS=$ch timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_lb${lb}_ch${ch}.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['fwd_stages']
print('fwd_ms',round(d['fwd_ms'],4),'bwd_ms',round(d['bwd_ms'],4),{k:round(v*1e3,1) for k,v in f.items() if 'sort' in k})"
done; done
for c in dtu_500k owlii_2m; do for lb in 8 16; do echo "== $c LB=$lb"; SFB_SORT_LB=$lb timeout 120 python scripts/quick_perf.py --config $c | tee -a $OUT/quick_perf_${c}_lb$lb.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['fwd_stages']
print('fwd_ms',round(d['fwd_ms'],4),'bwd_ms',round(d['bwd_ms'],4),{k:round(v*1e3,1) for k,v in f.items() if 'sort' in k})"
done; done
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-400 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
