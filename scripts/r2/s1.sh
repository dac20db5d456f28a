#!/bin/bash
# Round-2 session 1 (1 GPU): network probe for the pinned reference rasterizer, device capability probe, smoke of the
# new staging paths, the GPU suite, the A/B matrix of the render-kernel changes, bench, ncu of the two render kernels.
TAG=${1:-r2s1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{
  echo "== date"; date -u
  echo "== pip download (pinned reference rasterizer, README.md:28)"
  timeout 25 python -m pip download --no-deps -d /tmp/refdl "git+https://github.com/ingra14m/depth-diff-gaussian-rasterization.git@f2d8fa9921ea9a6cb9ac1c33a34ebd1b11510657" 2>&1 | tail -4
  echo "== git clone"
  timeout 25 git clone --depth 1 https://github.com/ingra14m/depth-diff-gaussian-rasterization.git /tmp/refclone 2>&1 | tail -3
  echo "== curl github / pypi"
  timeout 15 curl -sS -m 10 -o /dev/null -w "%{http_code}\n" https://github.com 2>&1 | tail -1
  timeout 15 curl -sS -m 10 -o /dev/null -w "%{http_code}\n" https://pypi.org/simple/ 2>&1 | tail -1
  echo "== resolv / routes"
  cat /etc/resolv.conf 2>/dev/null | head -3
  (ip route 2>/dev/null || route -n 2>/dev/null) | head -5
  echo "== local copies"
  python -c "import diff_gaussian_rasterization" 2>&1 | tail -1
  find / -xdev \( -iname '*diff_gaussian*' -o -iname '*diff-gaussian*' -o -iname 'simple_knn*' \) -not -path '*/repo*' -not -path '/proc/*' 2>/dev/null | grep -v "$PWD" | head
} > $OUT/network_probe.log 2>&1
tail -22 $OUT/network_probe.log
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > $OUT/device.log 2>&1
nvidia-smi topo -m >> $OUT/device.log 2>&1
python - >> $OUT/device.log 2>&1 <<'PY'
import torch, os
from cuda import cuda
cuda.cuInit(0)
err, dev = cuda.cuDeviceGet(0)
for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
             "CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN", "CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_MULTIPROCESSOR",
             "CU_DEVICE_ATTRIBUTE_NUMA_ID", "CU_DEVICE_ATTRIBUTE_HOST_NUMA_ID"):
    a = getattr(cuda.CUdevice_attribute, name, None)
    print(name, cuda.cuDeviceGetAttribute(a, dev) if a is not None else "n/a")
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("lscpu | grep -i -E 'numa|model name|socket' | head -8")
from splatfields_b200.host_api import numa_local
with numa_local(torch.device("cuda:0")) as n:
    print("numa_local cpus:", getattr(n, "cpus", None) and (len(n.cpus), n.cpus[:4], n.cpus[-4:]))
PY
tail -16 $OUT/device.log
echo "== TMA row gather primitive"; timeout 200 python -m pytest tests/test_gpu_properties.py -m gpu -q -k tma_row 2>&1 | tail -4 | tee $OUT/tma_probe.log
# smoke of both staging paths first: a broken TMA path must not take the whole session down
for st in ldg tma; do echo "== smoke SFB_FWD_STAGE=$st SFB_BWD_STAGE=$st"; SFB_FWD_STAGE=$st SFB_BWD_STAGE=$st timeout 300 python __graft_entry__.py --smoke-only 2>&1 | tail -2; done | tee $OUT/smoke.log
echo "== smoke SFB_BWD_SWEEP=pred SFB_BWD_BATCH=128"; SFB_BWD_SWEEP=pred SFB_BWD_BATCH=128 timeout 300 python __graft_entry__.py --smoke-only 2>&1 | tail -2 | tee -a $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log | cut -c1-400
qp() { # name, env...
  local name=$1; shift
  env "$@" timeout 150 python scripts/quick_perf.py --config lego_1m --iters 30 > $OUT/qp_$name.json 2>$OUT/qp_$name.err
  python - "$name" $OUT/qp_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    f,b=d['fwd_stages'],d['bwd_stages']
    print("%-14s total %.3f fwd %.3f bwd %.3f | pre %.3f rfwd %.3f rbwd %.3f geom %.3f tsort %.3f dsort %.3f dup %.3f" % (sys.argv[1], d['total_ms'], d['fwd_ms'], d['bwd_ms'], f.get('preprocess',0), f.get('render_forward',0), b.get('render_backward',0), b.get('geom_backward',0), f.get('tile_sort.scatter',0)+f.get('tile_sort.hist',0), f.get('depth_sort.scatter',0)+f.get('depth_sort.hist',0), f.get('duplicate',0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
{
qp default SFB_X=0
qp fwd_tma SFB_FWD_STAGE=tma
qp bwd_tma SFB_BWD_STAGE=tma
qp bwd_pred SFB_BWD_SWEEP=pred
qp bwd_b128 SFB_BWD_BATCH=128
qp bwd_b128_pred SFB_BWD_BATCH=128 SFB_BWD_SWEEP=pred
qp bwd_b128x3 SFB_BWD_BATCH=128x3
qp bwd_b128x3_pred SFB_BWD_BATCH=128x3 SFB_BWD_SWEEP=pred
qp bwd_noorder SFB_BWD_ORDER=0
qp pre_occ4 SFB_PRE_OCC4=1
qp default2 SFB_X=0
} | tee $OUT/ab_matrix.txt
for c in dtu_500k owlii_2m lego_100k; do timeout 150 python scripts/quick_perf.py --config $c >> $OUT/quick_perf.jsonl; done; cut -c1-700 $OUT/quick_perf.jsonl
timeout 400 python bench.py --steps 100 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-600 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
KERNELS='render_forward_kernel|render_backward_mma_kernel|preprocess_kernel|geom_backward_kernel'
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -s 16 -c 4 -f -o $OUT/prof \
    python scripts/quick_perf.py --config lego_1m --iters 1 --warmup 3 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/ncu_raw.csv > $OUT/ncu_full_summary.txt 2>&1; grep -c "^==" $OUT/ncu_full_summary.txt
K='activate_forward_kernel|activate_backward_kernel|knn_query_kernel|knn_boxes_kernel|knn_morton_kernel|sh_grad_combine_kernel|ssim_stats_kernel|ssim_grad_kernel|densify_stats_kernel|densify_masks_kernel'
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 20 -f -o $OUT/prof_next_rows \
    python scripts/quick_perf_next_rows.py --iters 1 --warmup 0 > $OUT/ncu_next_rows.log 2>&1
ncu -i $OUT/prof_next_rows.ncu-rep --page raw --csv > $OUT/ncu_next_rows_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/ncu_next_rows_raw.csv > $OUT/ncu_next_rows_summary.txt 2>&1; grep -c "^==" $OUT/ncu_next_rows_summary.txt
rm -f $OUT/prof_next_rows.ncu-rep
ls $OUT | head -40
