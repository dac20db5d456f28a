#!/bin/bash
# Round-2 session 1 (1 GPU): network probe for the pinned reference rasterizer, device capability probe,
# the GPU suite as it stands, and the A/Bs left open by round 1 (preprocess occupancy, tight rectangles).
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
  find / -xdev \( -iname '*diff_gaussian*' -o -iname '*diff-gaussian*' -o -iname 'simple_knn*' \) -not -path '*/repo*' 2>/dev/null | head
} > $OUT/network_probe.log 2>&1
tail -30 $OUT/network_probe.log
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > $OUT/device.log 2>&1
nvidia-smi topo -m >> $OUT/device.log 2>&1
python - >> $OUT/device.log 2>&1 <<'PY'
import torch
from cuda import cuda
cuda.cuInit(0)
err, dev = cuda.cuDeviceGet(0)
for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
             "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED", "CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_SUPPORTED",
             "CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN", "CU_DEVICE_ATTRIBUTE_NUMA_ID", "CU_DEVICE_ATTRIBUTE_HOST_NUMA_ID"):
    a = getattr(cuda.CUdevice_attribute, name, None)
    if a is None:
        print(name, "n/a"); continue
    print(name, cuda.cuDeviceGetAttribute(a, dev))
import os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("lscpu | grep -i -E 'numa|model name|socket' | head -8")
try:
    import torch.distributed._symmetric_memory as sm
    print("symmetric_memory:", [n for n in dir(sm) if not n.startswith('__')][:60])
except Exception as e:
    print("symm import failed", e)
PY
cat $OUT/device.log | tail -40
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log | cut -c1-300
for v in 0 1; do echo "== SFB_PRE_OCC4=$v"; SFB_PRE_OCC4=$v timeout 120 python scripts/quick_perf.py --config lego_1m | tee -a $OUT/quick_perf_pre_occ4_$v.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['fwd_ms'], d['bwd_ms'], d['fwd_stages'], d['bwd_stages'])"; done
for c in lego_1m dtu_500k; do timeout 300 python scripts/ab_tight_rect.py --config $c >> $OUT/ab_tight_rect.jsonl 2>> $OUT/ab_tight_rect.err; done; cat $OUT/ab_tight_rect.jsonl | cut -c1-600
for c in dtu_500k owlii_2m; do timeout 120 python scripts/quick_perf.py --config $c >> $OUT/quick_perf.jsonl; done; cut -c1-900 $OUT/quick_perf.jsonl
timeout 300 python bench.py --steps 100 --warmup 10 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-400 $OUT/bench_n1.json
