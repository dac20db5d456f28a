#!/bin/bash
# Round-2 last measurements (1 GPU, ~1 min): parity + timings of the 4-CTA/SM tile-sort pass (bare keys; second call: split instances too).
mkdir -p gpurun_out/r2s12
timeout 60 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py -m gpu -q -x -k "owlii or no_pack or small" > gpurun_out/r2s12/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s12/pytest.log; tail -3 gpurun_out/r2s12/pytest.log | cut -c1-300
timeout 40 python scripts/quick_perf.py --config owlii_2m >> gpurun_out/r2s12/quick_perf.jsonl 2>/dev/null; cut -c1-800 gpurun_out/r2s12/quick_perf.jsonl
