"""GPU (-m gpu): the NVLink gradient exchange (csrc/exchange.cu, geom_backward_kernel<PUSH>) with 2, 3 and 4 ranks
emulated on one device — tests/exchange_emulation.py, run in a process of its own."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_emulated_ranks_match_the_serial_sum(cuda_lib):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "exchange_emulation.py")], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and len(line["cases"]) == 10 and sum(bool(c.get("fused")) for c in line["cases"]) == 6, tail
