"""GPU (-m gpu): oracle parity at BASELINE.json's FULL sizes — the workloads bench.py and quick_perf time.

Same bars as tests/test_gpu_parity.py (`_compare`): bit-exact radii / rectangles / depth bits / conics / colours /
point_list_keys / point_list / ranges, 1e-5 abs on RGB / depth away from decision thresholds, 1e-3 rel on every
gradient.  `owlii_2m` (2 M splats, 1080p: ceil(log2 P) + ceil(log2 T) = 34 > 32) is the only configuration that takes
the UNPACKED instance path and the precomputed-RGB colour path at full size."""
import os
import subprocess
import sys

import numpy as np
import pytest

from splatfields_b200 import synth
from tests.test_gpu_parity import _compare

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FULL = [
    # config,    camera, smallest accepted fraction of pixels away from a decision threshold
    ("lego_1m",  0, 0.99),
    ("lego_1m",  3, 0.99),
    ("dtu_500k", 0, 0.99),
    ("owlii_2m", 0, 0.99),
]


@pytest.mark.parametrize("name,k,min_ok", FULL, ids=[f"{n}_cam{k}" for n, k, _ in FULL])
def test_parity_full_size(oracle, cuda_lib, name, k, min_ok):
    cfg = synth.CONFIGS[name]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=cfg["scale_mult"], precomp_rgb=cfg["precomp_rgb"])
    cam = synth.config_camera(name, k)
    deg = 0 if cfg["precomp_rgb"] else 3
    f, c = _compare(oracle, sc, cam, cfg["H"], cfg["W"], deg, min_ok=min_ok)
    assert f["num_rendered"] > 1_000_000


# Every build / run-time variant that ships goes through the same parity file in a process of its own (the knobs are
# read once per process).  SFB_NO_PACK=1: unpacked (tile, index) instance pairs on small scenes; exactexp: the
# library built with expf instead of ex2.approx in the compositing kernels.
VARIANTS = [
    ("no_pack", {"SFB_NO_PACK": "1"}),
    ("exactexp", {"SFB_LIB_VARIANT": "exactexp"}),
    # the measured A/B arms of round 2 that stay in the library
    ("tma_staging", {"SFB_FWD_STAGE": "tma", "SFB_BWD_STAGE": "tma"}),
    ("bwd_b128_pred", {"SFB_BWD_BATCH": "128", "SFB_BWD_SWEEP": "pred"}),
    ("bwd_b128x3_launch_order", {"SFB_BWD_BATCH": "128x3", "SFB_BWD_ORDER": "0"}),
]


@pytest.mark.parametrize("name,env", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_variant_passes_the_parity_file(cuda_lib, name, env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu",
                        "-x", "-q", "-p", "no:cacheprovider"], cwd=ROOT, env=e, capture_output=True, text=True,
                       timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], tail
