"""CPU: the C-ABI library loads and exports every symbol include/splat_b200.h declares; the Python
mirror keeps the reference's names, field order and error messages.  No compute calls (no GPU here)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    from splatfields_b200 import build
    build.build()


def test_library_exports_every_declared_symbol():
    _built()
    from splatfields_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "splat_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sfb_[a-z_0-9]+)\s*\(", hdr)) - {"sfb_alloc_fn"})
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.sfb_abi_version() == 6
    assert lib.sfb_last_error() == b""


def test_committed_sass_manifest_describes_the_built_library():
    """splatfields_b200/sass_manifest.json (checked by __graft_entry__.build() after a forced rebuild) is up to date:
    regenerate it with `python -m splatfields_b200.build --write-manifest` after touching a kernel."""
    _built()
    from splatfields_b200 import build
    assert build.check_manifest()["kernels"] > 50


def test_loss_window_is_the_reference_window_bit_for_bit():
    """Host-side piece of the fused loss: the 11 weights equal loss_utils.gaussian(11, 1.5) as the reference computed
    them (golden file), and the scratch size follows the documented layout (3 maps + block partials)."""
    import ctypes as C
    import numpy as np
    _built()
    from splatfields_b200 import _lib
    lib = _lib.load()
    buf = (C.c_float * 11)()
    lib.sfb_loss_window(buf)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "next_rows.npz"))
    assert np.array_equal(np.array(list(buf), dtype=np.float32), gold["window_1d"])
    n = lib.sfb_loss_scratch_bytes(3, 800, 800)
    assert 3 * 3 * 800 * 800 * 4 <= n <= 3 * 3 * 800 * 800 * 4 + (1 << 20)
    assert lib.sfb_loss_scratch_bytes(0, 800, 800) == 0


def test_next_row_mirrors_refuse_cpu_tensors():
    from splatfields_b200 import losses, densify
    from splatfields_b200._lib import SplatB200Error
    with pytest.raises(SplatB200Error, match="no CPU fallback"):
        losses.photometric_loss(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8), 0.2)
    with pytest.raises(Exception, match="CUDA tensor"):
        densify.add_densification_stats(torch.zeros(4, 1), torch.zeros(4, 1), torch.zeros(4, 3), radii=torch.ones(4))
    from splatfields_b200 import activate_parameters, distCUDA2, rasterizer
    with pytest.raises(SplatB200Error, match="no CPU fallback"):
        activate_parameters(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 4), torch.zeros(4, 1),
                            torch.zeros(4, 1, 3), torch.zeros(4, 15, 3))
    with pytest.raises(SplatB200Error, match="no CPU fallback"):
        distCUDA2(torch.zeros(8, 3))
    with pytest.raises(SplatB200Error, match="no CPU fallback"):
        rasterizer.sh_grad_combine(torch.zeros(4, 3), torch.zeros(1, 3), torch.zeros(1, 4, 3), 3, torch.zeros(4, 16, 3))


def test_library_is_sm100a_only():
    _built()
    import subprocess
    from splatfields_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_settings_fields_match_reference_order():
    from splatfields_b200 import GaussianRasterizationSettings
    # gaussian_renderer/__init__.py:59-72 passes exactly these keywords
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")


def test_alias_package_resolves_like_the_reference_import():
    import diff_gaussian_rasterization as D
    from splatfields_b200 import rasterizer
    assert D.GaussianRasterizer is rasterizer.GaussianRasterizer
    assert D.GaussianRasterizationSettings is rasterizer.GaussianRasterizationSettings


def _rast():
    from splatfields_b200 import GaussianRasterizationSettings, GaussianRasterizer
    rs = GaussianRasterizationSettings(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                       torch.zeros(3), False, False)
    return GaussianRasterizer(rs)


def test_argument_validation_messages():
    r = _rast()
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=torch.ones(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=torch.ones(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=torch.zeros(4, 3),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=torch.ones(4, 1), colors_precomp=torch.zeros(4, 3))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=torch.ones(4, 1), colors_precomp=torch.zeros(4, 3),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    """CPU tensors must fail loudly, not silently run somewhere else."""
    from splatfields_b200._lib import SplatB200Error
    r = _rast()
    m = torch.zeros(4, 3)
    with pytest.raises(SplatB200Error, match="no CPU fallback"):
        r(means3D=m, means2D=m, opacities=torch.ones(4, 1), colors_precomp=torch.zeros(4, 3),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "splatfields_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "splat_oracle" not in txt, f


def test_exchange_buffer_layout_sizes():
    """sfb_xchg_bytes (host-only arithmetic, no GPU): the symmetric buffer holds the flag page, the chunk flags
    ((16 + 1) words per 1024 splats), the packed records for both step parities and — with shs — the colour tables
    for both parities and every rank; everything 256-byte aligned; identical on every rank (no rank argument)."""
    _built()
    from splatfields_b200 import _lib
    lib = _lib.load()
    al = lambda v: (v + 255) // 256 * 256
    for P, world, ngeo, with_gc in ((1_000_000, 8, 12, 1), (1_000_000, 2, 12, 1), (2_000_000, 8, 16, 0), (7_777, 3, 12, 1),
                                    (1, 2, 16, 0)):
        nch = (P + 1023) // 1024
        want = 256 + al(17 * nch * 4) + 2 * al(P * ngeo * 4)
        if with_gc:
            want += 2 * al(world * al(P * 3 * 4))
        assert lib.sfb_xchg_bytes(P, world, ngeo, with_gc) == want, (P, world, ngeo, with_gc)
    assert lib.sfb_xchg_bytes(100, 2, 13, 1) == 0          # records are 12 (shs) or 16 (colors_precomp) floats
    d = _lib.XchgDesc()
    assert hasattr(d, "campos_views") and _lib.XCHG_MAX_RANKS == 16
