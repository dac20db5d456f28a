"""CPU: the reference's OWN gaussian_renderer/__init__.py, loaded unmodified from /root/reference, runs against the
alias package `diff_gaussian_rasterization` at the repo root, and drives the rasterizer with exactly the calls our
mirror splatfields_b200.renderer.render() makes (settings tuples, tensors, call order, output dict).  The native call
underneath (rasterize_gaussians) is replaced by a recorder — there is no GPU here; the kernels behind it are covered
by the -m gpu suite.  Skipped where the reference tree is not mounted (the GPU box)."""
import importlib.util
import math
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/gaussian_renderer/__init__.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")


class _Cam:
    def __init__(self, H, W):
        g = torch.Generator().manual_seed(4)
        self.FoVx, self.FoVy = 0.7, 0.55
        self.image_height, self.image_width = H, W
        self.world_view_transform = torch.randn(4, 4, generator=g)
        self.full_proj_transform = torch.randn(4, 4, generator=g)
        self.camera_center = torch.randn(3, generator=g)


class _Pipe:
    debug = False


def _load_reference_render(monkeypatch):
    # the reference module imports scene.gaussian_model (plyfile, simple_knn, ...) only for a type name
    scene = types.ModuleType("scene")
    gm = types.ModuleType("scene.gaussian_model")
    gm.GaussianModel = object
    monkeypatch.setitem(sys.modules, "scene", scene)
    monkeypatch.setitem(sys.modules, "scene.gaussian_model", gm)
    spec = importlib.util.spec_from_file_location("ref_gaussian_renderer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _record(monkeypatch):
    from splatfields_b200 import rasterizer
    calls = []

    def fake(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
             with_alpha=False):
        calls.append(dict(means3D=means3D, means2D=means2D, sh=sh, colors_precomp=colors_precomp, opacities=opacities,
                          scales=scales, rotations=rotations, cov=cov3Ds_precomp, rs=raster_settings, alpha=with_alpha))
        H, W = raster_settings.image_height, raster_settings.image_width
        k = float(len(calls))
        return (torch.full((3, H, W), k), torch.arange(means3D.shape[0], dtype=torch.int32) % 3,
                torch.full((1, H, W), 10 + k))
    monkeypatch.setattr(rasterizer, "rasterize_gaussians", fake)
    # gaussian_renderer/__init__.py:49 allocates on device="cuda": this box has none
    real = torch.zeros_like
    monkeypatch.setattr(torch, "zeros_like", lambda t, **kw: real(t, **{k: v for k, v in kw.items() if k != "device"}))
    return calls


def _same(a, b):
    if torch.is_tensor(a) or torch.is_tensor(b):
        return torch.is_tensor(a) and torch.is_tensor(b) and a.shape == b.shape and torch.equal(a, b)
    return a == b


@pytest.mark.parametrize("colour", ["sh", "rgb", "rgb_fnc"])
@pytest.mark.parametrize("return_opacity", [True, False])
def test_reference_render_and_mirror_make_the_same_rasterizer_calls(monkeypatch, colour, return_opacity):
    import diff_gaussian_rasterization as alias
    from splatfields_b200 import rasterizer, renderer
    assert alias.GaussianRasterizer is rasterizer.GaussianRasterizer
    ref = _load_reference_render(monkeypatch)
    assert ref.GaussianRasterizer is rasterizer.GaussianRasterizer      # the reference file resolved OUR classes
    calls = _record(monkeypatch)
    P, H, W = 11, 6, 9
    g = torch.Generator().manual_seed(1)
    gd = dict(means3D=torch.randn(P, 3, generator=g), active_sh_degree=2, gaussian_opacity=torch.rand(P, 1, generator=g),
              gaussian_scales=torch.rand(P, 3, generator=g), gaussian_rotations=torch.randn(P, 4, generator=g))
    if colour == "sh":
        gd["gaussian_features"] = torch.randn(P, 16, 3, generator=g)
    elif colour == "rgb":
        gd["gaussian_rgb"] = torch.rand(P, 3, generator=g)
    else:
        gd["gaussian_rgb_fnc"] = lambda ray_d: 0.5 + 0.5 * ray_d
    cam, bg = _Cam(H, W), torch.tensor([1.0, 0.5, 0.25])
    out_ref = ref.render(cam, gd, _Pipe(), bg, scaling_modifier=1.3, return_opacity=return_opacity)
    n_ref = len(calls)
    out_mir = renderer.render(cam, gd, _Pipe(), bg, scaling_modifier=1.3, return_opacity=return_opacity)
    a, b = calls[:n_ref], calls[n_ref:]
    assert n_ref == (2 if return_opacity else 1) and len(b) == n_ref
    for ca, cb in zip(a, b):
        for k in ("means3D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov", "alpha"):
            assert _same(ca[k], cb[k]), k
        assert ca["means2D"].shape == cb["means2D"].shape and ca["means2D"].requires_grad and cb["means2D"].requires_grad
        assert ca["rs"]._fields == cb["rs"]._fields
        for f in ca["rs"]._fields:
            assert _same(getattr(ca["rs"], f), getattr(cb["rs"], f)), f
    # both passes of one render() share ONE screen-space tensor (its .grad sums both passes)
    if return_opacity:
        assert a[0]["means2D"] is a[1]["means2D"] and b[0]["means2D"] is b[1]["means2D"]
        assert torch.equal(a[1]["rs"].bg, bg * 0.0) and a[1]["sh"] is None
        assert torch.equal(a[1]["colors_precomp"], torch.ones(P, 3))
    assert math.isclose(a[0]["rs"].tanfovx, math.tan(0.35)) and a[0]["rs"].scale_modifier == 1.3
    # output dicts: same keys, same values (the recorder numbers its calls, so call order shows in the images)
    assert list(out_ref.keys()) == list(out_mir.keys())
    for k in out_ref:
        if k == "viewspace_points":
            assert out_ref[k].shape == out_mir[k].shape == (P, 3)
        elif out_ref[k] is None:
            assert out_mir[k] is None
        else:
            va, vb = out_ref[k], out_mir[k]
            if k in ("render", "depth", "opacity"):
                vb = vb - float(n_ref)            # the mirror's calls are numbered after the reference's
            assert va.shape == vb.shape and torch.equal(va.float(), vb.float()), k
