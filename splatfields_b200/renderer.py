"""render(): same contract as the reference's gaussian_renderer.render()
(gaussian_renderer/__init__.py:30-124) — takes a camera + `gaussian_dict`, returns
{"render", "viewspace_points", "visibility_filter", "radii", "opacity", "depth"} — built on the sm_100a
rasterizer.  The reference file itself also runs unmodified against the `diff_gaussian_rasterization`
alias package at the repo root; this module exists so that the path can be exercised without the
reference tree (tests, bench, smoke on the GPU box).
"""
from __future__ import annotations

import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


class _Pipe:
    debug = False


def render(viewpoint_camera, gaussian_dict: dict, pipe=None, bg_color: torch.Tensor = None,
           scaling_modifier: float = 1.0, return_opacity: bool = True, fused_alpha: bool = False):
    """fused_alpha=False reproduces the reference call for call (two rasterizer passes when
    return_opacity); fused_alpha=True obtains the opacity image from the SAME pass."""
    pipe = pipe or _Pipe()
    means3D = gaussian_dict["means3D"]
    active_sh_degree = gaussian_dict["active_sh_degree"]
    opacity = gaussian_dict["gaussian_opacity"]
    scales = gaussian_dict["gaussian_scales"]
    rotations = gaussian_dict["gaussian_rotations"]
    features = gaussian_dict.get("gaussian_features", None)
    rgb = gaussian_dict.get("gaussian_rgb", None)
    if rgb is None and "gaussian_rgb_fnc" in gaussian_dict:           # reference :43-46
        ray_d = means3D - viewpoint_camera.camera_center[None]
        ray_d = ray_d / torch.norm(ray_d, dim=-1, keepdim=True)
        rgb = gaussian_dict["gaussian_rgb_fnc"](ray_d)

    # dummy leaf that receives the screen-space mean gradients (reference :49-53)
    screenspace_points = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)

    def settings(bg):
        return GaussianRasterizationSettings(
            image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
            tanfovx=tanfovx, tanfovy=tanfovy, bg=bg, scale_modifier=scaling_modifier,
            viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
            sh_degree=active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False,
            debug=pipe.debug)

    rasterizer = GaussianRasterizer(raster_settings=settings(bg_color))
    opacity_image = None
    if return_opacity and fused_alpha:
        rendered_image, radii, depth, opacity_image = rasterizer(
            means3D=means3D, means2D=screenspace_points, shs=features, colors_precomp=rgb, opacities=opacity,
            scales=scales, rotations=rotations, cov3D_precomp=None, with_alpha=True)
        return {"render": rendered_image, "viewspace_points": screenspace_points,
                "visibility_filter": radii > 0, "radii": radii, "opacity": opacity_image, "depth": depth}
    rendered_image, radii, depth = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=features, colors_precomp=rgb, opacities=opacity,
        scales=scales, rotations=rotations, cov3D_precomp=None)
    if return_opacity:                                                # reference :104-115
        rasterizer_mask = GaussianRasterizer(raster_settings=settings(bg_color * 0.0))
        opacity_image = rasterizer_mask(
            means3D=means3D, means2D=screenspace_points, shs=None,
            colors_precomp=torch.ones(opacity.shape[0], 3, device=opacity.device), opacities=opacity,
            scales=scales, rotations=rotations, cov3D_precomp=None)[0]
        opacity_image = opacity_image[:1]
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii, "opacity": opacity_image, "depth": depth}
