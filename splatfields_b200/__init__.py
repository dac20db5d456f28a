"""splatfields_b200 — B200-native (sm_100a) differentiable Gaussian-splat rasterizer: a drop-in for the
`diff_gaussian_rasterization` extension behind the reference's gaussian_renderer.render().
See DESIGN.md for the path, the boundary and the kernels."""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401
from .renderer import render  # noqa: F401
from .losses import l1_loss, ssim, photometric_loss  # noqa: F401
from .densify import add_densification_stats, densify_masks  # noqa: F401
from .activations import activate_parameters  # noqa: F401
from .knn import distCUDA2  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "render",
           "l1_loss", "ssim", "photometric_loss", "add_densification_stats", "densify_masks",
           "activate_parameters", "distCUDA2"]
