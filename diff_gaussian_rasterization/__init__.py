"""Drop-in alias: with the repo root on sys.path, the reference's own
`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(gaussian_renderer/__init__.py:14) resolves to the sm_100a rasterizer unmodified."""
from splatfields_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                         rasterize_gaussians)
