"""Drop-in module name: SCGaussian does
`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(reference gaussian_renderer/__init__.py:15).  With this repo on PYTHONPATH that import resolves
here and the reference's render()/train.py/render.py run unchanged on the B200-native rasterizer."""
from scgaussian_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                        _RasterizeGaussians, mark_visible, rasterize_gaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "mark_visible"]
