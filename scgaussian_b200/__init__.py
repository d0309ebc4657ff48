"""scgaussian_b200 -- B200-native (sm_100a) differentiable Gaussian rasterizer: the hot path of
SCGaussian's gaussian_renderer.render() (reference gaussian_renderer/__init__.py:20-118) rebuilt
from scratch behind the reference's own operator surface.

    from scgaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer

or, as a literal drop-in for the package the reference imports,

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

Everything computes in libscgr.so (include/scgr.h, built by scgaussian_b200/build.py); there is no
CPU fallback.
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, ScgrError, mark_visible,  # noqa: F401
                         rasterize_gaussians, rasterize_forward_raw, rasterize_backward_raw)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "ScgrError", "mark_visible",
           "rasterize_gaussians", "rasterize_forward_raw", "rasterize_backward_raw"]
__version__ = "0.1.0"
