"""Drop-in import name of the reference plugin.

OcRFDet's render wrapper does `from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer` (/root/reference/mmdet3d/models/necks/MVSGaussian/lib/gaussian_renderer/__init__.py:14).
With this repository on PYTHONPATH that import resolves here and runs the B200-native kernels.
"""
from ocrfdet_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                     rasterize_gaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
