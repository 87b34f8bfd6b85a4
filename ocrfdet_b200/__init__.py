"""ocrfdet_b200 -- Blackwell-native implementation of OcRFDet's Gaussian render path.

The package holds only what the hot path needs:
  csrc/           hand-written sm_100a CUDA kernels + the C ABI (include/ocrf_raster.h)
  rasterizer.py   host-side mirror of the reference plugin (`diff_gaussian_rasterization`)
  opacity_lift.py stage 5, the HOA opacity mask
  cameras.py      camera-matrix conventions of the caller (the input contract)
  sharding.py     (sample, view) partition over ranks + the opacity-map all-gather
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, render_batch,  # noqa
                         pack_cameras, pack_camera_dicts)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "render_batch",
           "pack_cameras", "pack_camera_dicts"]
