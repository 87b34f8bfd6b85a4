"""ocrfdet_b200 -- Blackwell-native implementation of OcRFDet's Gaussian render path.

The package holds only what the hot path needs:
  csrc/           hand-written sm_100a CUDA kernels + the C ABI (include/ocrf_raster.h)
  rasterizer.py   host-side mirror of the reference plugin (`diff_gaussian_rasterization`)
  opacity_lift.py stage 5, the HOA opacity mask (and BEVGeomAttention's gate: the same op)
  graphs.py       CUDA-graph replay of a whole forward + backward step
  cameras.py      camera-matrix conventions of the caller (the input contract)
  sharding.py     (sample, view) partition over ranks + the opacity-map all-gather
and the callers either side of the rasterizer ("next" rows of the scope table), behind the reference's names:
  gaussian_heads.py  the S/R/A/C MLP heads that turn voxel features into Gaussian parameters
  bev_pool.py        bev_pool_v2 (forward, backward incl. the regrouping of the point list)
  voxel_color.py     voxel colouring (grid_sample + masked mean over cameras) and retain_valid_pixels
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, render_batch,  # noqa
                         pack_cameras, pack_camera_dicts)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "render_batch",
           "pack_cameras", "pack_camera_dicts"]
