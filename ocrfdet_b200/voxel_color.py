"""Voxel colouring and the sparse supervision image (scope row f-4: the callers upstream of the Gaussian heads).

Mirrors three methods of the OcRF view transformer
(/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py):
  * `lidar_points_to_image_values` (:924-942) + `color_voxels` (:945-971) -> `color_voxels_from_images`
  * `retain_valid_pixels` (:1004-1022)                                   -> `retain_valid_pixels`
with the reference's argument shapes.  Scope: the RGB path (`imgs_wo_norm` -> the colour head's `rgb` input,
:1065-1071), which carries no gradient in the reference.  The same two reference methods are also called on the
`alpha_img` planes (:1116-1118), where `grid_sample` back-propagates into the sigma MLP: `color_voxels_from_images`
REFUSES an input that requires grad instead of silently cutting that gradient.
There is no PyTorch fallback: both call libocrf_raster.so.
"""
import torch

from . import _lib

__all__ = ["color_voxels_from_images", "retain_valid_pixels"]


def _mask_u8(mask):
    m = mask.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).view(torch.uint8)


def color_voxels_from_images(pillars, imgs, mask, divisor=1.0):
    """pillars [B,N,P,Q,2] pixel coordinates (x, y) of the voxel centres in every camera, imgs [B,N,C,H,W], mask
    [B,N,P,Q,1] (bool) -> (avg_color [B,P,Q,C], valid_mask [B,P,Q] bool): what `color_voxels(voxels,
    lidar_points_to_image_values(pillars, imgs, mask), mask)` returns as its 2nd and 3rd value
    (`colored_voxels` is `cat(voxels, avg_color)`); `divisor=255.0` folds the caller's `/ 255.0` (:1071)."""
    L = _lib.lib()
    if not (pillars.is_cuda and imgs.is_cuda and mask.is_cuda):
        raise _lib.OcrfError("color_voxels_from_images: tensors must live on a CUDA device (there is no CPU path)")
    if imgs.requires_grad and torch.is_grad_enabled():
        raise _lib.OcrfError("color_voxels_from_images has no backward: it covers the RGB path (no gradient in the "
                             "reference); for the alpha planes of view_transformer_ocrf.py:1116-1118 keep the "
                             "differentiable torch formulation or detach explicitly")
    B, N, P, Q, two = pillars.shape
    if two != 2 or tuple(mask.shape[:4]) != (B, N, P, Q):
        raise ValueError("pillars must be [B,N,P,Q,2] and mask [B,N,P,Q,1]")
    _, _, C, H, W = imgs.shape
    if imgs.shape[0] != B or imgs.shape[1] != N:
        raise ValueError("imgs must be [B,N,C,H,W] with the B, N of pillars")
    coords = pillars.contiguous().float()
    images = imgs.contiguous().float()
    m8 = _mask_u8(mask)
    avg = torch.empty((B, P, Q, C), dtype=torch.float32, device=pillars.device)
    valid = torch.empty((B, P, Q), dtype=torch.uint8, device=pillars.device)
    _lib.check(L.ocrf_color_voxels(_lib.current_stream(), B, N, P * Q, C, H, W, _lib.ptr(coords), _lib.ptr(m8),
                                   _lib.ptr(images), float(divisor), _lib.ptr(avg), _lib.ptr(valid)), "ocrf_color_voxels")
    return avg, valid.view(torch.bool)


def retain_valid_pixels(image_matrix, pseudo_point_cloud, mask, fill=255.0):
    """image_matrix [B,N,C,H,W], pseudo_point_cloud [B,N,...,2] (x, y) pixel coordinates, mask [B,N,...,1] (bool)
    -> [B,N,C,H,W]: `fill` everywhere except the pixels hit by a visible point, which keep the image value."""
    L = _lib.lib()
    if not (image_matrix.is_cuda and pseudo_point_cloud.is_cuda and mask.is_cuda):
        raise _lib.OcrfError("retain_valid_pixels: tensors must live on a CUDA device (there is no CPU path)")
    B, N, C, H, W = image_matrix.shape
    if pseudo_point_cloud.shape[-1] != 2 or pseudo_point_cloud.shape[:2] != (B, N):
        raise ValueError("pseudo_point_cloud must be [B,N,...,2]")
    img = image_matrix.contiguous().float()
    coords = pseudo_point_cloud.contiguous().float().reshape(B * N, -1, 2)
    m8 = _mask_u8(mask).reshape(B * N, -1)
    if m8.shape[1] != coords.shape[1]:
        raise ValueError("mask must have one entry per point")
    out = torch.empty_like(img)
    _lib.check(L.ocrf_retain_valid_pixels(_lib.current_stream(), B * N, coords.shape[1], C, H, W, _lib.ptr(coords),
                                          _lib.ptr(m8), _lib.ptr(img), float(fill), _lib.ptr(out)),
               "ocrf_retain_valid_pixels")
    return out.to(image_matrix.dtype)
