"""Stage 5: the opacity mask of OcRFDet's height-aware opacity (HOA) lift, as one fused op.

Mirrors `ObatinOpacityMask.forward` + its application
(/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:230-242, 1197-1199):

    mask = sigmoid(conv2d(cat(mean_c(x), max_c(x)), weight, padding=K//2) + opacity_bev);  out = x * mask

`OpacityMask` is a drop-in nn.Module with the same parameter name/shape (`conv.weight` [1,2,7,7],
no bias), so a reference checkpoint loads unchanged.
"""
import ctypes as C  # noqa: F401

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, current_stream, ptr


class _OpacityMaskFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, opacity_bev):
        if not x.is_cuda:
            raise _lib.OcrfError("opacity_mask needs CUDA tensors: there is no CPU implementation")
        x, weight, opacity_bev = x.float().contiguous(), weight.float().contiguous(), opacity_bev.float().contiguous()
        B, Cc, H, W = x.shape
        K = weight.shape[-1]
        out = torch.empty_like(x)
        mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        stats = torch.empty((B, 2, H, W), dtype=torch.float32, device=x.device)
        check(_lib.lib().ocrf_opacity_mask_forward(current_stream(), B, Cc, H, W, K, ptr(x), ptr(weight),
                                                   ptr(opacity_bev), ptr(out), ptr(mask), ptr(stats)),
              "ocrf_opacity_mask_forward")
        ctx.save_for_backward(x, weight, mask, stats)
        ctx.mark_non_differentiable(mask)
        return out, mask

    @staticmethod
    def backward(ctx, g_out, _g_mask):
        x, weight, mask, stats = ctx.saved_tensors
        B, Cc, H, W = x.shape
        K = weight.shape[-1]
        g_out = g_out.float().contiguous()
        g_x = torch.empty_like(x)
        g_w = torch.zeros_like(weight)
        g_ob = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        scratch = torch.empty((B * 3 * H * W,), dtype=torch.float32, device=x.device)
        check(_lib.lib().ocrf_opacity_mask_backward(current_stream(), B, Cc, H, W, K, ptr(x), ptr(weight), ptr(mask),
                                                    ptr(stats), ptr(g_out), ptr(g_x), ptr(g_w), ptr(g_ob), ptr(scratch)),
              "ocrf_opacity_mask_backward")
        return g_x, g_w, g_ob


def opacity_mask(x, weight, opacity_bev):
    """x [B,C,H,W], weight [1,2,K,K], opacity_bev [B,1,H,W] -> (x * mask, mask)."""
    return _OpacityMaskFn.apply(x, weight, opacity_bev)


class OpacityMask(nn.Module):
    """`ObatinOpacityMask` followed by the gating multiply (returns the gated feature)."""

    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)  # parameters only

    def forward(self, x, opacity_bev):
        return opacity_mask(x, self.conv.weight, opacity_bev)[0]


class GeomAttentionGate(nn.Module):
    """`BEVGeomAttention` (view_transformer_ocrf.py:215-228) followed by its gating multiply (:1190,
    `geom_feat = self.geom_att(channel_feat, bev_mask_logit) * channel_feat`): the same arithmetic as the opacity
    mask with the BEV-mask logit in place of the opacity logit, so it is the same fused op.  The parameter keeps the
    reference's name (`conv1.weight` [1,2,7,7], no bias): `geom_att.*` checkpoint entries load unchanged."""

    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv1 = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)  # parameters only

    def forward(self, x, bev_prob):
        return opacity_mask(x, self.conv1.weight, bev_prob)[0]
