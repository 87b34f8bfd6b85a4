"""BEV pooling v2 (scope row f-3): the reference's `bev_pool_v2` with the same signature.

Mirrors /root/reference/mmdet3d/ops/bev_pool_v2/bev_pool.py:10-88 (`QuickCumsumCuda` + `bev_pool_v2`): same
arguments, same output `[B, C, Z, Y, X]`, gradients for `depth` and `feat`.  The backward takes the forward's rank
arrays as they are: the per-step argsort / re-indexing / interval construction of the reference's Python
(bev_pool.py:47-60) happens inside `ocrf_bev_pool_backward` (a 2-3 pass radix sort + binary search).

There is no PyTorch fallback: both directions call libocrf_raster.so.
"""
import torch

from . import _lib

__all__ = ["bev_pool_v2", "QuickCumsumCuda"]


class QuickCumsumCuda(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths):
        L = _lib.lib()
        if not (depth.is_cuda and feat.is_cuda):
            raise _lib.OcrfError("bev_pool_v2: tensors must live on a CUDA device (there is no CPU path)")
        ranks_bev = ranks_bev.int().contiguous()
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        ranks_depth = ranks_depth.contiguous().int()
        ranks_feat = ranks_feat.contiguous().int()
        interval_lengths = interval_lengths.contiguous().int()
        interval_starts = interval_starts.contiguous().int()
        c = feat.shape[-1]
        if tuple(bev_feat_shape)[-1] != c:
            raise ValueError("bev_feat_shape[-1] must equal the feature channels")
        out = feat.new_zeros(bev_feat_shape)
        _lib.check(L.ocrf_bev_pool_forward(_lib.current_stream(), c, interval_starts.numel(), _lib.ptr(depth),
                                           _lib.ptr(feat), _lib.ptr(ranks_depth), _lib.ptr(ranks_feat),
                                           _lib.ptr(ranks_bev), _lib.ptr(interval_starts), _lib.ptr(interval_lengths),
                                           _lib.ptr(out)), "ocrf_bev_pool_forward")
        ctx.save_for_backward(ranks_bev, depth, feat, ranks_feat, ranks_depth)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        L = _lib.lib()
        ranks_bev, depth, feat, ranks_feat, ranks_depth = ctx.saved_tensors
        c = feat.shape[-1]
        n_points = ranks_bev.numel()
        n_feat = feat.numel() // c
        out_grad = out_grad.contiguous().float()
        depth_grad = torch.zeros_like(depth)
        feat_grad = torch.empty_like(feat)
        ws = torch.empty(L.ocrf_bev_pool_backward_workspace_bytes(n_points), dtype=torch.uint8, device=feat.device)
        _lib.check(L.ocrf_bev_pool_backward(_lib.current_stream(), c, n_points, n_feat, _lib.ptr(out_grad),
                                            _lib.ptr(depth), _lib.ptr(feat), _lib.ptr(ranks_depth), _lib.ptr(ranks_feat),
                                            _lib.ptr(ranks_bev), _lib.ptr(depth_grad), _lib.ptr(feat_grad), _lib.ptr(ws)),
                   "ocrf_bev_pool_backward")
        return depth_grad, feat_grad, None, None, None, None, None, None


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths):
    """depth [B,N,D,H,W], feat [B,N,H,W,C], int rank / interval tensors as produced by
    `voxel_pooling_prepare_v2`, bev_feat_shape (B,Z,Y,X,C) -> [B,C,Z,Y,X] (bev_pool.py:82-88)."""
    x = QuickCumsumCuda.apply(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                              interval_lengths)
    return x.permute(0, 4, 1, 2, 3).contiguous()
