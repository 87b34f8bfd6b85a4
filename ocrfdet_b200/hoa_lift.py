"""Stage 5a: the height-aware opacity lift as one fused op (scope rows a13 / f-4).

Mirrors what OcRFDet does per sample at /root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:1159-1161,

    opacity_up    = F.interpolate(opacity.view(1, 13, 128, 128), size=(21, 21), mode='bilinear', align_corners=True)
    alpha_up      = F.interpolate(alpha_lidar, size=(21, 21), mode='bilinear', align_corners=True)
    opacity_alpha = F.interpolate(self.defor_cross_attention(opacity_up, alpha_up), size=(128, 128), ...) + opacity

with `defor_cross_attention = DeformableAttention2D(dim=13, dim_head=8, heads=1, dropout=0.1, downsample_factor=4,
offset_scale=4, offset_groups=None, offset_kernel_size=6)` (:639-648, class at mmdet3d/ops/cross_attention_2d.py:93-220).
`DeformableAttention2D` below keeps the reference's constructor and state_dict names (`to_offsets.0.weight`, ...,
`to_out.bias`), stored packed in ONE flat parameter (the layout the kernels read); `opacity_alpha_lift` is the three
reference lines for a whole batch, forward and backward in libocrf_raster.so.  There is no PyTorch fallback.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib

# reference parameter name -> shape, in named_parameters() order (= packed order, include/ocrf_raster.h)
ATTN_PARAMS = OrderedDict([
    ("to_offsets.0.weight", (8, 1, 6, 6)), ("to_offsets.0.bias", (8,)), ("to_offsets.2.weight", (2, 8, 1, 1)),
    ("rel_pos_bias.mlp.0.0.weight", (3, 2)), ("rel_pos_bias.mlp.0.0.bias", (3,)),
    ("rel_pos_bias.mlp.1.0.weight", (3, 3)), ("rel_pos_bias.mlp.1.0.bias", (3,)),
    ("rel_pos_bias.mlp.2.weight", (1, 3)), ("rel_pos_bias.mlp.2.bias", (1,)),
    ("to_q.weight", (8, 13, 1, 1)), ("to_k.weight", (8, 13, 1, 1)), ("to_v.weight", (8, 13, 1, 1)),
    ("to_out.weight", (13, 8, 1, 1)), ("to_out.bias", (13,))])
ATTN_TOTAL = 766


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def pack_parameters(params, layout, device=None):
    """{reference name: tensor} -> flat float32 vector in `layout` order."""
    ref = next(iter(params.values()))
    parts = []
    for name, shape in layout.items():
        t = torch.as_tensor(params[name]).to(device if device is not None else ref.device, torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
        parts.append(t.reshape(-1))
    return torch.cat(parts)


def unpack_views(flat, layout):
    out, off = OrderedDict(), 0
    for name, shape in layout.items():
        n = _numel(shape)
        out[name] = flat[off:off + n].view(shape)
        off += n
    return out


class _PackedModule(nn.Module):
    """An nn.Module whose parameters live packed in `self.packed` but load and save under the reference's names."""
    LAYOUT = None

    def reference_parameters(self, grads=False):
        return unpack_views(self.packed.grad if grads else self.packed.data, self.LAYOUT)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name, t in self.reference_parameters().items():
            destination[prefix + name] = t.clone()
        for name, b in self._buffers.items():
            if b is not None:
                destination[prefix + name.replace("__", ".")] = b if keep_vars else b.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if prefix + "packed" in state_dict:
            self.packed.data.copy_(state_dict[prefix + "packed"])
        else:
            missing = [prefix + k for k in self.LAYOUT if prefix + k not in state_dict]
            if missing:
                if strict:
                    missing_keys.extend(missing)
            else:
                try:
                    self.packed.data.copy_(pack_parameters({k: state_dict[prefix + k] for k in self.LAYOUT}, self.LAYOUT,
                                                           device=self.packed.device))
                except RuntimeError as e:
                    error_msgs.append("%s: %s" % (type(self).__name__, e))
        for name, b in self._buffers.items():
            key = prefix + name.replace("__", ".")
            if key in state_dict:
                b.copy_(state_dict[key])
            elif strict and b is not None:
                missing_keys.append(key)


class _LiftFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, opacity, alpha, packed, keep):
        L = _lib.lib()
        if not (opacity.is_cuda and alpha.is_cuda and packed.is_cuda):
            raise _lib.OcrfError("opacity_alpha_lift: tensors must live on a CUDA device (there is no CPU path)")
        opacity, alpha = opacity.float().contiguous(), alpha.float().contiguous()
        packed = packed.float().contiguous()
        B, D, H, W = opacity.shape
        ws = torch.empty(L.ocrf_hoa_lift_workspace_floats(B, D, H, W), dtype=torch.float32, device=opacity.device)
        if ws.numel() == 0:
            raise _lib.OcrfError("opacity_alpha_lift: unsupported shape %s (13 height planes, >= 36 x 36)" % (tuple(opacity.shape),))
        out = torch.empty_like(opacity)
        _lib.check(L.ocrf_hoa_lift_forward(_lib.current_stream(), B, D, H, W, _lib.ptr(opacity), _lib.ptr(alpha),
                                           _lib.ptr(packed), _lib.ptr(keep), _lib.ptr(out), _lib.ptr(ws)),
                   "ocrf_hoa_lift_forward")
        ctx.save_for_backward(opacity, alpha, packed, keep)
        return out

    @staticmethod
    def backward(ctx, g_out):
        L = _lib.lib()
        opacity, alpha, packed, keep = ctx.saved_tensors
        B, D, H, W = opacity.shape
        g_out = g_out.float().contiguous()
        ws = torch.empty(L.ocrf_hoa_lift_workspace_floats(B, D, H, W), dtype=torch.float32, device=opacity.device)
        g_opacity, g_alpha = torch.empty_like(opacity), torch.empty_like(alpha)
        g_packed = torch.zeros_like(packed)
        _lib.check(L.ocrf_hoa_lift_backward(_lib.current_stream(), B, D, H, W, _lib.ptr(opacity), _lib.ptr(alpha),
                                            _lib.ptr(packed), _lib.ptr(keep), _lib.ptr(g_out), _lib.ptr(g_opacity),
                                            _lib.ptr(g_alpha), _lib.ptr(g_packed), _lib.ptr(ws)), "ocrf_hoa_lift_backward")
        return g_opacity, g_alpha, g_packed, None


def key_grid(H, W, downsample=4, ksize=6):
    """(queries, keys) of the attention for a full-resolution H x W map."""
    ch, cw = int(H / 6), int(W / 6)
    pad = (ksize - downsample) // 2
    return ch * cw, ((ch + 2 * pad - ksize) // downsample + 1) * ((cw + 2 * pad - ksize) // downsample + 1)


class DeformableAttention2D(_PackedModule):
    """Parameter holder with the reference's constructor (cross_attention_2d.py:94-147) for OcRFDet's configuration;
    used through `opacity_alpha_lift` (the attention never runs on its own in OcRFDet)."""
    LAYOUT = ATTN_PARAMS

    def __init__(self, *, dim=13, dim_head=8, heads=1, dropout=0.0, downsample_factor=4, offset_scale=None,
                 offset_groups=None, offset_kernel_size=6, group_queries=True, group_key_values=True):
        super().__init__()
        offset_scale = downsample_factor if offset_scale is None else offset_scale
        offset_groups = heads if offset_groups is None else offset_groups
        if (dim, dim_head, heads, downsample_factor, offset_scale, offset_groups, offset_kernel_size) != (13, 8, 1, 4, 4, 1, 6):
            raise ValueError("the fused lift implements OcRFDet's configuration: dim=13, dim_head=8, heads=1, "
                             "downsample_factor=4, offset_scale=4, one offset group, offset_kernel_size=6")
        self.dropout_p = float(dropout)
        init = OrderedDict()
        ref = OrderedDict([
            ("to_offsets.0", nn.Conv2d(8, 8, 6, groups=8, stride=4, padding=1)), ("to_offsets.2", nn.Conv2d(8, 2, 1, bias=False)),
            ("rel_pos_bias.mlp.0.0", nn.Linear(2, 3)), ("rel_pos_bias.mlp.1.0", nn.Linear(3, 3)),
            ("rel_pos_bias.mlp.2", nn.Linear(3, 1)), ("to_q", nn.Conv2d(13, 8, 1, bias=False)),
            ("to_k", nn.Conv2d(13, 8, 1, bias=False)), ("to_v", nn.Conv2d(13, 8, 1, bias=False)), ("to_out", nn.Conv2d(8, 13, 1))])
        for prefix, m in ref.items():  # the reference's default initialisations
            for n, p in m.named_parameters():
                init[prefix + "." + n] = p.data
        self.packed = nn.Parameter(pack_parameters(init, ATTN_PARAMS))

    def lift(self, opacity, alpha_lidar):
        return opacity_alpha_lift(opacity, alpha_lidar, self)


def opacity_alpha_lift(opacity, alpha_lidar, attention, keep=None):
    """view_transformer_ocrf.py:1159-1161 for a batch: opacity, alpha_lidar [B,13,W,L] -> opacity_alpha [B,13,W,L].
    In training mode the attention's dropout (p = attention.dropout_p) is applied with a keep-mask drawn here (torch's
    generator), unless `keep` [B, queries, keys] (already divided by 1 - p) is given."""
    if keep is None and attention.training and attention.dropout_p > 0.0:
        nq, nk = key_grid(opacity.shape[-2], opacity.shape[-1])
        p = attention.dropout_p
        keep = (torch.rand((opacity.shape[0], nq, nk), device=opacity.device) >= p).float() / (1.0 - p)
    if keep is not None:
        keep = keep.float().contiguous()
    return _LiftFn.apply(opacity, alpha_lidar, attention.packed, keep)
