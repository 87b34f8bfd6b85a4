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


# ---------------------------------------------------------------------------------------------------------------------
# Stage 5b: OpacityVoxelToBEVConverter (+ HeightAttention), view_transformer_ocrf.py:421-518
# ---------------------------------------------------------------------------------------------------------------------
BLOCKS = (("encoder1", 13, 4), ("encoder2", 4, 8), ("bottleneck", 8, 16), ("decoder2", 16, 8), ("decoder1", 8, 4))
_GATE_OF = {"encoder1": "ca1", "encoder2": "ca2", "bottleneck": "ca_bottleneck", "decoder2": "ca_dec2", "decoder1": "ca_dec1"}
_UPCONV_BEFORE = {"decoder2": ("upconv2", 16, 8), "decoder1": ("upconv1", 8, 4)}


def _converter_modules():
    """The reference module tree, layer by layer, in named_parameters() order: [(prefix, nn.Module)]."""
    mods = []
    for name, cin, cout in BLOCKS:
        if name in _UPCONV_BEFORE:
            up, a, b = _UPCONV_BEFORE[name]
            mods.append((up, nn.ConvTranspose2d(a, b, kernel_size=2, stride=2)))
        mods.append((name + ".0", nn.Conv2d(cin, cin, kernel_size=3, padding=1, groups=cin)))
        mods.append((name + ".1", nn.Conv2d(cin, cout, kernel_size=1)))
        mods.append((name + ".2", nn.BatchNorm2d(cout)))
        cs = cout // 4
        for s in range(1, 5):
            mods.append(("%s.conv%d.0" % (_GATE_OF[name], s), nn.Conv2d(cs, cs, 1, bias=False)))
            mods.append(("%s.conv%d.2" % (_GATE_OF[name], s), nn.Conv2d(cs, cs, 1, bias=False)))
    mods.append(("output_conv", nn.Conv2d(4, 1, kernel_size=1)))
    return mods


def _converter_layout_and_init():
    layout, init = OrderedDict(), OrderedDict()
    for prefix, m in _converter_modules():
        for n, p in m.named_parameters():
            layout[prefix + "." + n] = tuple(p.shape)
            init[prefix + "." + n] = p.data
    return layout, init


CONVERTER_PARAMS, _ = _converter_layout_and_init()
CONVERTER_TOTAL = 1847
BN_MOMENTUM = 0.1


class _ConverterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, position, packed, run_mean, run_var, train):
        L = _lib.lib()
        if not (x.is_cuda and position.is_cuda and packed.is_cuda):
            raise _lib.OcrfError("OpacityVoxelToBEVConverter: tensors must live on a CUDA device (there is no CPU path)")
        x, position, packed = x.float().contiguous(), position.float().contiguous(), packed.float().contiguous()
        B, C, S, S2 = x.shape
        if C != 13 or S != S2 or S % 4 or position.shape[1:] != (4, S, S) or position.shape[0] not in (1, B):
            raise _lib.OcrfError("OpacityVoxelToBEVConverter: x must be [B,13,S,S] (S % 4 == 0), position [1 or B,4,S,S]")
        pos_batched = int(position.shape[0] == B and B > 1)
        ws = torch.empty(L.ocrf_hoa_converter_workspace_floats(B, S), dtype=torch.float32, device=x.device)
        out = torch.empty((B, 1, S, S), dtype=torch.float32, device=x.device)
        stats = torch.empty((5, 16, 2), dtype=torch.float32, device=x.device) if train else None
        _lib.check(L.ocrf_hoa_converter_forward(_lib.current_stream(), B, S, int(train), _lib.ptr(x), _lib.ptr(position),
                                                pos_batched, _lib.ptr(packed), _lib.ptr(run_mean), _lib.ptr(run_var),
                                                _lib.ptr(out), _lib.ptr(stats), _lib.ptr(ws)), "ocrf_hoa_converter_forward")
        ctx.save_for_backward(x, position, packed, run_mean, run_var, ws)
        ctx.train, ctx.pos_batched = bool(train), pos_batched
        ctx.mark_non_differentiable(*([stats] if stats is not None else []))
        return (out, stats) if stats is not None else (out, None)

    @staticmethod
    def backward(ctx, g_out, _g_stats):
        L = _lib.lib()
        x, position, packed, run_mean, run_var, ws = ctx.saved_tensors
        B, _, S, _ = x.shape
        g_out = g_out.float().contiguous()
        g_x, g_pos = torch.empty_like(x), torch.empty_like(position)
        g_packed = torch.zeros_like(packed)
        _lib.check(L.ocrf_hoa_converter_backward(_lib.current_stream(), B, S, int(ctx.train), _lib.ptr(x),
                                                 _lib.ptr(position), ctx.pos_batched, _lib.ptr(packed), _lib.ptr(run_mean),
                                                 _lib.ptr(run_var), _lib.ptr(g_out), _lib.ptr(g_x), _lib.ptr(g_pos),
                                                 _lib.ptr(g_packed), _lib.ptr(ws)), "ocrf_hoa_converter_backward")
        return g_x, g_pos, g_packed, None, None, None


class OpacityVoxelToBEVConverter(_PackedModule):
    """Drop-in for the reference class (view_transformer_ocrf.py:463-518): `forward(x, position)` with the lifted
    opacity [B,13,S,S] and the learned BEV position encoding [1 or B,4,S,S] -> BEV opacity logit [B,1,S,S].
    Same state_dict keys (parameters and batch-norm buffers), same default initialisation; batch norm follows
    `self.training` (batch statistics + running-statistics update, or running statistics)."""
    LAYOUT = CONVERTER_PARAMS

    def __init__(self, input_channel=13):
        super().__init__()
        if input_channel != 13:
            raise ValueError("the fused converter implements OcRFDet's 13 height planes")
        _, init = _converter_layout_and_init()
        self.packed = nn.Parameter(pack_parameters(init, CONVERTER_PARAMS))
        # batch-norm buffers of the five blocks, packed [block, channel] as the kernels read them (one module call is
        # one kernel chain plus a handful of vector ops, not fifty tiny launches); state_dict keys are the reference's
        self.register_buffer("bn_mean", torch.zeros(5, 16), persistent=False)
        self.register_buffer("bn_var", torch.ones(5, 16), persistent=False)
        self.register_buffer("bn_batches", torch.zeros(5, dtype=torch.long), persistent=False)
        mask = torch.zeros(5, 16)
        for i, (_name, _cin, cout) in enumerate(BLOCKS):
            mask[i, :cout] = 1.0
        self.register_buffer("bn_mask", mask, persistent=False)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name, t in self.reference_parameters().items():
            destination[prefix + name] = t.clone()
        for i, (name, _cin, cout) in enumerate(BLOCKS):
            destination[prefix + name + ".2.running_mean"] = self.bn_mean[i, :cout].clone()
            destination[prefix + name + ".2.running_var"] = self.bn_var[i, :cout].clone()
            destination[prefix + name + ".2.num_batches_tracked"] = self.bn_batches[i].clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        saved, self._buffers = self._buffers, type(self._buffers)()  # the base class handles the parameters only
        try:
            super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        finally:
            self._buffers = saved
        for i, (name, _cin, cout) in enumerate(BLOCKS):
            for key, dst in ((".2.running_mean", self.bn_mean[i, :cout]), (".2.running_var", self.bn_var[i, :cout]),
                             (".2.num_batches_tracked", self.bn_batches[i])):
                k = prefix + name + key
                if k in state_dict:
                    dst.copy_(state_dict[k])
                elif strict:
                    missing_keys.append(k)

    def forward(self, x, position):
        # (training mode updates the buffers in place below; the autograd node keeps what the forward saw)
        mean, var = (self.bn_mean.clone(), self.bn_var.clone()) if self.training else (self.bn_mean, self.bn_var)
        out, stats = _ConverterFn.apply(x, position, self.packed, mean, var, self.training)
        if self.training:
            with torch.no_grad():  # nn.BatchNorm2d: momentum 0.1, running_var from the unbiased batch variance
                B, _, S, _ = x.shape
                n = torch.tensor([float(B * (S >> sh) * (S >> sh)) for sh in (0, 1, 2, 1, 0)], device=stats.device)[:, None]
                m = stats[:, :, 0] / n
                v = (stats[:, :, 1] / n - m * m).clamp_min_(0.0) * (n / (n - 1.0).clamp_min(1.0))
                self.bn_mean.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * m * self.bn_mask)
                self.bn_var.lerp_(torch.where(self.bn_mask > 0, v, self.bn_var), BN_MOMENTUM)
                self.bn_batches.add_(1)
        return out
