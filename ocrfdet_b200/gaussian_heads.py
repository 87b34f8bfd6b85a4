"""OcRF Gaussian construction: the S/R/A/C heads as one fused op (scope row a12, "next" f-1).

Mirrors the four modules the reference builds at
/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:611-622 (classes at :272-320) and evaluates at
:1130-1133.  ``GaussianHeads`` stores the 16 reference parameter tensors PACKED in one flat ``nn.Parameter`` (the
layout the kernel reads, see include/ocrf_raster.h), so a step costs one forward and one backward launch and no
packing kernels; ``state_dict()`` / ``load_state_dict()`` speak the reference's names (``S_MLP.fc1.weight`` ...),
so the corresponding slice of an OcRFDet checkpoint loads and saves unchanged.  ``forward`` returns the four tensors
in the order the reference computes them (opacity, scaling, rotation, color).

There is no PyTorch fallback: the op calls ``ocrf_gaussian_heads_forward/backward`` in libocrf_raster.so.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib

HIDDEN = 4
HEADS = (("S_MLP", 3), ("R_MLP", 4), ("A_MLP", 1), ("C_MLP", 3))  # packed order; see include/ocrf_raster.h
MAX_FEAT = 125


def packed_sizes(feat_dim):
    """(offset, shape) of w1t, b1, w2, b2 inside the flat parameter vector."""
    k = feat_dim + 3
    out, off = OrderedDict(), 0
    for name, shape in (("w1t", (k, 16)), ("b1", (16,)), ("w2", (11, 4)), ("b2", (11,))):
        out[name] = (off, shape)
        off += int(math.prod(shape))
        off = (off + 3) // 4 * 4   # keep every block 16-byte aligned
    return out, off


def _views(flat, feat_dim):
    lay, _ = packed_sizes(feat_dim)
    return {name: flat[off:off + int(math.prod(shape))].view(shape) for name, (off, shape) in lay.items()}


def _head_slices(feat_dim):
    """reference parameter name -> function extracting it (as a view) from the packed views."""
    out, row = OrderedDict(), 0
    for i, (name, n_out) in enumerate(HEADS):
        cols = slice(4 * i, 4 * i + 4)
        k = feat_dim + 3 if name == "C_MLP" else feat_dim
        out[name + ".fc1.weight"] = (lambda v, cols=cols, k=k: v["w1t"][:k, cols].t())
        out[name + ".fc1.bias"] = (lambda v, cols=cols: v["b1"][cols])
        out[name + ".fc2.weight"] = (lambda v, r=row, n=n_out: v["w2"][r:r + n])
        out[name + ".fc2.bias"] = (lambda v, r=row, n=n_out: v["b2"][r:r + n])
        row += n_out
    return out


class _GaussianHeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rgb, packed):
        L = _lib.lib()
        if not (feat.is_cuda and rgb.is_cuda and packed.is_cuda):
            raise _lib.OcrfError("gaussian_heads: tensors must live on a CUDA device (there is no CPU path)")
        n, Fd = feat.shape
        feat, rgb, packed = feat.contiguous().float(), rgb.contiguous().float(), packed.contiguous().float()
        v = _views(packed, Fd)
        # one allocation: hidden [n,16] | rotations [n,4] | scales [n,3] | colors [n,3] | opacity [n,1]
        buf = torch.empty(n * 27, device=feat.device, dtype=torch.float32)
        hidden = buf[:16 * n].view(n, 16)
        rotations = buf[16 * n:20 * n].view(n, 4)
        scales = buf[20 * n:23 * n].view(n, 3)
        colors = buf[23 * n:26 * n].view(n, 3)
        opacity = buf[26 * n:].view(n, 1)
        _lib.check(L.ocrf_gaussian_heads_forward(_lib.current_stream(), n, Fd, _lib.ptr(feat), _lib.ptr(rgb),
                                                 _lib.ptr(v["w1t"]), _lib.ptr(v["b1"]), _lib.ptr(v["w2"]),
                                                 _lib.ptr(v["b2"]), _lib.ptr(opacity), _lib.ptr(scales),
                                                 _lib.ptr(rotations), _lib.ptr(colors), _lib.ptr(hidden)),
                   "ocrf_gaussian_heads_forward")
        ctx.save_for_backward(feat, rgb, packed, hidden)
        ctx.mark_non_differentiable(hidden)
        return opacity, scales, rotations, colors, hidden

    @staticmethod
    def backward(ctx, g_opacity, g_scales, g_rotations, g_colors, _g_hidden):
        L = _lib.lib()
        feat, rgb, packed, hidden = ctx.saved_tensors
        n, Fd = feat.shape
        v = _views(packed, Fd)
        g_opacity, g_scales = g_opacity.contiguous().float(), g_scales.contiguous().float()
        g_rotations, g_colors = g_rotations.contiguous().float(), g_colors.contiguous().float()
        g_feat = torch.empty_like(feat)
        g_packed = torch.zeros_like(packed)
        gv = _views(g_packed, Fd)
        ws = torch.empty(L.ocrf_gaussian_heads_backward_workspace_bytes(n), dtype=torch.uint8, device=feat.device)
        _lib.check(L.ocrf_gaussian_heads_backward(_lib.current_stream(), n, Fd, _lib.ptr(feat), _lib.ptr(rgb),
                                                  _lib.ptr(v["w1t"]), _lib.ptr(v["w2"]), _lib.ptr(v["b2"]),
                                                  _lib.ptr(hidden), _lib.ptr(g_opacity), _lib.ptr(g_scales),
                                                  _lib.ptr(g_rotations), _lib.ptr(g_colors), _lib.ptr(g_feat),
                                                  _lib.ptr(gv["w1t"]), _lib.ptr(gv["b1"]), _lib.ptr(gv["w2"]),
                                                  _lib.ptr(gv["b2"]), _lib.ptr(ws)), "ocrf_gaussian_heads_backward")
        return g_feat, None, g_packed


def gaussian_heads(feat, rgb, packed):
    """feat [..., F], rgb [..., 3], packed = flat parameter vector of ``packed_sizes(F)``
    -> opacity [...,1], scaling [...,3], rotation [...,4], color [...,3]."""
    lead = feat.shape[:-1]
    Fd = feat.shape[-1]
    if Fd > MAX_FEAT:
        raise ValueError("gaussian_heads supports at most %d feature channels, got %d" % (MAX_FEAT, Fd))
    if rgb.shape[:-1] != lead or rgb.shape[-1] != 3:
        raise ValueError("rgb must be [..., 3] with the leading shape of feat")
    if packed.dim() != 1 or packed.numel() != packed_sizes(Fd)[1]:
        raise ValueError("packed parameter vector does not match %d feature channels" % Fd)
    out = _GaussianHeadsFn.apply(feat.reshape(-1, Fd), rgb.reshape(-1, 3), packed)[:4]
    return tuple(o.reshape(*lead, o.shape[-1]) for o in out)


def pack_reference_parameters(params, feat_dim, device=None):
    """params: {"S_MLP.fc1.weight": tensor, ...} (reference names) -> flat packed vector (rgb rows of S/R/A zero)."""
    _, total = packed_sizes(feat_dim)
    ref = next(iter(params.values()))
    flat = torch.zeros(total, dtype=torch.float32, device=device if device is not None else ref.device)
    v = _views(flat, feat_dim)
    for name, get in _head_slices(feat_dim).items():
        get(v).copy_(params[name].to(flat.device, torch.float32))
    return flat


class GaussianHeads(nn.Module):
    """Drop-in for the (A_MLP, S_MLP, R_MLP, C_MLP) quartet: same initialisation (nn.Linear defaults), same
    checkpoint keys; one flat parameter ``packed`` is what the optimiser sees."""

    def __init__(self, input_dim=80):
        super().__init__()
        if input_dim > MAX_FEAT:
            raise ValueError("at most %d feature channels" % MAX_FEAT)
        self.input_dim = input_dim
        init = {}
        for name, n_out in HEADS:
            fc1 = nn.Linear(input_dim + 3 if name == "C_MLP" else input_dim, HIDDEN)
            fc2 = nn.Linear(HIDDEN, n_out)
            init.update({name + ".fc1.weight": fc1.weight.data, name + ".fc1.bias": fc1.bias.data,
                         name + ".fc2.weight": fc2.weight.data, name + ".fc2.bias": fc2.bias.data})
        self.packed = nn.Parameter(pack_reference_parameters(init, input_dim))

    # ---- the reference's view of the parameters -------------------------------------------------------------
    def reference_parameters(self, grads=False):
        """OrderedDict of reference-named VIEWS into the packed parameter (or into its .grad)."""
        src = self.packed.grad if grads else self.packed.data
        v = _views(src, self.input_dim)
        return OrderedDict((name, get(v)) for name, get in _head_slices(self.input_dim).items())

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name, t in self.reference_parameters().items():
            destination[prefix + name] = t.clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if prefix + "packed" in state_dict:
            self.packed.data.copy_(state_dict[prefix + "packed"])
            return
        names = list(_head_slices(self.input_dim))
        missing = [prefix + k for k in names if prefix + k not in state_dict]
        if missing:
            if strict:
                missing_keys.extend(missing)
            return
        try:
            flat = pack_reference_parameters({k: state_dict[prefix + k] for k in names}, self.input_dim,
                                             device=self.packed.device)
        except RuntimeError as e:  # shape mismatch
            error_msgs.append("GaussianHeads: %s" % e)
            return
        self.packed.data.copy_(flat)

    def forward(self, voxel_feat, colored_voxels):
        opacity, scaling, rotation, color = gaussian_heads(voxel_feat, colored_voxels, self.packed)
        return opacity, scaling, rotation, color
