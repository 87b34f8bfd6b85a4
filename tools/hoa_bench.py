"""The height-aware opacity lift (stage 5) against a torch formulation of the same modules on the same GPU.

  lift       view_transformer_ocrf.py:1159-1161: two 128 -> 21 bilinear resizes, DeformableAttention2D
             (mmdet3d/ops/cross_attention_2d.py:93-220 in OcRFDet's configuration), 21 -> 128 resize + residual
  converter  OpacityVoxelToBEVConverter + HeightAttention (view_transformer_ocrf.py:421-518)

The torch side below is written from this repository's numpy oracle (oracle/hoa.py) with stock torch ops
(F.interpolate, F.conv2d, F.grid_sample, softmax, F.batch_norm, F.max_pool2d, F.conv_transpose2d) -- the op sequence the
reference's modules issue; the reference classes themselves are not available on the GPU box.  Both sides get the same
parameters and inputs; the script first checks that they agree (forward 1e-4, lift gradients 1e-3 of the largest
element, converter gradients by the share of elements that differ: its arg-max gates are discrete), then times forward and forward + backward with CUDA events.
Prints one JSON line; `python tools/hoa_bench.py [B]` (default: 1 and 8 samples)."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrfdet_b200 import hoa_lift as HL  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


# ---------------------------------------------------------------------------------------------------------------------
def torch_attention(p, xq, xkv, downsample=4, offset_scale=4.0, ksize=6):
    B, dim, h, w = xq.shape
    inner = p["to_q.weight"].shape[0]
    scale = inner ** -0.5
    pad = (ksize - downsample) // 2
    q = F.conv2d(xq, p["to_q.weight"])
    dw = F.conv2d(q, p["to_offsets.0.weight"], p["to_offsets.0.bias"], stride=downsample, padding=pad, groups=inner)
    off = torch.tanh(F.conv2d(F.gelu(dw), p["to_offsets.2.weight"])) * offset_scale
    hk, wk = off.shape[-2:]
    gy, gx = torch.meshgrid(torch.arange(hk, device=xq.device, dtype=xq.dtype),
                            torch.arange(wk, device=xq.device, dtype=xq.dtype), indexing="ij")
    vgrid = torch.stack([gx, gy])[None] + off
    vn = torch.stack([2.0 * vgrid[:, 0] / max(hk - 1, 1) - 1.0, 2.0 * vgrid[:, 1] / max(wk - 1, 1) - 1.0], -1)
    kvf = F.grid_sample(xkv, vn, mode="bilinear", padding_mode="zeros", align_corners=False)
    k = F.conv2d(kvf, p["to_k.weight"]).reshape(B, inner, hk * wk)
    v = F.conv2d(kvf, p["to_v.weight"]).reshape(B, inner, hk * wk)
    qs = (q * scale).reshape(B, inner, h * w)
    sim = torch.einsum("bdi,bdj->bij", qs, k)
    qy, qx = torch.meshgrid(torch.arange(h, device=xq.device, dtype=xq.dtype),
                            torch.arange(w, device=xq.device, dtype=xq.dtype), indexing="ij")
    gq = torch.stack([2.0 * qx / max(h - 1, 1) - 1.0, 2.0 * qy / max(w - 1, 1) - 1.0], -1).reshape(h * w, 2)
    pos = gq[None, :, None, :] - vn.reshape(B, hk * wk, 2)[:, None, :, :]
    bb = torch.sign(pos) * torch.log(pos.abs() + 1.0)
    h1 = F.relu(F.linear(bb, p["rel_pos_bias.mlp.0.0.weight"], p["rel_pos_bias.mlp.0.0.bias"]))
    h2 = F.relu(F.linear(h1, p["rel_pos_bias.mlp.1.0.weight"], p["rel_pos_bias.mlp.1.0.bias"]))
    bias = F.linear(h2, p["rel_pos_bias.mlp.2.weight"], p["rel_pos_bias.mlp.2.bias"])[..., 0]
    attn = (sim + bias).softmax(-1)
    out = torch.einsum("bij,bdj->bdi", attn, v).reshape(B, inner, h, w)
    return F.conv2d(out, p["to_out.weight"], p["to_out.bias"])


def torch_lift(p, opacity, alpha):
    S = opacity.shape[-1]
    c = int(S / 6)
    up = F.interpolate(opacity, size=(c, c), mode="bilinear", align_corners=True)
    al = F.interpolate(alpha, size=(c, c), mode="bilinear", align_corners=True)
    return F.interpolate(torch_attention(p, up, al), size=(S, S), mode="bilinear", align_corners=True) + opacity


def torch_converter(p, x, position, train=True):
    def block(name, t):
        cin = t.shape[1]
        t = F.conv2d(t, p[name + ".0.weight"], p[name + ".0.bias"], padding=1, groups=cin)
        t = F.conv2d(t, p[name + ".1.weight"], p[name + ".1.bias"])
        t = F.batch_norm(t, None, None, p[name + ".2.weight"], p[name + ".2.bias"], training=True)
        return F.relu(t)

    def gate(name, e):
        cs = e.shape[1] // 4
        mx = F.adaptive_max_pool2d(e, 1)
        parts = []
        for s in range(4):
            m = mx[:, s * cs:(s + 1) * cs]
            hid = F.relu(F.conv2d(m, p["%s.conv%d.0.weight" % (name, s + 1)]))
            parts.append(torch.sigmoid(F.conv2d(hid, p["%s.conv%d.2.weight" % (name, s + 1)])))
        return e * torch.cat(parts, 1)

    enc1 = gate("ca1", block("encoder1", x) + position)
    enc2 = gate("ca2", block("encoder2", F.max_pool2d(enc1, 2)))
    bott = gate("ca_bottleneck", block("bottleneck", F.max_pool2d(enc2, 2)))
    u2 = F.conv_transpose2d(bott, p["upconv2.weight"], p["upconv2.bias"], stride=2)
    dec2 = gate("ca_dec2", block("decoder2", torch.cat([u2, enc2], 1)))
    u1 = F.conv_transpose2d(dec2, p["upconv1.weight"], p["upconv1.bias"], stride=2)
    dec1 = gate("ca_dec1", block("decoder1", torch.cat([u1, enc1], 1)))
    return F.conv2d(dec1, p["output_conv.weight"], p["output_conv.bias"])


# ---------------------------------------------------------------------------------------------------------------------
def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def run(B, S=128):
    torch.manual_seed(3)
    dev = "cuda"
    attn = HL.DeformableAttention2D(dim=13, dim_head=8, heads=1, dropout=0.0, downsample_factor=4, offset_scale=4,
                                    offset_groups=None, offset_kernel_size=6).to(dev).eval()
    conv = HL.OpacityVoxelToBEVConverter(input_channel=13).to(dev).train()
    with torch.no_grad():
        attn.packed.add_(0.05 * torch.randn_like(attn.packed))
    pa = {k: v.detach().clone().requires_grad_(True) for k, v in HL.unpack_views(attn.packed, attn.LAYOUT).items()}
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in HL.unpack_views(conv.packed, conv.LAYOUT).items()}
    opacity = torch.sigmoid(1.5 * torch.randn(B, 13, S, S, device=dev)).requires_grad_(True)
    alpha = ((1 - torch.exp(-torch.randn(B, 13, S, S, device=dev).abs())) * (torch.rand(B, 13, S, S, device=dev) < 0.7)).requires_grad_(True)
    position = (0.5 * torch.randn(1, 4, S, S, device=dev)).requires_grad_(True)
    g_lift, g_bev = torch.randn(B, 13, S, S, device=dev), torch.randn(B, 1, S, S, device=dev)

    # ---- agreement
    ours = HL.opacity_alpha_lift(opacity, alpha, attn)
    ref = torch_lift(pa, opacity, alpha)
    chk = {"lift_fwd": rel(ours, ref)}
    go = torch.autograd.grad(ours, [opacity, alpha], g_lift)
    gr = torch.autograd.grad(ref, [opacity, alpha], g_lift)
    chk["lift_dopacity"], chk["lift_dalpha"] = rel(go[0], gr[0]), rel(go[1], gr[1])
    x = ours.detach().requires_grad_(True)
    oc = conv(x, position)
    rc = torch_converter(pc, x, position)
    chk["conv_fwd"] = rel(oc, rc)
    # The converter's gates route their gradient through the ARG-MAX of a whole plane: two fp32 implementations whose
    # activations differ in the last bits may pick different elements among near-equal maxima (torch in fp32 and in
    # fp64 do, too), which moves the gradient of a patch.  So the gradient is judged by the share of elements that
    # differ, not by the largest difference.
    gx_o, gx_r = torch.autograd.grad(oc, x, g_bev)[0], torch.autograd.grad(rc, x, g_bev)[0]
    chk["conv_dx_share_beyond_1e-4"] = float(((gx_o - gx_r).abs() > 1e-4 * gx_r.abs().max()).float().mean())
    assert chk["lift_fwd"] <= 1e-4 and chk["conv_fwd"] <= 1e-4, chk
    assert max(chk["lift_dopacity"], chk["lift_dalpha"]) <= 1e-3 and chk["conv_dx_share_beyond_1e-4"] <= 0.02, chk

    # ---- timing
    def ours_lift_f():
        with torch.no_grad():
            HL.opacity_alpha_lift(opacity, alpha, attn)

    def ref_lift_f():
        with torch.no_grad():
            torch_lift(pa, opacity, alpha)

    def ours_lift_fb():
        torch.autograd.grad(HL.opacity_alpha_lift(opacity, alpha, attn), [opacity, alpha, attn.packed], g_lift)

    def ref_lift_fb():
        torch.autograd.grad(torch_lift(pa, opacity, alpha), [opacity, alpha] + list(pa.values()), g_lift, allow_unused=True)

    def ours_conv_f():
        with torch.no_grad():
            conv(x, position)

    def ref_conv_f():
        with torch.no_grad():
            torch_converter(pc, x, position)

    def ours_conv_fb():
        torch.autograd.grad(conv(x, position), [x, position, conv.packed], g_bev)

    def ref_conv_fb():
        torch.autograd.grad(torch_converter(pc, x, position), [x, position] + list(pc.values()), g_bev)

    t = {k: timed(f) for k, f in (("lift_fwd_ours", ours_lift_f), ("lift_fwd_torch", ref_lift_f),
                                  ("lift_fwdbwd_ours", ours_lift_fb), ("lift_fwdbwd_torch", ref_lift_fb),
                                  ("conv_fwd_ours", ours_conv_f), ("conv_fwd_torch", ref_conv_f),
                                  ("conv_fwdbwd_ours", ours_conv_fb), ("conv_fwdbwd_torch", ref_conv_fb))}
    return {"samples": B, "ms": {k: round(v, 4) for k, v in t.items()}, "agreement": {k: float("%.2e" % v) for k, v in chk.items()},
            "speedup": {"lift_fwdbwd": round(t["lift_fwdbwd_torch"] / t["lift_fwdbwd_ours"], 2),
                        "conv_fwdbwd": round(t["conv_fwdbwd_torch"] / t["conv_fwdbwd_ours"], 2)}}


if __name__ == "__main__":
    Bs = [int(a) for a in sys.argv[1:]] or [1, 8]
    print(json.dumps({"what": "HOA lift + converter, ours (fused CUDA) vs a torch formulation on the same GPU; ms per call",
                      "runs": [run(B) for B in Bs]}))
