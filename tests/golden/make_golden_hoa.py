"""Golden vectors for the height-aware opacity (HOA) lift, produced by the REFERENCE's own module code.

Run in the build container (needs /root/reference; CPU only):   python tests/golden/make_golden_hoa.py

  * `DeformableAttention2D` is imported from /root/reference/mmdet3d/ops/cross_attention_2d.py where it lies (the file
    needs only torch + einops) and wrapped exactly as view_transformer_ocrf.py:1159-1161 does: bilinear 128 -> 21
    (align_corners=True) of the per-voxel opacity and of alpha_lidar, cross attention, bilinear back + residual.
  * `HeightAttention` and `OpacityVoxelToBEVConverter` are lifted out of view_transformer_ocrf.py:421-518 with `ast`
    (the file itself needs mmcv) and executed unmodified, in training mode (batch-norm batch statistics) and in eval
    mode (running statistics).
Inputs are regenerated from seeds by the tests (`hoa_inputs` below is imported by them); the files hold the parameters,
every small tensor in full, and the full-resolution outputs / input gradients on a stride-3 grid plus their sums.
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

VT = "/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py"
CA = "/root/reference/mmdet3d/ops/cross_attention_2d.py"
HERE = os.path.dirname(os.path.abspath(__file__))
STRIDE = 3


def hoa_inputs(seed, B, Hh=13, S=128):
    """Per-voxel opacity (sigmoid of the A_MLP) and alpha_lidar (1 - exp(-sigma), zero where no camera sees the voxel)
    of B samples, the BEV position encoding [1,4,S,S], and upstream gradients; all float32, from numpy's PCG64."""
    rng = np.random.default_rng(seed)
    opacity = (1.0 / (1.0 + np.exp(-rng.normal(0.0, 1.5, size=(B, Hh, S, S))))).astype(np.float32)
    alpha = (1.0 - np.exp(-np.abs(rng.normal(0.0, 1.0, size=(B, Hh, S, S))))).astype(np.float32)
    alpha *= (rng.random(size=(B, Hh, S, S)) < 0.7)
    position = rng.normal(0.0, 0.5, size=(1, 4, S, S)).astype(np.float32)
    g_lift = rng.normal(size=(B, Hh, S, S)).astype(np.float32)
    g_bev = rng.normal(size=(B, 1, S, S)).astype(np.float32)
    return dict(opacity=opacity, alpha=alpha, position=position, g_lift=g_lift, g_bev=g_bev)


def reference_attention():
    spec = importlib.util.spec_from_file_location("_ref_cross_attention_2d", CA)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DeformableAttention2D


def reference_converter():
    tree = ast.parse(open(VT).read())
    wanted = {"HeightAttention", "OpacityVoxelToBEVConverter"}
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in wanted]
    assert len(body) == 2
    ns = {"torch": torch, "nn": torch.nn, "F": F}
    exec(compile(ast.Module(body=body, type_ignores=[]), VT, "exec"), ns)
    return ns["OpacityVoxelToBEVConverter"]


def sub(t):
    return t.detach().numpy()[..., ::STRIDE, ::STRIDE].copy()


def lift_case(name, seed, B, S):
    torch.manual_seed(seed)
    # the constructor call of view_transformer_ocrf.py:639-648
    attn = reference_attention()(dim=13, dim_head=8, heads=1, dropout=0.1, downsample_factor=4, offset_scale=4,
                                 offset_groups=None, offset_kernel_size=6).eval()  # eval: dropout off (deterministic)
    with torch.no_grad():  # default init leaves the offsets and the position bias almost flat: spread them
        for n_, p in attn.named_parameters():
            if "to_offsets" in n_ or "rel_pos_bias" in n_:
                p.mul_(3.0)
    io = hoa_inputs(seed, B, S=S)
    opacity = torch.from_numpy(io["opacity"]).requires_grad_(True)
    alpha = torch.from_numpy(io["alpha"]).requires_grad_(True)
    Width = Length = S
    # view_transformer_ocrf.py:1159-1161 (per sample there; the module is batch-agnostic)
    opacity_up = F.interpolate(opacity, size=(int(Width / 6), int(Length / 6)), mode="bilinear", align_corners=True)
    alpha_up = F.interpolate(alpha, size=(int(Width / 6), int(Length / 6)), mode="bilinear", align_corners=True)
    att, vgrid = attn(opacity_up, alpha_up, return_vgrid=True)
    opacity_alpha = F.interpolate(att, size=(Width, Length), mode="bilinear", align_corners=True) + opacity
    (opacity_alpha * torch.from_numpy(io["g_lift"])).sum().backward()
    rec = dict(seed=seed, B=B, S=S, opacity_up=opacity_up.detach().numpy(), alpha_up=alpha_up.detach().numpy(),
               att=att.detach().numpy(), vgrid=vgrid.detach().numpy(), opacity_alpha_sub=sub(opacity_alpha),
               opacity_alpha_sum=opacity_alpha.detach().double().sum().numpy(), g_opacity_sub=sub(opacity.grad),
               g_alpha_sub=sub(alpha.grad), g_opacity_sum=opacity.grad.double().sum().numpy(),
               g_alpha_sum=alpha.grad.double().sum().numpy(),
               g_alpha_abs_sum=alpha.grad.double().abs().sum().numpy())
    for n_, p in attn.named_parameters():
        rec["p." + n_] = p.detach().numpy()
        rec["g." + n_] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "att", tuple(att.shape), "params", sum(p.numel() for p in attn.parameters()))


def converter_case(name, seed, B, S, train):
    torch.manual_seed(seed)
    conv = reference_converter()(input_channel=13)
    with torch.no_grad():
        for m in conv.modules():
            if isinstance(m, torch.nn.BatchNorm2d):  # non-trivial affine and running statistics
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    conv.train(train)
    io = hoa_inputs(seed, B, S=S)
    # the converter's input is the lifted opacity (opacity + attention): same range as opacity + noise
    x = torch.from_numpy(io["opacity"] + 0.25 * io["alpha"]).requires_grad_(True)
    pos = torch.from_numpy(io["position"]).requires_grad_(True)
    running = {n_: b.clone() for n_, b in conv.named_buffers()}
    out = conv(x, pos)  # view_transformer_ocrf.py:1196
    (out * torch.from_numpy(io["g_bev"])).sum().backward()
    rec = dict(seed=seed, B=B, S=S, train=int(train), out=out.detach().numpy(), g_x_sub=sub(x.grad),
               g_x_sum=x.grad.double().sum().numpy(), g_x_abs_sum=x.grad.double().abs().sum().numpy(),
               g_pos=pos.grad.numpy())
    for n_, p in conv.named_parameters():
        rec["p." + n_] = p.detach().numpy()
        rec["g." + n_] = p.grad.numpy()
    for n_, b in conv.named_buffers():
        rec["b0." + n_] = running[n_].numpy()      # before the forward
        rec["b1." + n_] = b.detach().numpy()       # after (training mode updates the running statistics)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "out", tuple(out.shape), "params", sum(p.numel() for p in conv.parameters()))


def main():
    lift_case("hoa_lift_b2", seed=501, B=2, S=128)
    lift_case("hoa_lift_b1_s96", seed=502, B=1, S=96)
    converter_case("hoa_converter_train_b2", seed=511, B=2, S=128, train=True)
    converter_case("hoa_converter_eval_b1", seed=512, B=1, S=64, train=False)


if __name__ == "__main__":
    sys.exit(main())
