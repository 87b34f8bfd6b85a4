"""Stage 5 (HOA opacity mask) and the drop-in call sequence of OcRFDet's render() wrapper."""
import math

import numpy as np
import pytest
import torch

from oracle import oracle
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,H,W,K", [(2, 80, 128, 128, 7), (1, 5, 20, 33, 3), (3, 16, 40, 24, 7)])
def test_opacity_mask_forward_backward(B, C, H, W, K):
    from ocrfdet_b200.opacity_lift import opacity_mask
    rng = np.random.default_rng(B * 100 + C)
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    w = (rng.normal(size=(1, 2, K, K)) * 0.2).astype(np.float32)
    ob = rng.normal(size=(B, 1, H, W)).astype(np.float32)
    go = rng.normal(size=(B, C, H, W)).astype(np.float32)
    xt, wt, ot = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (x, w, ob))
    out, mask = opacity_mask(xt, wt, ot)
    wo, wm, ws = oracle.opacity_mask_forward(x, w, ob)
    assert np.abs(out.detach().cpu().numpy() - wo).max() <= 1e-5 * (1 + np.abs(wo).max())
    assert np.abs(mask.detach().cpu().numpy() - wm).max() <= 1e-5
    (out * torch.from_numpy(go).cuda()).sum().backward()
    gx, gw, gop = oracle.opacity_mask_backward(x, w, wm, ws, go)
    assert util.rel_err(xt.grad.cpu().numpy(), gx) <= 1e-4
    assert util.rel_err(wt.grad.cpu().numpy().reshape(2, K, K), gw) <= 1e-4
    assert util.rel_err(ot.grad.cpu().numpy(), gop) <= 1e-4
    # and against the reference's own torch formulation (view_transformer_ocrf.py:230-242)
    x2, w2, o2 = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (x, w, ob))
    s = torch.cat([x2.mean(1, keepdim=True), x2.max(1, keepdim=True)[0]], 1)
    ref_out = x2 * torch.sigmoid(torch.nn.functional.conv2d(s, w2, padding=K // 2) + o2)
    assert float((ref_out - out).abs().max()) <= 1e-5 * (1 + float(ref_out.abs().max()))


def test_geom_attention_gate_is_the_same_fused_op():
    """BEVGeomAttention + its multiply (view_transformer_ocrf.py:215-228, 1190) through GeomAttentionGate."""
    from ocrfdet_b200.opacity_lift import GeomAttentionGate
    torch.manual_seed(3)
    B, C, H, W = 2, 80, 128, 128
    gate = GeomAttentionGate().cuda()
    assert list(gate.state_dict()) == ["conv1.weight"]
    x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    logit = torch.randn(B, 1, H, W, device="cuda", requires_grad=True)
    out = gate(x, logit)
    # float64 torch formulation (cuDNN's float32 convolutions run in TF32 on this GPU)
    x2, l2 = x.detach().double().requires_grad_(True), logit.detach().double().requires_grad_(True)
    s = torch.cat([x2.mean(1, keepdim=True), x2.max(1, keepdim=True)[0]], 1)
    ref = torch.sigmoid(torch.nn.functional.conv2d(s, gate.conv1.weight.detach().double(), padding=3) + l2) * x2
    assert float((out.detach() - ref.detach()).abs().max()) <= 1e-5 * (1 + float(ref.abs().max()))
    g = torch.randn_like(out)
    out.backward(g)
    ref.backward(g.double())
    assert util.rel_err(x.grad.cpu().numpy(), x2.grad.cpu().numpy()) <= 1e-4
    assert util.rel_err(logit.grad.cpu().numpy(), l2.grad.cpu().numpy()) <= 1e-4


def test_two_tuple_flavour_of_the_vendored_package():
    """`rendered_image, radii = rasterizer(...)` as MVSGaussian's gaussian_renderer_ft unpacks it (PKG:98)."""
    from diff_gaussian_rasterization import GaussianRasterizer
    W, H = 96, 48
    g, cams = util.small_scene("frustum", P=500, seed=3, W=W, H=H)
    gc = util.to_cuda(g)
    st = util.settings_for(cams[0], [0.0, 0.0, 0.0])
    kw = dict(means3D=gc["means3D"], means2D=torch.zeros_like(gc["means3D"]), opacities=gc["opacities"],
              colors_precomp=gc["colors"], scales=gc["scales"], rotations=gc["rotations"])
    image, radii = GaussianRasterizer(st, return_depth=False)(**kw)
    image3, radii3, depth3 = GaussianRasterizer(st)(**kw)
    assert torch.equal(image, image3) and torch.equal(radii, radii3) and tuple(depth3.shape) == (1, H, W)
    image4, radii4, opacity4 = GaussianRasterizer(st, return_depth=False, return_opacity=True)(**kw)
    assert torch.equal(image4, image3) and tuple(opacity4.shape) == (1, H, W)


def test_dropin_render_wrapper_call_sequence():
    """OcRFDet's own `render()` (gaussian_renderer/__init__.py:17-75, called from view_transformer_ocrf.py:1153) against
    the import name it uses: the UNMODIFIED file, imported where it lies, when the reference tree is mounted; its
    restated call sequence (tests/util.py::replay_render, proven argument-identical in the CPU suite) otherwise."""
    W, H = 176, 64
    g, cams = util.small_scene("ring", P=8000, seed=31, W=W, H=H, n_views=3)
    cam = cams[2]
    gc = util.to_cuda(g)
    data = {"FovX": torch.tensor(2 * math.atan(cam["tanfovx"])).cuda(), "FovY": torch.tensor(2 * math.atan(cam["tanfovy"])).cuda(),
            "height": H, "width": W, "world_view_transform": torch.from_numpy(cam["viewmatrix"]).cuda(),
            "full_proj_transform": torch.from_numpy(cam["projmatrix"]).cuda(),
            "camera_center": torch.from_numpy(cam["campos"]).cuda()}
    pts_xyz, pts_rgb = gc["means3D"], gc["colors"].requires_grad_(True)
    render = util.load_reference_render() or util.replay_render
    rendered_image, rendered_depth = render(data, 2, pts_xyz, pts_rgb, gc["rotations"], gc["scales"], gc["opacities"],
                                            bg_color=[0, 0, 0])
    assert rendered_image.shape == (3, H, W) and rendered_depth.shape == (1, H, W)
    rendered_image.mean().backward()
    assert float(pts_rgb.grad.abs().sum()) > 0
    cam2 = dict(cam, tanfovx=math.tan(data["FovX"] * 0.5), tanfovy=math.tan(data["FovY"] * 0.5))
    want, _ = util.oracle_forward(g, cam2, W, H, [0, 0, 0])
    util.assert_image_close(rendered_image.detach().cpu().numpy(), want["color"], want["ambiguous"], 1e-5, "render()")
