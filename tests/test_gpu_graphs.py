"""CUDA-graph replay of forward + backward equals the eager step."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def test_graphed_step_matches_eager_and_tracks_new_inputs():
    from ocrfdet_b200 import rasterizer as R
    from ocrfdet_b200.graphs import GraphedRenderStep, INPUTS
    W, H, P, V = 176, 64, 6000, 3
    g, cams = util.small_scene("ring", P=P, seed=41, W=W, H=H, n_views=V)
    cam_t = util.cams_tensor(cams)
    gc = util.to_cuda(g)
    gcol = torch.randn(V, 3, H, W, device="cuda")
    gop = torch.randn(V, 1, H, W, device="cuda")

    def eager(t):
        leaf = {k: t[k].clone().unsqueeze(0).requires_grad_(True) for k in INPUTS}
        color, radii, depth, opac = R.render_batch(leaf["means3D"], leaf["opacities"], cam_t, H, W, torch.zeros(3, device="cuda"),
                                                   colors_precomp=leaf["colors"], scales=leaf["scales"],
                                                   rotations=leaf["rotations"])
        torch.autograd.backward([color, opac], [gcol, gop])
        return (color, radii, depth, opac), {k: leaf[k].grad for k in INPUTS}

    (c0, r0, d0, o0), g0 = eager(gc)
    step = GraphedRenderStep(S=1, P=P, cams=cam_t, height=H, width=W, channels=3, pair_capacity=600_000)
    ins = {k: gc[k].unsqueeze(0) for k in INPUTS}
    (c1, r1, d1, o1), g1 = step(grad_color=gcol, grad_opacity=gop, **ins)
    step.check_overflow()
    assert torch.equal(c0, c1) and torch.equal(r0, r1) and torch.equal(d0, d1) and torch.equal(o0, o1)
    for k in INPUTS:
        assert util.rel_err(g1[k].cpu().numpy(), g0[k].cpu().numpy()) <= 1e-6, k
    # replay with different inputs: the graph reads the static buffers, so new data gives new (correct) results
    g2, _ = util.small_scene("ring", P=P, seed=42, W=W, H=H, n_views=V)
    gc2 = util.to_cuda(g2)
    (c2e, r2e, d2e, o2e), g2e = eager(gc2)
    (c2, r2, d2, o2), gg2 = step(grad_color=gcol, grad_opacity=gop, **{k: gc2[k].unsqueeze(0) for k in INPUTS})
    step.check_overflow()
    assert torch.equal(c2e, c2) and torch.equal(r2e, r2) and torch.equal(o2e, o2)
    for k in INPUTS:
        assert util.rel_err(gg2[k].cpu().numpy(), g2e[k].cpu().numpy()) <= 1e-6, k
    assert not torch.equal(c0, c2)


def test_graphed_step_reports_overflow():
    from ocrfdet_b200 import _lib
    from ocrfdet_b200.graphs import GraphedRenderStep, INPUTS
    W, H, P, V = 176, 64, 6000, 3
    g, cams = util.small_scene("ring", P=P, seed=41, W=W, H=H, n_views=V)
    gc = util.to_cuda(g)
    step = GraphedRenderStep(S=1, P=P, cams=util.cams_tensor(cams), height=H, width=W, channels=3, pair_capacity=1000)
    (c, _, _, o), _ = step(**{k: gc[k].unsqueeze(0) for k in INPUTS})
    assert float(c.abs().max()) == 0.0 and float(o.abs().max()) == 0.0   # background only
    with pytest.raises(_lib.OcrfError):
        step.check_overflow()
    with pytest.raises(ValueError):
        GraphedRenderStep(S=1, P=P, cams=util.cams_tensor(cams), height=H, width=W)


def test_graphed_step_with_captured_host_copies_is_a_host_to_host_step():
    """`pre` / `post` hooks: pinned host parameters in, gradients out, ONE graph launch (what bench.py's `e2e` times)."""
    from ocrfdet_b200 import rasterizer as R
    from ocrfdet_b200.graphs import GraphedRenderStep, INPUTS
    W, H, P, V = 176, 64, 6000, 3
    g, cams = util.small_scene("ring", P=P, seed=43, W=W, H=H, n_views=V)
    cam_t = util.cams_tensor(cams)
    host = {k: torch.from_numpy(g[k]).unsqueeze(0).contiguous().pin_memory() for k in INPUTS}
    ghost = {k: torch.zeros_like(host[k]).pin_memory() for k in INPUTS}
    gcol = torch.randn(V, 3, H, W, device="cuda")
    gop = torch.randn(V, 1, H, W, device="cuda")

    def pre(st):
        for k in INPUTS:
            st.static_in[k].copy_(host[k], non_blocking=True)

    def post(st, outs, grads):
        for k in INPUTS:
            ghost[k].copy_(grads[k], non_blocking=True)

    step = GraphedRenderStep(S=1, P=P, cams=cam_t, height=H, width=W, channels=3, pair_capacity=600_000, pre=pre, post=post)
    step.capture(grad_color=gcol, grad_opacity=gop)
    # new host data AFTER the capture: the replay must pick it up from the same pinned buffers
    g2, _ = util.small_scene("ring", P=P, seed=44, W=W, H=H, n_views=V)
    for k in INPUTS:
        host[k].copy_(torch.from_numpy(g2[k]).unsqueeze(0))
        ghost[k].zero_()
    step.graph.replay()
    torch.cuda.synchronize()
    step.check_overflow()
    leaf = {k: host[k].cuda().requires_grad_(True) for k in INPUTS}
    color, radii, depth, opac = R.render_batch(leaf["means3D"], leaf["opacities"], cam_t, H, W, torch.zeros(3, device="cuda"),
                                               colors_precomp=leaf["colors"], scales=leaf["scales"], rotations=leaf["rotations"])
    torch.autograd.backward([color, opac], [gcol, gop])
    assert torch.equal(step.outs[0], color)
    for k in INPUTS:
        assert util.rel_err(ghost[k].numpy(), leaf[k].grad.cpu().numpy()) <= 1e-6, k


def test_graphed_step_with_80_channels_runs_the_tensor_core_kernels_and_survives_overflow():
    """The tcgen05 blend kernels under CUDA-graph capture (tensor-memory allocation, mbarrier pipelines and the
    dependent-launch edges are all inside the graph), equal to the eager step; and with a capacity far below the pair
    count they render background (empty lists: no product is ever issued) and the overflow is reported."""
    from ocrfdet_b200 import _lib, rasterizer as R
    from ocrfdet_b200.graphs import GraphedRenderStep, INPUTS
    W, H, P, V, C = 176, 64, 4000, 2, 80
    g, cams = util.small_scene("ring", P=P, seed=61, W=W, H=H, n_views=V, channels=C)
    cam_t = util.cams_tensor(cams)
    gc = util.to_cuda(g)
    bg = torch.linspace(0.0, 1.0, C, device="cuda")
    gcol = torch.randn(V, C, H, W, device="cuda")
    gop = torch.randn(V, 1, H, W, device="cuda")
    leaf = {k: gc[k].clone().unsqueeze(0).requires_grad_(True) for k in INPUTS}
    color, radii, depth, opac = R.render_batch(leaf["means3D"], leaf["opacities"], cam_t, H, W, bg,
                                               colors_precomp=leaf["colors"], scales=leaf["scales"], rotations=leaf["rotations"])
    torch.autograd.backward([color, opac], [gcol, gop])
    step = GraphedRenderStep(S=1, P=P, cams=cam_t, height=H, width=W, channels=C, pair_capacity=400_000, bg=bg)
    (c1, r1, d1, o1), g1 = step(grad_color=gcol, grad_opacity=gop, **{k: gc[k].unsqueeze(0) for k in INPUTS})
    step.check_overflow()
    assert torch.equal(color, c1) and torch.equal(depth, d1) and torch.equal(opac, o1)
    for k in INPUTS:
        assert util.rel_err(g1[k].cpu().numpy(), leaf[k].grad.cpu().numpy()) <= 2e-6, k
    small = GraphedRenderStep(S=1, P=P, cams=cam_t, height=H, width=W, channels=C, pair_capacity=500, bg=bg)
    (c2, _, _, o2), g2 = small(grad_color=gcol, grad_opacity=gop, **{k: gc[k].unsqueeze(0) for k in INPUTS})
    assert torch.equal(c2, bg.view(1, C, 1, 1).expand(V, C, H, W)) and float(o2.abs().max()) == 0.0
    assert all(float(g2[k].abs().max()) == 0.0 for k in INPUTS)
    with pytest.raises(_lib.OcrfError):
        small.check_overflow()
