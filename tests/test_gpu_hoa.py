"""GPU parity of the HOA lift (stage 5): the fused kernels through the C ABI against the numpy oracle (oracle/hoa.py)
and against the golden vectors the reference's own modules produced (tests/golden/make_golden_hoa.py).
Forward 1e-5, gradients 1e-4 (relative to the largest element; parameters whose gradient is mathematically zero are
judged against the scale of their neighbours)."""
import os

import numpy as np
import pytest
import torch

from ocrfdet_b200 import hoa_lift as HL
from oracle import hoa
from tests.golden.make_golden_hoa import STRIDE, hoa_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol, what, floor=0.0):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    err = np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-12)
    assert err <= tol, "%s: rel err %.3g" % (what, err)


def _attention_from(params):
    m = HL.DeformableAttention2D(dim=13, dim_head=8, heads=1, dropout=0.1, downsample_factor=4, offset_scale=4,
                                 offset_groups=None, offset_kernel_size=6)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    return m.cuda()


@pytest.mark.parametrize("name", ["hoa_lift_b2", "hoa_lift_b1_s96"])
def test_lift_against_reference_golden_and_oracle(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    io = hoa_inputs(int(g["seed"]), int(g["B"]), S=int(g["S"]))
    params = {k[2:]: g[k] for k in g.files if k.startswith("p.")}
    attn = _attention_from(params).eval()
    op = torch.from_numpy(io["opacity"]).cuda().requires_grad_(True)
    al = torch.from_numpy(io["alpha"]).cuda().requires_grad_(True)
    out = HL.opacity_alpha_lift(op, al, attn)
    want, cache = hoa.lift_forward(params, io["opacity"], io["alpha"])
    _close(out, want, 1e-5, "opacity_alpha vs oracle")
    _close(out[..., ::STRIDE, ::STRIDE], g["opacity_alpha_sub"], 1e-5, "opacity_alpha vs reference modules")
    out.backward(torch.from_numpy(io["g_lift"]).cuda())
    g_op, g_al, G = hoa.lift_backward(params, cache, io["g_lift"])
    _close(op.grad, g_op, 1e-4, "d/d opacity vs oracle")
    _close(al.grad, g_al, 1e-4, "d/d alpha vs oracle")
    _close(al.grad[..., ::STRIDE, ::STRIDE], g["g_alpha_sub"], 1e-4, "d/d alpha vs reference modules")
    got = attn.reference_parameters(grads=True)
    floor = 1e-3 * max(np.abs(g["g." + k]).max() for k in params)
    for k in params:
        _close(got[k], G[k], 2e-4, "d/d %s vs oracle" % k, floor)
        _close(got[k], g["g." + k], 3e-4, "d/d %s vs reference modules" % k, floor)


def test_lift_dropout_mask_and_training_mode():
    """Dropout is an explicit keep-mask: the kernels with a given mask equal the oracle with the same mask; in training
    mode a mask is drawn (p = 0.1) and the result differs from eval mode."""
    g = np.load(os.path.join(GOLD, "hoa_lift_b2.npz"))
    io = hoa_inputs(77, 1, S=128)
    params = {k[2:]: g[k] for k in g.files if k.startswith("p.")}
    attn = _attention_from(params)
    nq, nk = HL.key_grid(128, 128)
    assert (nq, nk) == (441, 25)
    rng = np.random.default_rng(5)
    keep = ((rng.random((1, nq, nk)) >= 0.1) / 0.9).astype(np.float32)
    op = torch.from_numpy(io["opacity"]).cuda().requires_grad_(True)
    al = torch.from_numpy(io["alpha"]).cuda().requires_grad_(True)
    out = HL.opacity_alpha_lift(op, al, attn, keep=torch.from_numpy(keep).cuda())
    want, cache = hoa.lift_forward(params, io["opacity"], io["alpha"], keep=keep)
    _close(out, want, 1e-5, "opacity_alpha with a dropout mask")
    out.backward(torch.from_numpy(io["g_lift"]).cuda())
    g_op, g_al, G = hoa.lift_backward(params, cache, io["g_lift"])
    _close(op.grad, g_op, 1e-4, "d/d opacity with a dropout mask")
    _close(al.grad, g_al, 1e-4, "d/d alpha with a dropout mask")
    attn.train()
    torch.manual_seed(0)
    o_train = HL.opacity_alpha_lift(op.detach(), al.detach(), attn)
    attn.eval()
    o_eval = HL.opacity_alpha_lift(op.detach(), al.detach(), attn)
    assert not torch.equal(o_train, o_eval)
    with pytest.raises(Exception):
        HL.opacity_alpha_lift(op.detach().cpu(), al.detach().cpu(), attn)   # no CPU path


def test_lift_full_batch_against_torch_formulation():
    """Eight samples at the reference size against the three reference lines written with torch ops on the GPU
    (interpolate + the attention's own math via the oracle is too slow at B = 8; here: batch independence)."""
    g = np.load(os.path.join(GOLD, "hoa_lift_b2.npz"))
    params = {k[2:]: g[k] for k in g.files if k.startswith("p.")}
    attn = _attention_from(params).eval()
    io = hoa_inputs(9, 8, S=128)
    op, al = torch.from_numpy(io["opacity"]).cuda(), torch.from_numpy(io["alpha"]).cuda()
    full = HL.opacity_alpha_lift(op, al, attn)
    for b in (0, 3, 7):
        one = HL.opacity_alpha_lift(op[b:b + 1], al[b:b + 1], attn)
        assert torch.equal(one[0], full[b])


def _converter_from(g):
    m = HL.OpacityVoxelToBEVConverter(input_channel=13)
    sd = {}
    for k in g.files:
        if k.startswith("p."):
            sd[k[2:]] = torch.from_numpy(np.asarray(g[k]))
        elif k.startswith("b0."):
            sd[k[3:]] = torch.from_numpy(np.asarray(g[k]))
    m.load_state_dict(sd)
    return m.cuda()


@pytest.mark.parametrize("name", ["hoa_converter_train_b2", "hoa_converter_eval_b1"])
def test_converter_against_reference_golden_and_oracle(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    io = hoa_inputs(int(g["seed"]), int(g["B"]), S=int(g["S"]))
    params = {k[2:]: g[k] for k in g.files if k.startswith("p.")}
    buffers = {k[3:]: g[k] for k in g.files if k.startswith("b0.")}
    train = bool(int(g["train"]))
    conv = _converter_from(g).train(train)
    x_np = io["opacity"] + np.float32(0.25) * io["alpha"]
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    pos = torch.from_numpy(io["position"]).cuda().requires_grad_(True)
    out = conv(x, pos)
    want, cache, stats = hoa.converter_forward(params, x_np, io["position"], train=train, buffers=buffers)
    _close(out, want, 1e-5, "BEV opacity logit vs oracle")
    _close(out, g["out"], 2e-5, "BEV opacity logit vs reference modules")
    if train:  # running statistics were updated like nn.BatchNorm2d does
        sd = conv.state_dict()
        for k in ("encoder1.2.running_mean", "bottleneck.2.running_var", "decoder1.2.running_var"):
            _close(sd[k], g["b1." + k], 1e-4, k)
        assert int(sd["decoder2.2.num_batches_tracked"]) == int(g["b1.decoder2.2.num_batches_tracked"])
    out.backward(torch.from_numpy(io["g_bev"]).cuda())
    g_x, g_pos, G = hoa.converter_backward(params, cache, io["g_bev"])
    _close(x.grad, g_x, 1e-4, "d/d x vs oracle")
    _close(x.grad[..., ::STRIDE, ::STRIDE], g["g_x_sub"], 1e-4, "d/d x vs reference modules")
    _close(pos.grad, g_pos, 1e-4, "d/d position vs oracle")
    got = conv.reference_parameters(grads=True)
    floor = 1e-2 * max(np.abs(g["g." + k]).max() for k in params)
    for k in params:
        _close(got[k], G[k], 3e-4, "d/d %s vs oracle" % k, floor)
        _close(got[k], g["g." + k], 5e-4, "d/d %s vs reference modules" % k, floor)


def test_converter_batch_of_eight_and_batched_position():
    """The training shape (8 samples, 128 x 128): eval mode is sample-independent, so a batch equals its samples one by
    one; a per-sample position tensor [B,4,S,S] gives per-sample position gradients."""
    g = np.load(os.path.join(GOLD, "hoa_converter_eval_b1.npz"))
    conv = _converter_from(g).eval()
    io = hoa_inputs(21, 8, S=128)
    x = torch.from_numpy(io["opacity"]).cuda()
    pos = torch.from_numpy(io["position"]).cuda()
    full = conv(x, pos)
    for b in (0, 5):
        assert torch.allclose(conv(x[b:b + 1], pos)[0], full[b], rtol=0, atol=1e-6)
    posb = pos.expand(8, -1, -1, -1).contiguous().requires_grad_(True)
    pos1 = pos.clone().requires_grad_(True)
    gb = torch.from_numpy(io["g_bev"]).cuda()
    conv(x, posb).backward(gb)
    conv(x, pos1).backward(gb)
    _close(posb.grad.sum(0, keepdim=True), pos1.grad.cpu().numpy(), 1e-5, "position gradient, batched vs shared")
