"""CPU suite: the numpy oracle of the HOA lift (oracle/hoa.py) against the golden vectors produced by the reference's
own modules (tests/golden/make_golden_hoa.py: DeformableAttention2D imported where it lies, OpacityVoxelToBEVConverter
+ HeightAttention lifted with ast), forward and every gradient."""
import os

import numpy as np
import pytest

from oracle import hoa
from tests.golden.make_golden_hoa import STRIDE, hoa_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol, what, floor=0.0):
    """max |a - b| <= tol * max(max |b|, floor).  `floor`: the scale below which a gradient is rounding noise (the bias
    in front of a softmax has a mathematically ZERO gradient; autograd returns ~1e-9 of its neighbours)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    err = np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-12)
    assert err <= tol, "%s: rel err %.3g" % (what, err)


@pytest.mark.parametrize("name", ["hoa_lift_b2", "hoa_lift_b1_s96"])
def test_lift_oracle_matches_reference_modules(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    io = hoa_inputs(int(g["seed"]), int(g["B"]), S=int(g["S"]))
    params = {k[2:]: g[k] for k in g.files if k.startswith("p.")}
    out, cache = hoa.lift_forward(params, io["opacity"], io["alpha"])
    _close(cache["opacity_up"], g["opacity_up"], 1e-6, "opacity_up")
    _close(cache["alpha_up"], g["alpha_up"], 1e-6, "alpha_up")
    _close(cache["vgrid"], g["vgrid"], 1e-5, "sampling grid")
    _close(cache["att"], g["att"], 1e-5, "attention output")
    _close(out[..., ::STRIDE, ::STRIDE], g["opacity_alpha_sub"], 1e-5, "opacity_alpha")
    assert abs(out.sum() - float(g["opacity_alpha_sum"])) <= 1e-6 * abs(float(g["opacity_alpha_sum"]))
    g_op, g_al, G = hoa.lift_backward(params, cache, io["g_lift"])
    _close(g_op[..., ::STRIDE, ::STRIDE], g["g_opacity_sub"], 1e-4, "d/d opacity")
    _close(g_al[..., ::STRIDE, ::STRIDE], g["g_alpha_sub"], 1e-4, "d/d alpha")
    assert abs(np.abs(g_al).sum() - float(g["g_alpha_abs_sum"])) <= 1e-4 * float(g["g_alpha_abs_sum"])
    floor = 1e-3 * max(np.abs(g["g." + k]).max() for k in params)
    for k in params:
        _close(G[k], g["g." + k], 2e-4, "d/d " + k, floor)


@pytest.mark.parametrize("name", ["hoa_converter_train_b2", "hoa_converter_eval_b1"])
def test_converter_oracle_matches_reference_modules(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    io = hoa_inputs(int(g["seed"]), int(g["B"]), S=int(g["S"]))
    params = {k[2:]: g[k] for k in g.files if k.startswith("p.")}
    buffers = {k[3:]: g[k] for k in g.files if k.startswith("b0.")}
    train = bool(int(g["train"]))
    x = io["opacity"] + np.float32(0.25) * io["alpha"]
    out, cache, stats = hoa.converter_forward(params, x, io["position"], train=train, buffers=buffers)
    _close(out, g["out"], 1e-5, "BEV opacity logit")
    if train:  # running statistics: momentum 0.1, unbiased variance (torch.nn.BatchNorm2d)
        n = x.shape[0] * x.shape[2] * x.shape[3]
        mean, var = stats["encoder1"]
        _close(0.9 * buffers["encoder1.2.running_mean"] + 0.1 * mean, g["b1.encoder1.2.running_mean"], 1e-5, "running mean")
        _close(0.9 * buffers["encoder1.2.running_var"] + 0.1 * var * n / (n - 1), g["b1.encoder1.2.running_var"], 1e-5,
               "running var")
    g_x, g_pos, G = hoa.converter_backward(params, cache, io["g_bev"])
    _close(g_x[..., ::STRIDE, ::STRIDE], g["g_x_sub"], 1e-4, "d/d x")
    assert abs(np.abs(g_x).sum() - float(g["g_x_abs_sum"])) <= 1e-4 * float(g["g_x_abs_sum"])
    _close(g_pos, g["g_pos"], 1e-4, "d/d position")
    # (biases in front of a training-mode batch norm have a mathematically zero gradient: float32 cancellation noise)
    floor = 1e-2 * max(np.abs(g["g." + k]).max() for k in params)
    for k in params:
        _close(G[k], g["g." + k], 3e-4, "d/d " + k, floor)
