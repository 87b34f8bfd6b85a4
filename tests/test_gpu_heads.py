"""Scope row f-1: the fused Gaussian-construction heads against the C oracle, the golden vectors recorded from
the reference's module code, and the reference's torch formulation at full size."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import oracle
from tests import util
from tests.test_oracle_golden import GOLDEN_HEADS, heads_params

pytestmark = pytest.mark.gpu


def load_heads(params, Fd):
    from ocrfdet_b200.gaussian_heads import GaussianHeads
    m = GaussianHeads(Fd).cuda()
    sd = {"%s.%s" % (h, k): torch.from_numpy(np.asarray(v)) for h, d in params.items() for k, v in d.items()}
    m.load_state_dict(sd)   # reference parameter names
    return m


def random_params(rng, Fd, scale=1.0):
    out = {}
    for h, o in zip(oracle.HEAD_ORDER, oracle.HEAD_OUTS):
        fin = Fd + 3 if h == "C_MLP" else Fd
        out[h] = {"fc1.weight": (rng.normal(size=(4, fin)) * scale / np.sqrt(fin)).astype(np.float32),
                  "fc1.bias": (rng.normal(size=4) * 0.1).astype(np.float32),
                  "fc2.weight": (rng.normal(size=(o, 4)) * scale * 0.5).astype(np.float32),
                  "fc2.bias": (rng.normal(size=o) * 0.1).astype(np.float32)}
    return out


@pytest.mark.parametrize("path", GOLDEN_HEADS, ids=[os.path.basename(p) for p in GOLDEN_HEADS])
def test_heads_match_reference_golden(path):
    z = np.load(path)
    Fd = z["feat"].shape[1]
    m = load_heads(heads_params(z), Fd)
    feat = torch.from_numpy(z["feat"]).cuda().requires_grad_(True)
    outs = m(feat, torch.from_numpy(z["rgb"]).cuda())
    for got, key in zip(outs, ("opacity", "scaling", "rotation", "color")):
        assert got.shape == z[key].shape
        assert np.abs(got.detach().cpu().numpy() - z[key]).max() <= 1e-5 * (1 + np.abs(z[key]).max()), key
    sum((o * torch.from_numpy(z[k]).cuda()).sum()
        for o, k in zip(outs, ("g_opacity", "g_scaling", "g_rotation", "g_color"))).backward()
    assert util.rel_err(feat.grad.cpu().numpy(), z["g_feat"]) <= 1e-4
    for name, gp in m.reference_parameters(grads=True).items():
        assert util.rel_err(gp.cpu().numpy(), z["g." + name]) <= 1e-4, name


@pytest.mark.parametrize("n,Fd,seed", [(1, 80, 0), (127, 80, 1), (129, 80, 2), (1000, 5, 3), (700, 114, 4), (300, 125, 6), (515, 84, 7), (4097, 33, 5)])
def test_heads_match_oracle(n, Fd, seed):
    rng = np.random.default_rng(seed)
    params = random_params(rng, Fd, scale=2.0)
    feat = rng.normal(size=(n, Fd)).astype(np.float32)
    rgb = rng.uniform(size=(n, 3)).astype(np.float32)
    gs = [rng.normal(size=(n, w)).astype(np.float32) for w in (1, 3, 4, 3)]
    m = load_heads(params, Fd)
    ft = torch.from_numpy(feat).cuda().requires_grad_(True)
    outs = m(ft, torch.from_numpy(rgb).cuda())
    want = oracle.gaussian_heads_forward(feat, rgb, params)
    for got, w in zip(outs, want):
        assert np.abs(got.detach().cpu().numpy() - w).max() <= 1e-5 * (1 + np.abs(w).max())
    sum((o * torch.from_numpy(g).cuda()).sum() for o, g in zip(outs, gs)).backward()
    g_feat, grads = oracle.gaussian_heads_backward(feat, rgb, params, *gs)
    assert util.rel_err(ft.grad.cpu().numpy(), g_feat) <= 1e-4
    for name, gp in m.reference_parameters(grads=True).items():
        h, k = name.split(".", 1)
        assert util.rel_err(gp.cpu().numpy(), grads[h][k]) <= 1e-4, name
    # the S/R/A heads never see rgb: their packed rgb weights must keep a zero gradient
    from ocrfdet_b200.gaussian_heads import _views
    assert float(_views(m.packed.grad, Fd)["w1t"][Fd:, :12].abs().max()) == 0.0


def test_heads_empty_and_bad_arguments():
    from ocrfdet_b200.gaussian_heads import GaussianHeads, gaussian_heads
    m = GaussianHeads(80).cuda()
    outs = m(torch.zeros(0, 80, device="cuda"), torch.zeros(0, 3, device="cuda"))
    assert [tuple(o.shape) for o in outs] == [(0, 1), (0, 3), (0, 4), (0, 3)]
    with pytest.raises(ValueError):
        m(torch.zeros(4, 80, device="cuda"), torch.zeros(5, 3, device="cuda"))
    with pytest.raises(ValueError):
        gaussian_heads(torch.zeros(4, 200, device="cuda"), torch.zeros(4, 3, device="cuda"), torch.zeros(3328, device="cuda"))
    with pytest.raises(ValueError):
        gaussian_heads(torch.zeros(4, 80, device="cuda"), torch.zeros(4, 3, device="cuda"), torch.zeros(77, device="cuda"))
    with pytest.raises(Exception):
        m.cpu()(torch.zeros(4, 80), torch.zeros(4, 3))   # no CPU path


class _RefHead(nn.Module):  # the reference's formulation (view_transformer_ocrf.py:272-320), restated for the full-size check
    def __init__(self, fin, out, act):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(fin, 4), nn.Linear(4, out), act

    def forward(self, x):
        return self.act(self.fc2(F.relu(self.fc1(x))))


def test_heads_full_size_against_torch_formulation():
    """BASELINE config 2's voxel grid: [2 samples, 212992 voxels, 80 channels], batched leading dims."""
    from ocrfdet_b200.gaussian_heads import GaussianHeads
    torch.manual_seed(5)
    S, n, Fd = 2, 212992, 80
    acts = {"S_MLP": F.softplus, "R_MLP": lambda x: F.normalize(x, dim=-1), "A_MLP": torch.sigmoid, "C_MLP": torch.sigmoid}
    ref = nn.ModuleDict({h: _RefHead(Fd + 3 if h == "C_MLP" else Fd, o, acts[h])
                         for h, o in zip(oracle.HEAD_ORDER, oracle.HEAD_OUTS)}).cuda()
    m = GaussianHeads(Fd).cuda()
    m.load_state_dict({k.replace(".act", ""): v for k, v in ref.state_dict().items()})
    feat = torch.randn(S, n, Fd, device="cuda")
    rgb = torch.rand(S, n, 3, device="cuda")
    f1, f2 = feat.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    got = m(f1, rgb)
    want = (ref["A_MLP"](f2), ref["S_MLP"](f2), ref["R_MLP"](f2), ref["C_MLP"](torch.cat((f2, rgb), -1)))
    gs = [torch.randn_like(w) for w in want]
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert float((g - w).abs().max()) <= 1e-5 * (1 + float(w.abs().max()))
    sum((o * g).sum() for o, g in zip(got, gs)).backward()
    sum((o * g).sum() for o, g in zip(want, gs)).backward()
    assert util.rel_err(f1.grad.cpu().numpy(), f2.grad.cpu().numpy()) <= 1e-4
    rp = dict(ref.named_parameters())
    for name, gp in m.reference_parameters(grads=True).items():
        # 425 984-term float32 sums: compare with the tolerance of the reference's own (cuBLAS) summation order
        assert util.rel_err(gp.cpu().numpy(), rp[name].grad.cpu().numpy()) <= 2e-4, name
    # checkpoint keys are the reference's
    assert set(m.state_dict()) == set(k.replace(".act", "") for k in ref.state_dict())
