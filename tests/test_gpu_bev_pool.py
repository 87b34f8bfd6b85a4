"""Scope row f-3: BEV pooling v2 against the C oracle, the reference's own CUDA kernels (oracle/_ref) and a dense
torch formulation.  Forward and feat_grad follow the reference's summation order: bit-exact.  depth_grad is a
warp-parallel dot product: 1e-5 relative."""
import numpy as np
import pytest
import torch

from ocrfdet_b200.scenes import bev_pool_case
from oracle import oracle, ref
from tests import util

pytestmark = pytest.mark.gpu

CASES = [dict(B=1, N=2, D=24, H=8, W=22, C=80, bev=64, seed=1), dict(B=2, N=3, D=16, H=6, W=10, C=19, bev=32, seed=2),
         dict(B=1, N=6, D=30, H=8, W=16, C=128, bev=48, seed=3), dict(B=1, N=1, D=4, H=2, W=3, C=4, bev=8, seed=4),
         dict(B=2, N=6, D=88, H=16, W=44, C=80, bev=128, seed=5)]


def run_ours(c, out_grad):
    from ocrfdet_b200.bev_pool import bev_pool_v2
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items() if isinstance(v, np.ndarray)}
    depth, feat = t["depth"].requires_grad_(True), t["feat"].requires_grad_(True)
    out = bev_pool_v2(depth, feat, t["ranks_depth"], t["ranks_feat"], t["ranks_bev"], c["bev_feat_shape"],
                      t["interval_starts"], t["interval_lengths"])
    B, Z, Y, X, C = c["bev_feat_shape"]
    assert tuple(out.shape) == (B, C, Z, Y, X) and out.is_contiguous()
    og = torch.from_numpy(out_grad).cuda().reshape(B, Z, Y, X, C).permute(0, 4, 1, 2, 3)
    out.backward(og)
    flat = out.permute(0, 2, 3, 4, 1).reshape(-1, C)
    return flat.detach().cpu().numpy(), depth.grad.cpu().numpy(), feat.grad.cpu().numpy(), t


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_bev_pool_matches_oracle_and_reference(case):
    c = bev_pool_case(**case)
    og = np.random.default_rng(case["seed"] + 7).normal(size=(c["n_bev"], case["C"])).astype(np.float32)
    out, dg, fg, t = run_ours(c, og)
    want = oracle.bev_pool_forward(c["depth"], c["feat"], c["ranks_depth"], c["ranks_feat"], c["ranks_bev"], c["n_bev"],
                                   c["interval_starts"], c["interval_lengths"])
    assert np.array_equal(out, want), "forward must be bit-exact (same summation order, same fma contraction)"
    wdg, wfg = oracle.bev_pool_backward(og, c["depth"], c["feat"], c["ranks_depth"], c["ranks_feat"], c["ranks_bev"])
    assert np.array_equal(fg, wfg), "feat_grad must be bit-exact"
    assert util.rel_err(dg, wdg) <= 1e-5
    assert np.array_equal(dg == 0, wdg == 0) or util.rel_err(dg, wdg) <= 1e-5   # untouched depth bins stay exactly zero
    if ref.bev_available():
        rout = ref.bev_pool_forward(t["depth"].detach(), t["feat"].detach(), t["ranks_depth"], t["ranks_feat"], t["ranks_bev"],
                                    c["n_bev"], t["interval_starts"], t["interval_lengths"])
        assert np.array_equal(out, rout.cpu().numpy())
        rdg, rfg = ref.bev_pool_backward(torch.from_numpy(og).cuda(), t["depth"].detach(), t["feat"].detach(),
                                         t["ranks_depth"], t["ranks_feat"], t["ranks_bev"])
        assert np.array_equal(fg, rfg.cpu().numpy())
        assert util.rel_err(dg, rdg.cpu().numpy()) <= 1e-5


def test_bev_pool_against_dense_torch_autograd():
    from ocrfdet_b200.bev_pool import bev_pool_v2
    case = dict(B=1, N=3, D=12, H=6, W=9, C=16, bev=24, seed=9)
    c = bev_pool_case(**case)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items() if isinstance(v, np.ndarray)}
    d1, f1 = t["depth"].double().requires_grad_(True), t["feat"].double().requires_grad_(True)
    dense = torch.zeros(c["n_bev"], case["C"], dtype=torch.float64, device="cuda")
    dense = dense.index_add(0, t["ranks_bev"].long(), d1.reshape(-1)[t["ranks_depth"].long()][:, None]
                            * f1.reshape(-1, case["C"])[t["ranks_feat"].long()])
    B, Z, Y, X, C = c["bev_feat_shape"]
    dense = dense.reshape(B, Z, Y, X, C).permute(0, 4, 1, 2, 3)
    g = torch.randn_like(dense)
    (dense * g).sum().backward()
    d2, f2 = t["depth"].clone().requires_grad_(True), t["feat"].clone().requires_grad_(True)
    out = bev_pool_v2(d2, f2, t["ranks_depth"], t["ranks_feat"], t["ranks_bev"], c["bev_feat_shape"], t["interval_starts"],
                      t["interval_lengths"])
    (out * g.float()).sum().backward()
    assert float((out - dense).abs().max()) <= 1e-5 * (1 + float(dense.abs().max()))
    assert util.rel_err(d2.grad.cpu().numpy(), d1.grad.cpu().numpy()) <= 1e-5
    assert util.rel_err(f2.grad.cpu().numpy(), f1.grad.cpu().numpy()) <= 1e-5


def test_bev_pool_edge_cases():
    from ocrfdet_b200.bev_pool import bev_pool_v2
    dev = "cuda"
    e = lambda: torch.zeros(0, dtype=torch.int32, device=dev)  # noqa: E731
    depth = torch.rand(1, 1, 2, 2, 2, device=dev, requires_grad=True)
    feat = torch.rand(1, 1, 2, 2, 8, device=dev, requires_grad=True)
    out = bev_pool_v2(depth, feat, e(), e(), e(), (1, 1, 4, 4, 8), e(), e())     # no point inside the grid
    assert tuple(out.shape) == (1, 8, 1, 4, 4) and float(out.abs().max()) == 0.0
    out.sum().backward()
    assert float(depth.grad.abs().max()) == 0.0 and float(feat.grad.abs().max()) == 0.0
    # several points sharing one depth bin (the height-sampling variant, view_transformer_ocrf.py:784-848):
    # forward and feat_grad are well defined; depth_grad is a last-writer-wins store in the reference as well
    rb = torch.tensor([0, 0, 3, 3, 3], dtype=torch.int32, device=dev)
    rd = torch.tensor([1, 1, 2, 5, 2], dtype=torch.int32, device=dev)
    rf = torch.tensor([0, 0, 1, 2, 1], dtype=torch.int32, device=dev)
    st = torch.tensor([0, 2], dtype=torch.int32, device=dev)
    ln = torch.tensor([2, 3], dtype=torch.int32, device=dev)
    depth.grad = feat.grad = None
    out = bev_pool_v2(depth, feat, rd, rf, rb, (1, 1, 2, 2, 8), st, ln)
    d, f = depth.detach().reshape(-1), feat.detach().reshape(-1, 8)
    want0 = 2 * d[1] * f[0]
    want3 = d[2] * f[1] * 2 + d[5] * f[2]
    flat = out.permute(0, 2, 3, 4, 1).reshape(-1, 8)
    assert torch.allclose(flat[0], want0, atol=1e-6) and torch.allclose(flat[3], want3, atol=1e-6)
    out.sum().backward()
    assert torch.allclose(feat.grad.reshape(-1, 8)[0], (2 * d[1]).expand(8), atol=1e-6)
    assert torch.allclose(feat.grad.reshape(-1, 8)[1], (2 * d[2]).expand(8), atol=1e-6)
    assert float(feat.grad.reshape(-1, 8)[3].abs().max()) == 0.0
    with pytest.raises(Exception):
        bev_pool_v2(depth.cpu(), feat.cpu(), rd.cpu(), rf.cpu(), rb.cpu(), (1, 1, 2, 2, 8), st.cpu(), ln.cpu())
