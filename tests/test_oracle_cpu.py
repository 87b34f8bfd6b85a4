"""CPU suite: internal consistency of the oracle (C restatement vs the dense PyTorch autograd oracle),
its integer stages against numpy, and the stage-5 oracle against torch."""
import numpy as np
import pytest
import torch

from ocrfdet_b200.scenes import frustum_scene, ring_scene
from oracle import dense_torch, oracle
from tests import util


@pytest.mark.parametrize("seed,P,W,H", [(3, 400, 80, 48), (5, 150, 33, 21)])
def test_c_oracle_matches_dense_autograd_oracle(seed, P, W, H):
    g, cams = frustum_scene(P=P, seed=seed, width=W, height=H)
    cam = cams[0]
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    out, st = util.oracle_forward(g, cam, W, H, bg)
    T = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)  # noqa: E731
    m, s, r, o, c = T(g["means3D"]), T(g["scales"]), T(g["rotations"]), T(g["opacities"]), T(g["colors"])
    col, dep, opa, rad = dense_torch.render(m, s, r, o, c, cam, W, H, bg)
    amb = out["ambiguous"].astype(bool)
    assert np.array_equal(rad.numpy(), st["pre"]["radii"])
    util.assert_image_close(out["color"], col.detach().numpy(), amb, 1e-5, "colour")
    util.assert_image_close(out["opacity"], opa.detach().numpy(), amb, 1e-5, "opacity")
    assert np.array_equal(out["depth"][0][~amb], dep.detach().numpy()[0][~amb].astype(np.float32))
    rng = np.random.default_rng(seed)
    gc = rng.normal(size=(3, H, W)).astype(np.float32)
    go = rng.normal(size=(1, H, W)).astype(np.float32)
    ((col * torch.tensor(gc, dtype=torch.float64)).sum() + (opa * torch.tensor(go, dtype=torch.float64)).sum()).backward()
    for f64 in (True, False):
        gw = util.oracle_backward(g, cam, W, H, bg, out, st, gc, go, f64=f64)
        for name, ref in (("means3D", m.grad), ("scales", s.grad), ("rotations", r.grad),
                          ("opacities", o.grad.reshape(-1)), ("colors", c.grad)):
            assert util.rel_err(gw[name], ref.numpy()) <= 2e-5, "%s (f64=%s)" % (name, f64)


def test_sort_is_stable_and_matches_numpy():
    rng = np.random.default_rng(0)
    n = 50_000
    keys = (rng.integers(0, 704, n, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 64, n, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    k, v = oracle.sort_pairs(keys, vals, 42)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])
    # bits above end_bit must not take part
    k2, v2 = oracle.sort_pairs(keys | (np.uint64(1) << np.uint64(50)) * (vals % 2).astype(np.uint64), vals, 42)
    assert np.array_equal(v2, v)


def test_tile_ranges_and_higher_msb():
    keys = np.array([0, 0, 2, 2, 2, 5], dtype=np.uint64) << np.uint64(32)
    r = oracle.tile_ranges(keys, 8)
    assert r.tolist() == [[0, 2], [0, 0], [2, 5], [0, 0], [0, 0], [5, 6], [0, 0], [0, 0]]
    assert oracle.tile_ranges(np.zeros(0, np.uint64), 4).tolist() == [[0, 0]] * 4
    for n, want in ((0, 1), (1, 1), (2, 2), (3, 2), (704, 10), (1024, 11), (2816, 12), (4224, 13), (2 ** 31, 32)):
        assert oracle.higher_msb(n) == want


def test_duplicate_emits_row_major_tiles_in_scan_order():
    g, cams = ring_scene(P=3000, seed=1, width=176, height=64, n_views=1)
    cam = cams[0]
    _, st = util.oracle_forward(g, cam, 176, 64, [0, 0, 0])
    pre, b = st["pre"], st["bin"]
    assert b["N"] == int(pre["tiles_touched"].sum())
    ku, vu = b["keys_unsorted"], b["values_unsorted"]
    assert np.all(np.diff(vu.astype(np.int64)) >= 0)                      # Gaussian-major emission
    same = vu[1:] == vu[:-1]
    assert np.all((ku[1:] >> np.uint64(32))[same] > (ku[:-1] >> np.uint64(32))[same])  # row-major tiles inside one
    depth_bits = pre["depths"].view(np.uint32)[vu]
    assert np.array_equal((ku & np.uint64(0xFFFFFFFF)).astype(np.uint32), depth_bits)


@pytest.mark.parametrize("B,C,H,W,K", [(2, 6, 12, 9, 3), (1, 80, 16, 16, 7)])
def test_opacity_mask_oracle_matches_torch(B, C, H, W, K):
    rng = np.random.default_rng(B + C)
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    w = (rng.normal(size=(1, 2, K, K)) * 0.2).astype(np.float32)
    ob = rng.normal(size=(B, 1, H, W)).astype(np.float32)
    go = rng.normal(size=(B, C, H, W)).astype(np.float32)
    out, mask, stats = oracle.opacity_mask_forward(x, w, ob)
    xt, wt, ot = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w, ob))
    s = torch.cat([xt.mean(1, keepdim=True), xt.max(1, keepdim=True)[0]], 1)
    ref = xt * torch.sigmoid(torch.nn.functional.conv2d(s, wt, padding=K // 2) + ot)
    assert np.abs(out - ref.detach().numpy()).max() <= 1e-5
    (ref * torch.tensor(go, dtype=torch.float64)).sum().backward()
    gx, gw, gop = oracle.opacity_mask_backward(x, w, mask, stats, go)
    assert util.rel_err(gx, xt.grad.numpy()) <= 1e-5
    assert util.rel_err(gw, wt.grad.numpy().reshape(2, K, K)) <= 1e-5
    assert util.rel_err(gop, ot.grad.numpy()) <= 1e-5


@pytest.mark.parametrize("case", [dict(B=1, N=2, D=10, H=4, W=7, C=12, bev=16, seed=1),
                                  dict(B=2, N=3, D=16, H=6, W=10, C=19, bev=32, seed=2)])
def test_bev_pool_oracle_matches_dense_numpy(case):
    """ocrf_oracle_bev_pool_* (restating bev_pool_cuda.cu:21-121) against an order-free float64 formulation."""
    from ocrfdet_b200.scenes import bev_pool_case
    c = bev_pool_case(**case)
    C_ = case["C"]
    d, f = c["depth"].reshape(-1).astype(np.float64), c["feat"].reshape(-1, C_).astype(np.float64)
    out = oracle.bev_pool_forward(c["depth"], c["feat"], c["ranks_depth"], c["ranks_feat"], c["ranks_bev"], c["n_bev"],
                                  c["interval_starts"], c["interval_lengths"])
    want = np.zeros((c["n_bev"], C_))
    np.add.at(want, c["ranks_bev"], d[c["ranks_depth"]][:, None] * f[c["ranks_feat"]])
    assert np.abs(out - want).max() <= 1e-5 * (1 + np.abs(want).max())
    og = np.random.default_rng(5).normal(size=out.shape).astype(np.float32)
    dg, fg = oracle.bev_pool_backward(og, c["depth"], c["feat"], c["ranks_depth"], c["ranks_feat"], c["ranks_bev"])
    wdg = np.zeros(d.shape)
    wdg[c["ranks_depth"]] = (og[c["ranks_bev"]].astype(np.float64) * f[c["ranks_feat"]]).sum(1)
    wfg = np.zeros(f.shape)
    np.add.at(wfg, c["ranks_feat"], og[c["ranks_bev"]].astype(np.float64) * d[c["ranks_depth"]][:, None])
    assert np.abs(dg.reshape(-1) - wdg).max() <= 1e-5 * (1 + np.abs(wdg).max())
    assert np.abs(fg.reshape(-1, C_) - wfg).max() <= 1e-5 * (1 + np.abs(wfg).max())
    # intervals cover the point list exactly once, sorted by BEV cell (voxel_pooling_prepare_v2's contract)
    assert int(c["interval_lengths"].sum()) == len(c["ranks_bev"]) and np.all(np.diff(c["ranks_bev"]) >= 0)
