"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): tile keys / sort order / tile ranges bit-exact; forward outputs within
1e-5 abs/rel; gradients within 1e-4 rel.
"""
import numpy as np
import pytest
import torch

from ocrfdet_b200 import rasterizer as R
from oracle import oracle
from tests import util

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 1e-4


def _render_single(g, cam, bg, return_opacity=True, **kw):
    R.KEEP_STATE = True
    gc = util.to_cuda(g)
    for k in ("means3D", "scales", "rotations", "opacities", "colors"):
        if k in gc:
            gc[k].requires_grad_(True)
    st = util.settings_for(cam, bg, **kw)
    rast = R.GaussianRasterizer(st, return_opacity=return_opacity)
    means2D = torch.zeros_like(gc["means3D"], requires_grad=True)
    out = rast(means3D=gc["means3D"], means2D=means2D, opacities=gc["opacities"], colors_precomp=gc["colors"],
               scales=gc["scales"], rotations=gc["rotations"])
    return out, gc, means2D


SCENES = [
    ("frustum", dict(P=3000, seed=1, W=160, H=96)),
    ("frustum", dict(P=500, seed=2, W=50, H=37)),      # image not a multiple of the tile size
    ("ring", dict(P=20000, seed=3, W=352, H=128)),
    ("frustum", dict(P=10000, seed=4, W=704, H=256)),  # BASELINE config 1
]


@pytest.mark.parametrize("binning", ["split", "depthfirst", "pairsort"])
@pytest.mark.parametrize("kind,kw", SCENES)
def test_preprocess_and_binning_bit_exact(kind, kw, binning, monkeypatch):
    # all stage-2 algorithms (multi-split / depth-first / the reference's pair sort) must give the same bits
    monkeypatch.setenv("OCRF_BINNING", binning)
    g, cams = util.small_scene(kind, **kw)
    cam, W, H = cams[0], kw["W"], kw["H"]
    bg = [0.0, 0.0, 0.0]
    (_color, radii, _depth, _op), _, _ = _render_single(g, cam, bg)
    st = R.last_state(reference_lists=True)
    assert st["binning"] == binning
    want, wst = util.oracle_forward(g, cam, W, H, bg)
    pre, b = wst["pre"], wst["bin"]
    vis = pre["radii"] > 0
    assert np.array_equal(radii.cpu().numpy(), pre["radii"])
    assert np.array_equal(st["tiles_touched"][0].cpu().numpy().astype(np.uint32), pre["tiles_touched"])
    assert np.array_equal(st["offsets"].cpu().numpy().astype(np.uint32), b["offsets"])
    # float state: compare BIT PATTERNS on visible Gaussians (culled entries are never written)
    for name, ref in (("depths", pre["depths"]), ("xy", pre["xy"]), ("conic_opacity", pre["conic_opacity"])):
        got = st[name][0].cpu().numpy()
        assert np.array_equal(got[vis].view(np.uint32), ref[vis].view(np.uint32)), name
    assert st["num_pairs"] == b["N"]
    # the multi-split path does not materialise the pair lists: they come from the pair-sort kernels run on the
    # same geometry state (last_state(reference_lists=True)); the other two paths produce them themselves
    keys = st["keys"] if "keys" in st else st["keys_ref"]
    point_list = st["point_list"] if "point_list" in st else st["point_list_ref"]
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), b["keys"])
    assert np.array_equal(point_list.cpu().numpy().view(np.uint32), b["point_list"])
    assert np.array_equal(st["ranges"][0].cpu().numpy().view(np.uint32), b["ranges"])
    if binning != "pairsort":
        # vs the reference's pair sort on the same geometry state: identical bits everywhere
        for k in ("keys", "point_list", "ranges", "ranges_render"):
            if k in st:
                assert torch.equal(st[k], st[k + "_ref"]), k
        rr = st["ranges_render"][0].cpu().numpy()
        rec, ref = st["records"].cpu().numpy(), st["records_ref"].cpu().numpy()
        full = st["ranges"][0].cpu().numpy()
        for (lo, hi), (flo, fhi) in zip(rr, full):
            assert np.array_equal(rec[lo:hi], ref[lo:hi])
            # culled records are a sub-list of the tile's reference list: "orig" strictly increases and indexes it
            orig = rec[lo:hi, 6]
            assert np.all(np.diff(orig) > 0) and (hi == lo or (orig[0] >= 1 and orig[-1] <= fhi - flo))
            assert np.array_equal(rec[lo:hi, 10], b["point_list"][flo:fhi][orig - 1].astype(np.int32))


@pytest.mark.parametrize("P", [1, 31, 1023, 1025, 4097, 24576, 40000, 138000, 138200])  # 138 000 -> 130 950 visible (16 rows), 138 200 -> 131 120
def test_visible_sort_segment_sizes_and_depth_ties(P):
    # The depth sort of the visible Gaussians (visible_sort.cu) has two kernels: keys in registers up to 131 072 visible
    # Gaussians per view (R = ceil(n / 8192) rows of 32 per warp, 1 .. 16), keys through global memory above that.
    # Segment sizes on both sides of every boundary, depths quantised to half metres so that thousands of keys tie (ties
    # must stay in Gaussian-index order, rasterizer_impl.cu:303-308 sorts stably); lists and ranges bit-exact against
    # the oracle.
    W, H = 96, 64
    g, cams = util.small_scene("frustum", P=P, seed=11, W=W, H=H)
    g["means3D"][:, 2] = np.maximum(1.0, np.round(g["means3D"][:, 2] * 2.0) / 2.0)
    g["scales"] *= 0.3
    cam, bg = cams[0], [0.0, 0.0, 0.0]
    _render_single(g, cam, bg)
    st = R.last_state(reference_lists=True)
    _want, wst = util.oracle_forward(g, cam, W, H, bg)
    b = wst["bin"]
    assert st["num_pairs"] == b["N"]
    keys = st["keys"] if "keys" in st else st["keys_ref"]
    point_list = st["point_list"] if "point_list" in st else st["point_list_ref"]
    if "point_list_ref" in st and "point_list" in st:
        assert torch.equal(st["point_list"], st["point_list_ref"])
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), b["keys"])
    assert np.array_equal(point_list.cpu().numpy().view(np.uint32), b["point_list"])
    assert np.array_equal(st["ranges"][0].cpu().numpy().view(np.uint32), b["ranges"])
    # the culled records index the tile's reference list: this is what the multi-split derives from the depth order
    rr, full = st["ranges_render"][0].cpu().numpy(), st["ranges"][0].cpu().numpy()
    rec = st["records"].cpu().numpy()
    for (lo, hi), (flo, fhi) in zip(rr, full):
        orig = rec[lo:hi, 6]
        assert np.array_equal(rec[lo:hi, 10], b["point_list"][flo:fhi][orig - 1].astype(np.int32))


@pytest.mark.parametrize("kind,kw", SCENES)
def test_forward_outputs(kind, kw):
    g, cams = util.small_scene(kind, **kw)
    cam, W, H = cams[0], kw["W"], kw["H"]
    bg = [0.2, 0.5, 0.1]
    (color, _radii, depth, opac), _, _ = _render_single(g, cam, bg)
    want, _ = util.oracle_forward(g, cam, W, H, bg)
    amb = want["ambiguous"]
    util.assert_image_close(color.detach().cpu().numpy(), want["color"], amb, FWD_TOL, "color")
    util.assert_image_close(opac.detach().cpu().numpy(), want["opacity"], amb, FWD_TOL, "opacity")
    d = depth.cpu().numpy()
    assert np.array_equal(d[0][~amb.astype(bool)], want["depth"][0][~amb.astype(bool)]), "median depth"
    st = R.last_state()
    nc = st["n_contrib"][0].cpu().numpy()
    assert np.array_equal(nc[~amb.astype(bool)], want["n_contrib"].astype(np.int32)[~amb.astype(bool)])


@pytest.mark.parametrize("kind,kw", SCENES[:3])
def test_backward_gradients(kind, kw):
    g, cams = util.small_scene(kind, **kw)
    cam, W, H = cams[0], kw["W"], kw["H"]
    bg = [0.3, 0.1, 0.6]
    rng = np.random.default_rng(7)
    gcol = rng.normal(size=(3, H, W)).astype(np.float32)
    gop = rng.normal(size=(1, H, W)).astype(np.float32)
    (color, _radii, _depth, opac), gc, means2D = _render_single(g, cam, bg)
    loss = (color * torch.from_numpy(gcol).cuda()).sum() + (opac * torch.from_numpy(gop).cuda()).sum()
    loss.backward()
    want, wst = util.oracle_forward(g, cam, W, H, bg)
    # the oracle's backward must start from the SAME forward state as the GPU (n_contrib / final_T)
    gw = util.oracle_backward(g, cam, W, H, bg, want, wst, gcol, gop)
    pairs = [("means3D", gc["means3D"].grad, gw["means3D"]), ("scales", gc["scales"].grad, gw["scales"]),
             ("rotations", gc["rotations"].grad, gw["rotations"]),
             ("opacities", gc["opacities"].grad.reshape(-1), gw["opacities"]), ("colors", gc["colors"].grad, gw["colors"]),
             ("means2D", means2D.grad[:, :2], gw["means2D"])]
    for name, got, ref in pairs:
        e = util.rel_err(got.cpu().numpy(), ref)
        assert e <= GRAD_TOL, "%s gradient rel err %.3g" % (name, e)
    assert float(means2D.grad[:, 2].abs().max()) == 0.0


@pytest.mark.parametrize("binning", ["split", "depthfirst", "pairsort"])
def test_batch_matches_single_views(binning, monkeypatch):
    monkeypatch.setenv("OCRF_BINNING", binning)
    W, H = 352, 128
    g, cams = util.small_scene("ring", P=20000, seed=5, W=W, H=H, n_views=6)
    bg = [0.0, 0.0, 0.0]
    gc = util.to_cuda(g)
    u = lambda t: t.unsqueeze(0)  # noqa: E731
    color, radii, depth, opac = R.render_batch(u(gc["means3D"]), u(gc["opacities"]), util.cams_tensor(cams), H, W,
                                               torch.tensor(bg, device="cuda"), colors_precomp=u(gc["colors"]),
                                               scales=u(gc["scales"]), rotations=u(gc["rotations"]))
    for v, cam in enumerate(cams):
        (c1, r1, d1, o1), _, _ = _render_single(g, cam, bg)
        assert torch.equal(color[v], c1) and torch.equal(radii[v], r1) and torch.equal(depth[v], d1)
        assert torch.equal(opac[v], o1)


def test_batch_gradients_sum_over_views():
    W, H = 160, 96
    g, cams = util.small_scene("ring", P=6000, seed=6, W=W, H=H, n_views=3)
    bg = [0.1, 0.2, 0.3]
    rng = np.random.default_rng(3)
    gcol = torch.from_numpy(rng.normal(size=(3, 3, H, W)).astype(np.float32)).cuda()
    gc = util.to_cuda(g)
    names = ("means3D", "scales", "rotations", "opacities", "colors")
    for k in names:
        gc[k].requires_grad_(True)
    u = lambda t: t.unsqueeze(0)  # noqa: E731
    color, _, _, _ = R.render_batch(u(gc["means3D"]), u(gc["opacities"]), util.cams_tensor(cams), H, W,
                                    torch.tensor(bg, device="cuda"), colors_precomp=u(gc["colors"]),
                                    scales=u(gc["scales"]), rotations=u(gc["rotations"]))
    (color * gcol).sum().backward()
    got = {k: gc[k].grad.clone() for k in names}
    acc = {k: 0 for k in names}
    for v, cam in enumerate(cams):
        want, wst = util.oracle_forward(g, cam, W, H, bg)
        gw = util.oracle_backward(g, cam, W, H, bg, want, wst, gcol[v].cpu().numpy())
        for k in names:
            acc[k] = acc[k] + np.asarray(gw[k], np.float64)
    for k in names:
        e = util.rel_err(got[k].cpu().numpy().reshape(acc[k].shape), acc[k])
        assert e <= GRAD_TOL, "%s batched gradient rel err %.3g" % (k, e)


@pytest.mark.parametrize("C", [1, 8, 19, 40, 80])
def test_generic_channel_count(C):
    W, H = 96, 64
    g, cams = util.small_scene("frustum", P=1500, seed=8, W=W, H=H, channels=C)
    cam = cams[0]
    bg = list(np.linspace(0.1, 0.9, C).astype(np.float32))
    rng = np.random.default_rng(5)
    gcol = rng.normal(size=(C, H, W)).astype(np.float32)
    (color, _r, depth, opac), gc, means2D = _render_single(g, cam, bg)
    want, wst = util.oracle_forward(g, cam, W, H, bg)
    util.assert_image_close(color.detach().cpu().numpy(), want["color"], want["ambiguous"], FWD_TOL, "features")
    (color * torch.from_numpy(gcol).cuda()).sum().backward()
    gw = util.oracle_backward(g, cam, W, H, bg, want, wst, gcol)
    for name, got, ref in (("colors", gc["colors"].grad, gw["colors"]), ("means3D", gc["means3D"].grad, gw["means3D"]),
                           ("scales", gc["scales"].grad, gw["scales"]),
                           ("opacities", gc["opacities"].grad.reshape(-1), gw["opacities"])):
        e = util.rel_err(got.cpu().numpy(), ref)
        assert e <= GRAD_TOL, "%s gradient rel err %.3g (C=%d)" % (name, e, C)


def test_sh_and_cov3d_precomp_paths():
    W, H = 128, 80
    g, cams = util.small_scene("frustum", P=1200, seed=9, W=W, H=H)
    cam = cams[0]
    bg = [0.0, 0.1, 0.2]
    rng = np.random.default_rng(11)
    P = g["means3D"].shape[0]
    shs = (rng.normal(size=(P, 16, 3)) * 0.3).astype(np.float32)
    cov = util.cov3d_numpy(g["scales"], g["rotations"])
    gcol = rng.normal(size=(3, H, W)).astype(np.float32)
    for deg in (0, 2, 3):
        R.KEEP_STATE = True
        st = util.settings_for(cam, bg, sh_degree=deg)
        m = torch.from_numpy(g["means3D"]).cuda().requires_grad_(True)
        sh_t = torch.from_numpy(shs).cuda().requires_grad_(True)
        cov_t = torch.from_numpy(cov).cuda().requires_grad_(True)
        op_t = torch.from_numpy(g["opacities"]).cuda().requires_grad_(True)
        color, radii, depth = R.GaussianRasterizer(st)(means3D=m, means2D=torch.zeros_like(m), opacities=op_t, shs=sh_t,
                                                       cov3D_precomp=cov_t)
        want, wst = oracle.rasterize(g["means3D"], g["opacities"], None, cam["viewmatrix"], cam["projmatrix"], W, H,
                                     cam["tanfovx"], cam["tanfovy"], np.asarray(bg, np.float32), cov3D_precomp=cov,
                                     shs=shs, sh_degree=deg, campos=cam["campos"])
        util.assert_image_close(color.detach().cpu().numpy(), want["color"], want["ambiguous"], FWD_TOL, "SH colour")
        (color * torch.from_numpy(gcol).cuda()).sum().backward()
        gw = oracle.rasterize_backward(wst, g["means3D"], cam["viewmatrix"], cam["projmatrix"], W, H, cam["tanfovx"],
                                       cam["tanfovy"], np.asarray(bg, np.float32), want, gcol, shs=shs, sh_degree=deg,
                                       campos=cam["campos"])
        for name, got, ref in (("shs", sh_t.grad, gw["shs"]), ("cov3D", cov_t.grad, gw["cov3D"]),
                               ("means3D", m.grad, gw["means3D"]), ("opacities", op_t.grad.reshape(-1), gw["opacities"])):
            e = util.rel_err(got.cpu().numpy(), ref)
            assert e <= GRAD_TOL, "%s gradient rel err %.3g (deg %d)" % (name, e, deg)


def test_edge_cases():
    W, H = 64, 48
    g, cams = util.small_scene("frustum", P=300, seed=10, W=W, H=H)
    cam = cams[0]
    bg = [0.25, 0.5, 0.75]
    st = util.settings_for(cam, bg)
    rast = R.GaussianRasterizer(st, return_opacity=True)
    # (1) nothing visible: every Gaussian behind the camera -> pure background, zero gradients
    gb = dict(g)
    gb["means3D"] = g["means3D"] * np.array([1, 1, -1], np.float32)
    gc = util.to_cuda(gb)
    gc["colors"].requires_grad_(True)
    color, radii, depth, opac = rast(means3D=gc["means3D"], means2D=torch.zeros_like(gc["means3D"]),
                                     opacities=gc["opacities"], colors_precomp=gc["colors"], scales=gc["scales"],
                                     rotations=gc["rotations"])
    assert int(radii.abs().sum()) == 0
    want = torch.tensor(bg, device="cuda").view(3, 1, 1).expand(3, H, W)
    assert torch.equal(color, want) and float(opac.abs().max()) == 0.0 and float((depth - 15).abs().max()) == 0.0
    color.sum().backward()
    assert float(gc["colors"].grad.abs().max()) == 0.0
    # (2) P == 0: zero image as the reference binding (rasterize_points.cu:68,81)
    e = torch.zeros((0, 3), device="cuda")
    color, radii, depth, opac = rast(means3D=e, means2D=e, opacities=torch.zeros((0, 1), device="cuda"), colors_precomp=e,
                                     scales=e, rotations=torch.zeros((0, 4), device="cuda"))
    assert color.shape == (3, H, W) and float(color.abs().max()) == 0.0 and radii.numel() == 0
    # (3) argument errors of the reference wrapper
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(means3D=gc["means3D"], means2D=None, opacities=gc["opacities"], scales=gc["scales"], rotations=gc["rotations"])
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        rast(means3D=gc["means3D"], means2D=None, opacities=gc["opacities"], colors_precomp=gc["colors"])
    # (4) markVisible
    vis = rast.markVisible(torch.from_numpy(g["means3D"]).cuda())
    assert np.array_equal(vis.cpu().numpy(), oracle.mark_visible(g["means3D"], cam["viewmatrix"], cam["projmatrix"]))


def test_capacity_mode_no_sync_and_overflow():
    W, H = 160, 96
    g, cams = util.small_scene("frustum", P=3000, seed=12, W=W, H=H)
    bg = torch.zeros(3, device="cuda")
    gc = util.to_cuda(g)
    u = lambda t: t.unsqueeze(0)  # noqa: E731
    args = (u(gc["means3D"]), u(gc["opacities"]), util.cams_tensor(cams), H, W, bg)
    kw = dict(colors_precomp=u(gc["colors"]), scales=u(gc["scales"]), rotations=u(gc["rotations"]))
    R.KEEP_STATE = True
    exact = R.render_batch(*args, **kw)
    n = R.last_state()["num_pairs"]
    roomy = R.render_batch(*args, pair_capacity=2 * n + 1000, **kw)
    R.check_overflow()
    assert torch.equal(roomy[0], exact[0]) and torch.equal(roomy[2], exact[2])
    tight = R.render_batch(*args, pair_capacity=n // 2, **kw)
    with pytest.raises(Exception, match="overflow"):
        R.check_overflow()
    assert float(tight[0].abs().max()) == 0.0  # rendered background only, no overrun


def test_clear_gradients_touches_only_what_backward_accumulates_into():
    """ocrf_clear_gradients == the reference's torch::zeros for everything the backward reads: rows of visible pairs
    and the colour gradients are zeroed, rows of invisible pairs (never read nor written downstream) are left alone."""
    import ctypes as C
    from ocrfdet_b200 import _lib
    L = _lib.lib()
    V, P, S, Cc = 3, 1000, 1, 3
    shape = _lib.OcrfShape(S, P, V, V, 64, 48, Cc, 0, 0)
    rng = np.random.default_rng(0)
    radii = torch.from_numpy((rng.random((V, P)) < 0.3).astype(np.int32) * 5).cuda()
    ggrad = torch.full((V, P, 6), float("nan"), dtype=torch.float64, device="cuda")
    gcol = torch.full((S, P, Cc), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(L.ocrf_clear_gradients(_lib.current_stream(), C.byref(shape), 0, _lib.ptr(radii), _lib.ptr(ggrad),
                                      _lib.ptr(gcol)), "ocrf_clear_gradients")
    vis = radii > 0
    assert bool((ggrad[vis] == 0).all()) and bool(torch.isnan(ggrad[~vis]).all())
    assert bool((gcol == 0).all())


def test_colors_ready_event_late_binds_the_colour_upload():
    """render_batch(colors_ready=event): colours uploaded on a second stream while the preprocess runs."""
    W, H, P = 176, 64, 5000
    g, cams = util.small_scene("ring", P=P, seed=61, W=W, H=H, n_views=3)
    cam_t = util.cams_tensor(cams)
    gc = util.to_cuda(g)
    bg = torch.zeros(3, device="cuda")
    kw = dict(scales=gc["scales"].unsqueeze(0), rotations=gc["rotations"].unsqueeze(0))
    want = R.render_batch(gc["means3D"].unsqueeze(0), gc["opacities"].unsqueeze(0), cam_t, H, W, bg,
                          colors_precomp=gc["colors"].unsqueeze(0), **kw)
    host_col = torch.from_numpy(g["colors"]).unsqueeze(0).pin_memory()
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        torch.cuda._sleep(20_000_000)          # the upload is still in flight when render_batch is called
        col = host_col.to("cuda", non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(side)
    col.record_stream(torch.cuda.current_stream())
    got = R.render_batch(gc["means3D"].unsqueeze(0), gc["opacities"].unsqueeze(0), cam_t, H, W, bg, colors_precomp=col,
                         colors_ready=ready, **kw)
    for a, b in zip(want, got):
        assert torch.equal(a, b)


@pytest.mark.parametrize("C,half", [(80, 0), (48, 0), (36, 0), (80, 1), (40, 1)])
def test_tensor_core_blend_odd_image_batch_and_half_tiles(C, half):
    """The tcgen05 blend kernels (32 < C <= 80, C % 4 == 0) on an image that is not a multiple of the tile, two samples
    x two views in one call, channel counts that pad the MMA's N (36 -> 48), whole-tile and half-tile CTAs; a
    subprocess per case because the CTA shape is read from the environment once."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np, torch
from tests import util
from ocrfdet_b200 import rasterizer as R
C = %d
W, H, V = 90, 50, 2
gs = [util.small_scene("ring", P=1500, seed=20 + s, W=W, H=H, channels=C, n_views=V) for s in range(2)]
cams = gs[0][1] + gs[1][1]
names = ("means3D", "scales", "rotations", "opacities", "colors")
t = {k: torch.from_numpy(np.stack([g[0][k] for g in gs])).cuda().requires_grad_(True) for k in names}
bg = np.linspace(0.1, 0.9, C).astype(np.float32)
rng = np.random.default_rng(9)
gcol = rng.normal(size=(2 * V, C, H, W)).astype(np.float32)
gop = rng.normal(size=(2 * V, 1, H, W)).astype(np.float32)
outs = [util.oracle_forward(gs[i // V][0], cams[i], W, H, list(bg)) for i in range(2 * V)]
amb = np.stack([np.asarray(w["ambiguous"]).astype(bool).reshape(H, W) for w, _ in outs])
gcol, gop = gcol * ~amb[:, None], gop * ~amb[:, None]   # a borderline pixel may blend one Gaussian more or less
color, radii, depth, opac = R.render_batch(t["means3D"], t["opacities"], util.cams_tensor(cams), H, W,
                                           torch.from_numpy(bg).cuda(), colors_precomp=t["colors"], scales=t["scales"],
                                           rotations=t["rotations"])
torch.autograd.backward([color, opac], [torch.from_numpy(gcol).cuda(), torch.from_numpy(gop).cuda()])
for s in range(2):
    acc = {k: 0.0 for k in names}
    for v in range(V):
        i = s * V + v
        want, wst = outs[i]
        util.assert_image_close(color[i].detach().cpu().numpy(), want["color"], want["ambiguous"], 1e-5, "features")
        util.assert_image_close(opac[i].detach().cpu().numpy(), want["opacity"], want["ambiguous"], 1e-5, "opacity")
        gw = util.oracle_backward(gs[s][0], cams[i], W, H, list(bg), want, wst, gcol[i], gop[i])
        for k in names:
            acc[k] = acc[k] + np.asarray(gw[k], np.float64).reshape(gs[s][0][k].shape)
    for k in names:
        e = util.rel_err(t[k].grad[s].cpu().numpy(), acc[k])
        assert e <= 1e-4, "%%s gradient of sample %%d: rel err %%.3g (C=%%d)" %% (k, s, e, C)
print("PARITY_OK")
''' % C
    env = dict(os.environ, OCRF_TC_FWD_MB="1" if half else "2", OCRF_TC_BWD_MB="1" if half else "2",
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "PARITY_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_foreground_filter_changes_pair_counts_not_results():
    """render_batch(min_opacity=...): Gaussians below the blend threshold are culled before projection;
    images and gradients are the unfiltered ones, radii of the culled Gaussians are 0 and the pair count shrinks."""
    W, H, V = 176, 64, 3
    g, cams = util.small_scene("ring", P=6000, seed=51, W=W, H=H, n_views=V)
    rng = np.random.default_rng(2)
    g["opacities"] = np.where(rng.random(g["opacities"].shape) < 0.6, np.float32(0.003) * rng.random(g["opacities"].shape),
                              g["opacities"]).astype(np.float32)   # 60 % "empty space" voxels
    names = ("means3D", "scales", "rotations", "opacities", "colors")
    gcol = torch.randn(V, 3, H, W, device="cuda")
    gop = torch.randn(V, 1, H, W, device="cuda")
    res = []
    for thr in (0.0, 1.0):  # (values above 1/255 are clamped to it)
        t = {k: torch.from_numpy(g[k]).unsqueeze(0).cuda().requires_grad_(True) for k in names}
        R.KEEP_STATE = True
        color, radii, depth, opac = R.render_batch(t["means3D"], t["opacities"], util.cams_tensor(cams), H, W,
                                                   torch.zeros(3, device="cuda"), colors_precomp=t["colors"],
                                                   scales=t["scales"], rotations=t["rotations"], min_opacity=thr)
        n_pairs = int(R.last_state()["num_pairs"])
        R.KEEP_STATE = False
        torch.autograd.backward([color, opac], [gcol, gop])
        res.append((color.detach(), depth, opac.detach(), radii, n_pairs, {k: t[k].grad for k in names}))
    (c0, d0, o0, r0, n0, g0), (c1, d1, o1, r1, n1, g1) = res
    assert torch.equal(c0, c1) and torch.equal(d0, d1) and torch.equal(o0, o1)
    low = torch.from_numpy(g["opacities"].reshape(-1) < 1.0 / 255.0).cuda()
    assert int(r1[:, low].abs().max()) == 0 and torch.equal(r0[:, ~low], r1[:, ~low])
    assert n1 < 0.6 * n0, (n0, n1)
    for k in names:
        assert util.rel_err(g1[k].cpu().numpy(), g0[k].cpu().numpy()) <= 1e-6, k
    assert float(g0["opacities"].reshape(-1)[low].abs().max()) == 0.0   # they never blended
