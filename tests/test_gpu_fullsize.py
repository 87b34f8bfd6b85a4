"""Parity at the sizes that are BENCHMARKED: the exact bench.py workload (BASELINE config 2) in
exact and capacity mode, config 4 (C = 80) at full size, and a 512x1408 view (2 816 tiles, config 5's image), each
against the CPU oracle and -- where it has the output -- against the reference's own CUDA rasterizer
(oracle/_ref/libinria_ref.so, the vendored Inria kernels compiled unmodified).

Bars: integer / index state bit-exact; forward 1e-5 abs/rel on every pixel whose blend decisions were not borderline,
and a HARD cap of 5e-3 on the borderline ones (a flipped blend changes a pixel by at most alpha*T*|c|); gradients
1e-4 of the largest element AND per element |a-b| <= 1e-4 |b| + 1e-6 max|b| on colours / opacities.  Upstream
gradients are zeroed on the borderline pixels (they are < 0.1 % of the image), so that a flipped branch -- which both
implementations are entitled to -- does not enter any gradient sum.
"""
import numpy as np
import pytest
import torch

from ocrfdet_b200 import rasterizer as R
from ocrfdet_b200.scenes import gaussians_on_grid, ring_scene
from ocrfdet_b200.cameras import ego_ring_cameras
from oracle import ref
from tests import util

pytestmark = pytest.mark.gpu

NAMES = ("means3D", "scales", "rotations", "opacities", "colors")
BORDERLINE_CAP = 5e-3


def _leafs(g):
    gc = util.to_cuda(g)
    return {k: gc[k].unsqueeze(0).requires_grad_(True) for k in NAMES}


def _render(t, cam_t, H, W, bg, **kw):
    return R.render_batch(t["means3D"], t["opacities"], cam_t, H, W, bg, colors_precomp=t["colors"], scales=t["scales"],
                          rotations=t["rotations"], **kw)


def assert_images(got, want, amb, what):
    """1e-5 off the borderline pixels, <= 5e-3 on them, and few of them."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    amb = np.asarray(amb).astype(bool)
    err = np.abs(got - want) / (1.0 + np.abs(want))
    clean = err[..., ~amb]
    assert clean.size == 0 or clean.max() <= 1e-5, "%s: max err %.3g on unambiguous pixels" % (what, clean.max())
    assert err.max() <= BORDERLINE_CAP, "%s: a borderline pixel is off by %.3g > %.0e" % (what, err.max(), BORDERLINE_CAP)
    assert amb.mean() < 2e-3, "%s: too many borderline pixels (%.3f%%)" % (what, 100 * amb.mean())


def assert_grad(name, got, want, elementwise=False, tol=1e-4):
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    scale = np.abs(want).max() + 1e-30
    e = np.abs(got - want).max() / scale
    assert e <= tol, "%s gradient: max-norm rel err %.3g" % (name, e)
    if elementwise:
        bad = np.abs(got - want) > 1e-4 * np.abs(want) + 1e-6 * scale
        assert not bad.any(), "%s gradient: %d of %d elements beyond 1e-4|b| + 1e-6 max|b| (worst %.3g rel)" % (
            name, int(bad.sum()), bad.size, float((np.abs(got - want) / (np.abs(want) + 1e-6 * scale)).max()))


@pytest.fixture(scope="module")
def bench_scene():
    """bench.py's workload on rank 0: ring_scene(P = 100 000, seed 1234), 6 views 256x704, seeded upstream gradients."""
    W, H, P, V = 704, 256, 100_000, 6
    g, cams = ring_scene(P=P, seed=1234, width=W, height=H, channels=3, n_views=V)
    rng = np.random.default_rng(99)
    gcol = rng.normal(size=(V, 3, H, W)).astype(np.float32)
    gop = rng.normal(size=(V, 1, H, W)).astype(np.float32)
    bg = np.zeros(3, np.float32)
    oracle_out = []
    for v in range(V):
        want, wst = util.oracle_forward(g, cams[v], W, H, bg)
        oracle_out.append((want, wst))
    amb = np.stack([np.asarray(w["ambiguous"]).astype(bool).reshape(H, W) for w, _ in oracle_out])
    gcol = gcol * ~amb[:, None]
    gop = gop * ~amb[:, None]
    grads = {k: 0.0 for k in NAMES}
    for v in range(V):
        want, wst = oracle_out[v]
        gw = util.oracle_backward(g, cams[v], W, H, bg, want, wst, gcol[v], gop[v])
        for k in NAMES:
            grads[k] = grads[k] + np.asarray(gw[k], np.float64).reshape(g[k].shape)
    return dict(W=W, H=H, P=P, V=V, g=g, cams=cams, gcol=gcol, gop=gop, bg=bg, oracle=oracle_out, amb=amb, grads=grads)


@pytest.mark.parametrize("mode", ["exact", "capacity"])
def test_bench_workload_against_oracle(bench_scene, mode):
    s = bench_scene
    W, H, V = s["W"], s["H"], s["V"]
    t = _leafs(s["g"])
    cam_t = util.cams_tensor(s["cams"])
    bg = torch.zeros(3, device="cuda")
    kw = {}
    if mode == "capacity":  # what bench.py times: capacity = 1.3 x the previous step's pair count, no host sync
        R.KEEP_STATE = True
        with torch.no_grad():
            _render(t, cam_t, H, W, bg)
        kw["pair_capacity"] = int(R.last_state()["num_pairs"] * 1.3) + 4096
        R.KEEP_STATE = False
    color, radii, depth, opac = _render(t, cam_t, H, W, bg, **kw)
    torch.autograd.backward([color, opac], [torch.from_numpy(s["gcol"]).cuda(), torch.from_numpy(s["gop"]).cuda()])
    R.check_overflow()
    for v in range(V):
        want, wst = s["oracle"][v]
        amb = s["amb"][v]
        assert np.array_equal(radii[v].cpu().numpy(), wst["pre"]["radii"]), "radii of view %d" % v
        assert_images(color[v].detach().cpu().numpy(), want["color"], amb, "colour of view %d" % v)
        assert_images(opac[v].detach().cpu().numpy(), want["opacity"], amb, "opacity of view %d" % v)
        d = depth[v, 0].cpu().numpy()
        assert np.array_equal(d[~amb], np.asarray(want["depth"])[0][~amb]), "median depth of view %d" % v
    for k in NAMES:
        assert_grad(k, t[k].grad.cpu().numpy(), s["grads"][k], elementwise=k in ("colors", "opacities"))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libinria_ref.so not built")
def test_bench_workload_against_reference_cuda(bench_scene):
    """Every view of the benched batch against the reference's own kernels: radii, pair count, sorted keys, point list,
    tile ranges bit-exact; colour; accumulated opacity 1 - final_T and n_contrib (what the reference keeps in its image
    state); colour-driven gradients."""
    s = bench_scene
    W, H, V, P = s["W"], s["H"], s["V"], s["P"]
    t = _leafs(s["g"])
    cam_t = util.cams_tensor(s["cams"])
    bg3 = [0.0, 0.0, 0.0]
    R.KEEP_STATE = True
    color, radii, depth, opac = _render(t, cam_t, H, W, torch.zeros(3, device="cuda"))
    ms = R.last_state(reference_lists=True)
    R.KEEP_STATE = False
    gcol = torch.from_numpy(s["gcol"]).cuda()
    torch.autograd.backward([color], [gcol])
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    keys = (ms["keys"] if "keys" in ms else ms["keys_ref"]).cpu().numpy().view(np.uint64)
    plist = (ms["point_list"] if "point_list" in ms else ms["point_list_ref"]).cpu().numpy()
    ranges = ms["ranges"].cpu().numpy()
    gc = util.to_cuda(s["g"])
    rr = ref.RefRasterizer()
    acc = {k: 0.0 for k in ("opacities", "colors", "means3D", "scales", "rotations")}
    first = 0
    for v in range(V):
        st = util.settings_for(s["cams"][v], bg3)
        rcol, rradii, rN = rr.forward(gc["means3D"], gc["opacities"], gc["colors"], st.viewmatrix, st.projmatrix,
                                      st.campos, W, H, st.tanfovx, st.tanfovy, st.bg, scales=gc["scales"],
                                      rotations=gc["rotations"])
        rs = rr.state()
        assert torch.equal(radii[v], rradii), "radii of view %d" % v
        # our batch key = (view * tiles + tile) << 32 | depth: the view's slice, rebased, is the reference's list
        rk = rs["keys"].cpu().numpy().view(np.uint64)
        mine = keys[first:first + rN]
        assert mine.shape == rk.shape and np.array_equal(mine - (np.uint64(v * tiles) << np.uint64(32)), rk), \
            "sorted (tile|depth) keys of view %d" % v
        assert np.array_equal(plist[first:first + rN], rs["point_list"].cpu().numpy()), "sort order of view %d" % v
        rr_ranges = rs["ranges"].cpu().numpy().astype(np.int64)
        mr = ranges[v].astype(np.int64)
        nonempty = rr_ranges[:, 1] > rr_ranges[:, 0]
        assert np.array_equal(mr[nonempty] - first, rr_ranges[nonempty]) and not mr[~nonempty].any(), \
            "tile ranges of view %d" % v
        first += rN
        amb = s["amb"][v]
        assert_images(color[v].detach().cpu().numpy(), rcol.cpu().numpy(), amb, "colour vs reference, view %d" % v)
        assert_images(opac[v, 0].detach().cpu().numpy(), 1.0 - rs["final_T"].cpu().numpy(), amb,
                      "opacity vs 1 - final_T of the reference, view %d" % v)
        nc, rnc = ms["n_contrib"][v].cpu().numpy(), rs["n_contrib"].cpu().numpy()
        assert np.array_equal(nc[~amb], rnc[~amb]), "n_contrib of view %d" % v
        gr = rr.backward(gc["means3D"], gc["colors"], st.viewmatrix, st.projmatrix, st.campos, st.tanfovx, st.tanfovy,
                         st.bg, rradii, gcol[v], scales=gc["scales"], rotations=gc["rotations"])
        for k in acc:
            acc[k] = acc[k] + gr[k].double().cpu().numpy().reshape(s["g"][k].shape)
    assert first == ms["num_pairs"]
    rr.close()
    # colour-only upstream gradient: the oracle's sum for the same loss is the float64 truth the chain is judged by
    truth = {k: 0.0 for k in acc}
    for v in range(V):
        want, wst = s["oracle"][v]
        gw = util.oracle_backward(s["g"], s["cams"][v], W, H, s["bg"], want, wst, s["gcol"][v])
        for k in acc:
            truth[k] = truth[k] + np.asarray(gw[k], np.float64).reshape(s["g"][k].shape)
    for k in ("opacities", "colors"):
        assert_grad(k + " vs reference CUDA", t[k].grad.cpu().numpy(), acc[k], elementwise=False)
        assert_grad(k + " vs oracle", t[k].grad.cpu().numpy(), truth[k], elementwise=True)
    for k in ("means3D", "scales", "rotations"):
        e_mine, e_ref = util.rel_err(t[k].grad.cpu().numpy(), truth[k]), util.rel_err(acc[k], truth[k])
        e_pair = util.rel_err(t[k].grad.cpu().numpy(), acc[k])
        assert e_mine <= 1e-4, "%s vs float64 oracle: %.3g (reference: %.3g)" % (k, e_mine, e_ref)
        assert e_pair <= 1e-4 + e_ref, "%s vs reference CUDA: %.3g > 1e-4 + %.3g" % (k, e_pair, e_ref)


def test_config4_feature_rendering_full_size():
    """BASELINE config 4 at its stated size, two of its views: 80 feature channels, 100 000 Gaussians, 256x704,
    features + depth + opacity, forward and backward, against the oracle."""
    W, H, P, C, V = 704, 256, 100_000, 80, 2
    g, cams = ring_scene(P=P, seed=404, width=W, height=H, channels=C, n_views=6)
    cams = [cams[0], cams[3]]
    bgv = np.linspace(0.0, 0.5, C).astype(np.float32)
    rng = np.random.default_rng(44)
    gcol = rng.normal(size=(V, C, H, W)).astype(np.float32)
    gop = rng.normal(size=(V, 1, H, W)).astype(np.float32)
    t = _leafs(g)
    color, radii, depth, opac = _render(t, util.cams_tensor(cams), H, W, torch.from_numpy(bgv).cuda())
    outs = [util.oracle_forward(g, cams[v], W, H, bgv) for v in range(V)]
    amb = np.stack([np.asarray(w["ambiguous"]).astype(bool).reshape(H, W) for w, _ in outs])
    gcol, gop = gcol * ~amb[:, None], gop * ~amb[:, None]
    torch.autograd.backward([color, opac], [torch.from_numpy(gcol).cuda(), torch.from_numpy(gop).cuda()])
    truth = {k: 0.0 for k in NAMES}
    for v in range(V):
        want, wst = outs[v]
        assert np.array_equal(radii[v].cpu().numpy(), wst["pre"]["radii"])
        assert_images(color[v].detach().cpu().numpy(), want["color"], amb[v], "features of view %d" % v)
        assert_images(opac[v].detach().cpu().numpy(), want["opacity"], amb[v], "opacity of view %d" % v)
        assert np.array_equal(depth[v, 0].cpu().numpy()[~amb[v]], np.asarray(want["depth"])[0][~amb[v]])
        gw = util.oracle_backward(g, cams[v], W, H, bgv, want, wst, gcol[v], gop[v])
        for k in NAMES:
            truth[k] = truth[k] + np.asarray(gw[k], np.float64).reshape(g[k].shape)
    for k in NAMES:
        assert_grad(k, t[k].grad.cpu().numpy(), truth[k], elementwise=k in ("colors", "opacities"))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libinria_ref.so not built")
@pytest.mark.parametrize("binning", ["split", "pairsort"])
def test_512x1408_view_against_reference_cuda(binning, monkeypatch):
    """One view of config 5's image (1408x512 = 88 x 32 = 2 816 tiles, 44-bit keys) with 250 000 Gaussians from its
    refined 277^2 x 13 voxel grid: keys / point list / ranges bit-exact, colour, opacity, n_contrib, gradients."""
    monkeypatch.setenv("OCRF_BINNING", binning)
    W, H, P = 1408, 512, 250_000
    g = gaussians_on_grid(P, seed=55, channels=3, bev=277)
    cam = ego_ring_cameras(W, H)[1]
    bg = [0.1, 0.2, 0.3]
    gc = util.to_cuda(g)
    st = util.settings_for(cam, bg)
    rr = ref.RefRasterizer()
    rcol, rradii, rN = rr.forward(gc["means3D"], gc["opacities"], gc["colors"], st.viewmatrix, st.projmatrix, st.campos,
                                  W, H, st.tanfovx, st.tanfovy, st.bg, scales=gc["scales"], rotations=gc["rotations"])
    rs = rr.state()
    assert rN > 2_000_000
    R.KEEP_STATE = True
    leaf = {k: gc[k].clone().requires_grad_(True) for k in NAMES}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
    color, radii, depth, opac = R.GaussianRasterizer(st, return_opacity=True)(
        means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], colors_precomp=leaf["colors"],
        scales=leaf["scales"], rotations=leaf["rotations"])
    ms = R.last_state(reference_lists=True)
    R.KEEP_STATE = False
    assert ms["binning"] == binning and ms["num_pairs"] == rN
    assert torch.equal(radii, rradii)
    assert torch.equal(ms["keys"] if "keys" in ms else ms["keys_ref"], rs["keys"]), "sorted (tile|depth) keys"
    assert torch.equal(ms["point_list"] if "point_list" in ms else ms["point_list_ref"], rs["point_list"]), "sort order"
    assert torch.equal(ms["ranges"][0], rs["ranges"]), "tile ranges"
    want, wst = util.oracle_forward(g, cam, W, H, bg)
    amb = np.asarray(want["ambiguous"]).astype(bool).reshape(H, W)
    assert_images(color.detach().cpu().numpy(), rcol.cpu().numpy(), amb, "colour vs reference")
    assert_images(color.detach().cpu().numpy(), want["color"], amb, "colour vs oracle")
    assert_images(opac[0].detach().cpu().numpy(), 1.0 - rs["final_T"].cpu().numpy(), amb, "opacity vs 1 - final_T")
    assert np.array_equal(ms["n_contrib"][0].cpu().numpy()[~amb], rs["n_contrib"].cpu().numpy()[~amb]), "n_contrib"
    rng = np.random.default_rng(5)
    gcol_np = rng.normal(size=(3, H, W)).astype(np.float32) * ~amb[None]
    gcol = torch.from_numpy(gcol_np).cuda()
    color.backward(gcol)
    gr = rr.backward(gc["means3D"], gc["colors"], st.viewmatrix, st.projmatrix, st.campos, st.tanfovx, st.tanfovy, st.bg,
                     rradii, gcol, scales=gc["scales"], rotations=gc["rotations"])
    truth = util.oracle_backward(g, cam, W, H, bg, want, wst, gcol_np)
    for k in ("opacities", "colors"):
        assert_grad(k + " vs reference CUDA", leaf[k].grad.cpu().numpy(), gr[k].cpu().numpy())
        assert_grad(k + " vs oracle", leaf[k].grad.cpu().numpy(), truth[k], elementwise=True)
    assert_grad("means2D vs reference CUDA", means2D.grad.cpu().numpy(), gr["means2D"].cpu().numpy())
    for k in ("means3D", "scales", "rotations"):
        e_mine, e_ref = util.rel_err(leaf[k].grad.cpu().numpy(), truth[k]), util.rel_err(gr[k].cpu().numpy(), truth[k])
        assert e_mine <= 1e-4, "%s vs float64 oracle: %.3g (reference: %.3g)" % (k, e_mine, e_ref)
    rr.close()


def test_capacity_overflow_with_one_view_touches_nothing_outside_its_workspace():
    """with V = 1 and a capacity far below the real pair count the chunk scan of the multi-split
    used to index its tables by the REAL pair count.  Drive the C ABI directly with the binning workspace embedded in a
    larger buffer whose surroundings hold a pattern, overflow it, and require the pattern intact."""
    import ctypes as C
    from ocrfdet_b200 import _lib
    L = _lib.lib()
    W, H, P = 704, 256, 60_000
    g, cams = ring_scene(P=P, seed=77, width=W, height=H, n_views=1)
    gc = util.to_cuda(g)
    cam_t = util.cams_tensor(cams)
    shape = _lib.OcrfShape(1, P, 1, 1, W, H, 3, 0, 0)
    geo, img, binl = _lib.OcrfGeomLayout(), _lib.OcrfImageLayout(), _lib.OcrfBinLayout()
    _lib.check(L.ocrf_geom_layout(C.byref(shape), 0, C.byref(geo)), "geom layout")
    _lib.check(L.ocrf_image_layout(C.byref(shape), C.byref(img)), "image layout")
    cap = 5000  # the scene has several hundred thousand pairs
    _lib.check(L.ocrf_bin_layout(C.byref(shape), C.c_uint64(cap), C.byref(binl)), "bin layout")
    guard = 8 << 20
    arena = torch.full((guard + binl.total + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    binning = arena[guard:guard + binl.total]
    geom = torch.empty(geo.total, dtype=torch.uint8, device="cuda")
    image = torch.empty(img.total, dtype=torch.uint8, device="cuda")
    radii = torch.empty((1, P), dtype=torch.int32, device="cuda")
    sticky = torch.zeros(2, dtype=torch.int32, device="cuda")
    st, p = _lib.current_stream(), _lib.ptr
    _lib.check(L.ocrf_preprocess_forward(st, C.byref(shape), p(gc["means3D"]), p(gc["scales"]), p(gc["rotations"]), None,
                                         p(gc["opacities"]), None, p(cam_t), C.c_float(1.0), 0, p(radii), p(geom)),
               "preprocess")
    for flags in (0, _lib.OCRF_BIN_DEPTH_FIRST, _lib.OCRF_BIN_PAIR_SORT):
        _lib.check(L.ocrf_bin_forward(st, C.byref(shape), C.c_uint64(cap), p(radii), p(gc["colors"]), 0,
                                      C.c_uint32(flags), p(geom), p(binning), p(image), p(sticky)), "bin")
        torch.cuda.synchronize()
        assert bool((arena[:guard] == 0xA5).all()) and bool((arena[guard + binl.total:] == 0xA5).all()), \
            "binning flags %d wrote outside its workspace on overflow" % flags
        assert int(sticky[0]) & 1 and int(sticky[1]) > 100_000
    ranges = image[img.ranges_render:img.ranges_render + 8 * 704].view(torch.int32)
    assert int(ranges.abs().max()) == 0  # nothing is rendered


def test_overflow_is_reported_without_asking(monkeypatch):
    """Capacity mode: an overflowing call renders background; the NEXT entry into the path raises (sync-free, through
    the pinned mirror of the sticky status words), and the state is clean again afterwards."""
    from ocrfdet_b200 import _lib
    W, H = 160, 96
    g, cams = util.small_scene("frustum", P=3000, seed=12, W=W, H=H)
    t = _leafs(g)
    cam_t = util.cams_tensor(cams)
    bg = torch.zeros(3, device="cuda")
    R.check_overflow()
    out = _render(t, cam_t, H, W, bg, pair_capacity=500)
    assert float(out[0].abs().max()) == 0.0
    torch.cuda.synchronize()  # (the mirror copy has landed; in a training loop the next step comes later anyway)
    with pytest.raises(_lib.OcrfError, match="overflow"):
        _render(t, cam_t, H, W, bg, pair_capacity=500)
    ok = _render(t, cam_t, H, W, bg)  # flags were cleared by the raise
    assert float(ok[0].abs().max()) > 0.0
    R.check_overflow()


def test_debug_mode_checks_every_stage_and_snapshots_on_failure(tmp_path, monkeypatch):
    """GaussianRasterizationSettings.debug (PKG:83-90,132-139): per-stage synchronise + check; on an exception the CPU
    copy of the arguments is saved as snapshot_fw.dump and the exception re-raised."""
    from ocrfdet_b200 import _lib
    monkeypatch.chdir(tmp_path)
    W, H = 160, 96
    g, cams = util.small_scene("frustum", P=2000, seed=13, W=W, H=H)
    cam = cams[0]
    gc = util.to_cuda(g)
    st = util.settings_for(cam, [0.1, 0.2, 0.3])._replace(debug=True)
    plain = util.settings_for(cam, [0.1, 0.2, 0.3])
    args = dict(means3D=gc["means3D"].requires_grad_(True), means2D=torch.zeros_like(gc["means3D"]),
                opacities=gc["opacities"], colors_precomp=gc["colors"], scales=gc["scales"], rotations=gc["rotations"])
    c_dbg, r_dbg, d_dbg = R.GaussianRasterizer(st)(**args)
    c_ref, r_ref, d_ref = R.GaussianRasterizer(plain)(**args)
    assert torch.equal(c_dbg, c_ref) and torch.equal(r_dbg, r_ref) and torch.equal(d_dbg, d_ref)
    c_dbg.sum().backward()
    assert not (tmp_path / "snapshot_fw.dump").exists() and not (tmp_path / "snapshot_bw.dump").exists()
    # a failing forward in debug mode: capacity overflow is detected at the binning stage's check
    u = lambda x: x.unsqueeze(0)  # noqa: E731
    with pytest.raises(_lib.OcrfError, match="overflow"):
        R.render_batch(u(gc["means3D"]), u(gc["opacities"]), util.cams_tensor(cams), H, W, torch.zeros(3, device="cuda"),
                       colors_precomp=u(gc["colors"]), scales=u(gc["scales"]), rotations=u(gc["rotations"]),
                       pair_capacity=100, debug=True)
    dump = torch.load(tmp_path / "snapshot_fw.dump")
    assert torch.equal(dump[0][0], gc["means3D"].detach().cpu())
    R.check_overflow()
