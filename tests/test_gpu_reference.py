"""The CUDA path against the reference's OWN CUDA rasterizer (oracle/_ref/libinria_ref.so, built
unmodified from the vendored Inria sources).  Keys, sort order and tile ranges must be bit-exact;
colour within 1e-5; gradients within 1e-4 (atomic order differs)."""
import numpy as np
import pytest
import torch

from ocrfdet_b200 import rasterizer as R
from oracle import ref
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libinria_ref.so not built")]


@pytest.mark.parametrize("kind,kw", [("frustum", dict(P=10000, seed=21, W=704, H=256)),
                                     ("ring", dict(P=30000, seed=22, W=704, H=256)),
                                     ("frustum", dict(P=800, seed=23, W=90, H=70))])
@pytest.mark.parametrize("binning", ["split", "depthfirst", "pairsort"])
def test_against_reference_cuda(kind, kw, binning, monkeypatch):
    monkeypatch.setenv("OCRF_BINNING", binning)
    g, cams = util.small_scene(kind, **kw)
    cam, W, H = cams[0], kw["W"], kw["H"]
    bg = [0.4, 0.3, 0.2]
    gc = util.to_cuda(g)
    st = util.settings_for(cam, bg)
    rr = ref.RefRasterizer()
    rcol, rradii, rN = rr.forward(gc["means3D"], gc["opacities"], gc["colors"], st.viewmatrix, st.projmatrix, st.campos,
                                  W, H, st.tanfovx, st.tanfovy, st.bg, scales=gc["scales"], rotations=gc["rotations"])
    rs = rr.state()

    R.KEEP_STATE = True
    for k in ("means3D", "scales", "rotations", "opacities", "colors"):
        gc[k].requires_grad_(True)
    means2D = torch.zeros_like(gc["means3D"], requires_grad=True)
    color, radii, _depth, opac = R.GaussianRasterizer(st, return_opacity=True)(
        means3D=gc["means3D"], means2D=means2D, opacities=gc["opacities"], colors_precomp=gc["colors"],
        scales=gc["scales"], rotations=gc["rotations"])
    ms = R.last_state(reference_lists=True)
    # ---- integer / index state: bit-exact ----
    assert ms["num_pairs"] == rN
    assert torch.equal(radii, rradii)
    assert torch.equal(ms["tiles_touched"][0], rs["tiles_touched"])
    assert torch.equal(ms["offsets"], rs["offsets"])
    assert torch.equal(ms["keys"] if "keys" in ms else ms["keys_ref"], rs["keys"]), "sorted (tile|depth) keys"
    assert torch.equal(ms["point_list"] if "point_list" in ms else ms["point_list_ref"], rs["point_list"]), "sort order"
    assert torch.equal(ms["ranges"][0], rs["ranges"]), "tile ranges"
    vis = radii > 0
    for name in ("depths", "xy", "conic_opacity"):
        a, b = ms[name][0][vis].contiguous().view(torch.int32), rs[name][vis].contiguous().view(torch.int32)
        assert torch.equal(a, b), "%s bit pattern" % name
    # ---- forward colour ----
    err = ((color - rcol).abs() / (1 + rcol.abs()))
    frac_bad = float((err > 1e-5).float().mean())
    assert frac_bad < 1e-3, "colour: %.4f%% of values off by more than 1e-5 (max %.3g)" % (100 * frac_bad, float(err.max()))
    assert float(err.max()) <= 5e-3, "colour: a borderline pixel is off by %.3g (one flipped blend is <= alpha*T*|c|)" % float(err.max())
    # ---- the reference's image state: accumulated opacity = 1 - final_T, and n_contrib (pinned; the median DEPTH
    #      output has no reference source in the tree -- the w-depth fork is absent -- and stays unpinned) ----
    eo = (opac[0] - (1.0 - rs["final_T"])).abs()
    assert float((eo > 1e-5).float().mean()) < 1e-3 and float(eo.max()) <= 5e-3, "opacity vs 1 - final_T: max %.3g" % float(eo.max())
    same_nc = ms["n_contrib"][0] == rs["n_contrib"]
    assert float((~same_nc).float().mean()) < 1e-3, "n_contrib differs on %.4f%% of the pixels" % (100 * float((~same_nc).float().mean()))
    # ---- gradients ----
    rng = np.random.default_rng(9)
    gcol = torch.from_numpy(rng.normal(size=(3, H, W)).astype(np.float32)).cuda()
    (color * gcol).sum().backward()
    gr = rr.backward(gc["means3D"].detach(), gc["colors"].detach(), st.viewmatrix, st.projmatrix, st.campos, st.tanfovx,
                     st.tanfovy, st.bg, rradii, gcol, scales=gc["scales"].detach(), rotations=gc["rotations"].detach())
    torch.cuda.synchronize()
    # Screen-space gradients (what the blend kernel produces) are well conditioned: 1e-4 directly.
    # The per-Gaussian chain behind them (conic -> covariance -> scale/quaternion) cancels heavily in
    # float32 -- the reference build itself is ~1e-4 away from the float64 value of its own formulas --
    # so for those outputs the bar is: we are within 1e-4 of the float64 value, and our distance to
    # the reference is explained by the reference's own distance to it.
    want, wst = util.oracle_forward(g, cam, W, H, bg)
    truth = util.oracle_backward(g, cam, W, H, bg, want, wst, gcol.cpu().numpy())
    for name, got, refv in (("opacities", gc["opacities"].grad.reshape(-1), gr["opacities"].reshape(-1)),
                            ("colors", gc["colors"].grad, gr["colors"]), ("means2D", means2D.grad, gr["means2D"])):
        e = util.rel_err(got.cpu().numpy(), refv.cpu().numpy())
        assert e <= 1e-4, "%s gradient vs reference CUDA: rel err %.3g" % (name, e)
    for name, got, refv in (("means3D", gc["means3D"].grad, gr["means3D"]), ("scales", gc["scales"].grad, gr["scales"]),
                            ("rotations", gc["rotations"].grad, gr["rotations"])):
        t = np.asarray(truth[name], np.float64)
        e_mine, e_ref = util.rel_err(got.cpu().numpy(), t), util.rel_err(refv.cpu().numpy(), t)
        e_pair = util.rel_err(got.cpu().numpy(), refv.cpu().numpy())
        assert e_mine <= 1e-4, "%s gradient vs float64 oracle: rel err %.3g (reference: %.3g)" % (name, e_mine, e_ref)
        assert e_pair <= 1e-4 + e_ref, "%s gradient vs reference CUDA: %.3g > 1e-4 + %.3g" % (name, e_pair, e_ref)
    rr.close()


def test_mark_visible_matches_reference():
    g, cams = util.small_scene("ring", P=5000, seed=24, W=352, H=128)
    cam = cams[0]
    st = util.settings_for(cam, [0, 0, 0])
    m = torch.from_numpy(g["means3D"]).cuda()
    assert torch.equal(R.GaussianRasterizer(st).markVisible(m), ref.mark_visible(m, st.viewmatrix, st.projmatrix))
