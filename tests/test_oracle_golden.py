"""CPU suite: the C oracle against the golden vectors recorded from the REFERENCE's own CUDA rasterizer
(tests/golden/*.npz, produced by tests/golden/make_golden.py on a B200).  This is what pins the oracle."""
import ast
import glob
import os

import numpy as np
import pytest

from oracle import oracle
from tests import util
from tests.golden.make_golden import scene

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("heads_", "bevpool_", "voxelcolor_", "hoa_")))
GOLDEN_HEADS = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "heads_*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_cuda_golden(path):
    z = np.load(path)
    case = ast.literal_eval(str(z["case"]))
    g, cam = scene(case)
    W, H, bg = case["W"], case["H"], case["bg"]
    out, st = util.oracle_forward(g, cam, W, H, bg)
    pre, b = st["pre"], st["bin"]
    # ---- integer / index state: bit-exact against the reference CUDA build ----
    assert b["N"] == int(z["num_rendered"])
    assert np.array_equal(pre["radii"], z["radii"])
    assert np.array_equal(pre["tiles_touched"], z["tiles_touched"].view(np.uint32))
    assert np.array_equal(b["offsets"], z["offsets"].view(np.uint32))
    assert np.array_equal(b["keys"], z["keys"].view(np.uint64)), "sorted (tile|depth) keys"
    assert np.array_equal(b["point_list"], z["point_list"].view(np.uint32)), "sort order"
    assert np.array_equal(b["ranges"], z["ranges"].view(np.uint32)), "tile ranges"
    vis = pre["radii"] > 0
    for name, key in (("depths", "depths"), ("xy", "xy"), ("conic_opacity", "conic_opacity")):
        assert np.array_equal(pre[name][vis].view(np.uint32), z[key][vis].view(np.uint32)), name + " bit pattern"
    # ---- forward: 1e-5 off the borderline pixels; blend bookkeeping identical there ----
    amb = out["ambiguous"].astype(bool)
    util.assert_image_close(out["color"], z["color"], amb, 1e-5, "colour vs reference CUDA")
    assert np.array_equal(out["n_contrib"][~amb], z["n_contrib"].view(np.uint32)[~amb])
    assert np.abs(out["final_T"] - z["final_T"])[~amb].max() <= 1e-6
    # ---- backward: the reference's float32 result vs the oracle ----
    gcol = np.random.default_rng(case["seed"] + 7).normal(size=(3, H, W)).astype(np.float32)
    gw = util.oracle_backward(g, cam, W, H, bg, out, st, gcol)
    for name, key in (("means2D", "g_means2D"), ("opacities", "g_opacities"), ("colors", "g_colors")):
        ref = z[key][:, :2] if name == "means2D" else z[key].reshape(np.asarray(gw[name]).shape)
        assert util.rel_err(ref, gw[name]) <= 1e-4, name
    conic = z["g_conic"].reshape(-1, 4)[:, [0, 1, 3]]
    assert util.rel_err(conic, gw["conic"]) <= 1e-4
    # per-Gaussian chain: ill-conditioned in float32 (see tests/test_gpu_reference.py) -> looser bound
    for name, key in (("means3D", "g_means3D"), ("scales", "g_scales"), ("rotations", "g_rotations")):
        assert util.rel_err(z[key], gw[name]) <= 5e-4, name


def heads_params(z):
    return {h: {k: z["%s.%s" % (h, k)] for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")}
            for h in oracle.HEAD_ORDER}


@pytest.mark.parametrize("path", GOLDEN_HEADS, ids=[os.path.basename(p) for p in GOLDEN_HEADS])
def test_heads_oracle_matches_reference_modules(path):
    """tests/golden/heads_*.npz were produced by executing the reference's own four nn.Module classes
    (tests/golden/make_golden_heads.py); float32, so 1e-5 forward / 1e-4 gradients."""
    z = np.load(path)
    params = heads_params(z)
    op, sc, rot, col = oracle.gaussian_heads_forward(z["feat"], z["rgb"], params)
    for got, key in ((op, "opacity"), (sc, "scaling"), (rot, "rotation"), (col, "color")):
        assert np.abs(got - z[key]).max() <= 1e-5 * (1 + np.abs(z[key]).max()), key
    g_feat, grads = oracle.gaussian_heads_backward(z["feat"], z["rgb"], params, z["g_opacity"], z["g_scaling"],
                                                   z["g_rotation"], z["g_color"])
    assert util.rel_err(g_feat, z["g_feat"]) <= 1e-4
    for h in oracle.HEAD_ORDER:
        for k, v in grads[h].items():
            assert util.rel_err(v, z["g.%s.%s" % (h, k)]) <= 1e-4, (h, k)


def test_heads_golden_present():
    assert len(GOLDEN_HEADS) >= 2


GOLDEN_BEV = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bevpool_*.npz")))


@pytest.mark.parametrize("path", GOLDEN_BEV, ids=[os.path.basename(p) for p in GOLDEN_BEV])
def test_bev_pool_oracle_matches_reference_cuda_golden(path):
    """tests/golden/bevpool_*.npz hold the outputs of the reference's own bev_pool_v2 CUDA kernels on a B200
    (tests/golden/make_golden_bev.py).  The oracle keeps their loops and fma contraction: bit-exact."""
    from tests.golden.make_golden_bev import case_inputs
    z = np.load(path)
    c = case_inputs(ast.literal_eval(str(z["case"])))
    out = oracle.bev_pool_forward(c["depth"], c["feat"], c["ranks_depth"], c["ranks_feat"], c["ranks_bev"], c["n_bev"],
                                  c["interval_starts"], c["interval_lengths"])
    assert np.array_equal(out, z["out"])
    dg, fg = oracle.bev_pool_backward(c["out_grad"], c["depth"], c["feat"], c["ranks_depth"], c["ranks_feat"],
                                      c["ranks_bev"])
    assert np.array_equal(fg, z["feat_grad"])
    assert np.array_equal(dg, z["depth_grad"])


def test_bev_pool_golden_present():
    assert len(GOLDEN_BEV) >= 2


GOLDEN_VC = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "voxelcolor_*.npz")))


@pytest.mark.parametrize("path", GOLDEN_VC, ids=[os.path.basename(p) for p in GOLDEN_VC])
def test_voxel_color_oracle_matches_reference_methods(path):
    """tests/golden/voxelcolor_*.npz were produced by executing the reference's own lidar_points_to_image_values,
    color_voxels and retain_valid_pixels (tests/golden/make_golden_voxel_color.py).  Colours live on a 0..255 scale:
    1e-5 relative to that scale; the sparse image and the validity mask are exact."""
    from tests.golden.make_golden_voxel_color import voxel_color_case
    z = np.load(path)
    pillars, imgs, mask = voxel_color_case(**ast.literal_eval(str(z["case"])))
    avg, valid = oracle.color_voxels(pillars, imgs, mask)
    assert np.array_equal(valid, z["valid"])
    assert np.abs(avg - z["avg"]).max() <= 1e-5 * 255
    assert np.array_equal(avg[~valid], np.zeros_like(avg[~valid]))
    sparse = oracle.retain_valid_pixels(imgs, pillars, mask)
    assert np.array_equal(sparse, z["sparse"])


def test_voxel_color_golden_present():
    assert len(GOLDEN_VC) >= 2
