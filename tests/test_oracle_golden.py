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

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


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
