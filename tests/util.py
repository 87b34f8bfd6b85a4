"""Shared helpers of the test-suite: scene construction, oracle / product / reference drivers."""
import numpy as np
import torch

from ocrfdet_b200 import rasterizer as R
from ocrfdet_b200.scenes import frustum_scene, ring_scene
from oracle import oracle


def small_scene(kind="frustum", P=2000, seed=0, W=160, H=96, channels=3, n_views=1):
    if kind == "frustum":
        g, cams = frustum_scene(P=P, seed=seed, width=W, height=H, channels=channels)
    else:
        g, cams = ring_scene(P=P, seed=seed, width=W, height=H, channels=channels, n_views=n_views)
    return g, cams


def oracle_forward(g, cam, W, H, bg, **kw):
    return oracle.rasterize(g["means3D"], g["opacities"], g["colors"], cam["viewmatrix"], cam["projmatrix"], W, H,
                            cam["tanfovx"], cam["tanfovy"], np.asarray(bg, np.float32), scales=g.get("scales"),
                            rots=g.get("rotations"), campos=cam["campos"], **kw)


def oracle_backward(g, cam, W, H, bg, out, state, dL_dcolor, dL_dopacity=None, **kw):
    return oracle.rasterize_backward(state, g["means3D"], cam["viewmatrix"], cam["projmatrix"], W, H, cam["tanfovx"],
                                     cam["tanfovy"], np.asarray(bg, np.float32), out, dL_dcolor, dL_dopacity,
                                     scales=g.get("scales"), rots=g.get("rotations"), campos=cam["campos"], **kw)


def to_cuda(g):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in g.items()}


def settings_for(cam, bg, sh_degree=0, scale_modifier=1.0):
    dev = "cuda"
    return R.GaussianRasterizationSettings(
        image_height=cam["height"], image_width=cam["width"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
        bg=torch.tensor(bg, dtype=torch.float32, device=dev), scale_modifier=scale_modifier,
        viewmatrix=torch.from_numpy(cam["viewmatrix"]).to(dev), projmatrix=torch.from_numpy(cam["projmatrix"]).to(dev),
        sh_degree=sh_degree, campos=torch.from_numpy(cam["campos"]).to(dev), prefiltered=False)


def cams_tensor(cams):
    return R.pack_camera_dicts(cams, "cuda")


def assert_grad_elementwise(a, b, what, rtol=1e-4, atol_of_max=1e-6):
    """Per element |a - b| <= rtol |b| + atol_of_max max|b| (the max-norm `rel_err` alone lets a small gradient be
    10 % off unnoticed)."""
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    scale = np.abs(b).max() + 1e-30
    bad = np.abs(a - b) > rtol * np.abs(b) + atol_of_max * scale
    assert not bad.any(), "%s: %d of %d elements beyond %.0e|b| + %.0e max|b|" % (what, int(bad.sum()), bad.size, rtol,
                                                                                 atol_of_max)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def assert_image_close(got, want, ambiguous, tol=1e-5, what="image"):
    """|got - want| <= tol * (1 + |want|) on every pixel whose blend decisions were not borderline;
    borderline pixels (a handful) may differ by one blended Gaussian (<= ~1/255 of the value range)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    amb = np.asarray(ambiguous).astype(bool)
    err = np.abs(got - want) / (1.0 + np.abs(want))
    clean = err[..., ~amb] if err.ndim == 3 else err[~amb]
    assert clean.size == 0 or clean.max() <= tol, "%s: max err %.3g on unambiguous pixels" % (what, clean.max())
    # a flipped blend decision changes a pixel by at most alpha * T * |c| of ONE Gaussian at a threshold: cap it
    assert err.size == 0 or err.max() <= 5e-3, "%s: a borderline pixel is off by %.3g > 5e-3" % (what, err.max())
    assert amb.mean() < 0.01, "%s: too many borderline pixels (%.2f%%)" % (what, 100 * amb.mean())


def cov3d_numpy(scales, rots):
    """[P,6] upper triangle of Sigma = A^T diag(s^2) A for quaternion (r,x,y,z) (forward.cu:118-152)."""
    r, x, y, z = (rots[:, i].astype(np.float64) for i in range(4))
    A = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y),
                  2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x),
                  2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    S = np.einsum("pki,pk,pkj->pij", A, scales.astype(np.float64) ** 2, A)
    return np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1).astype(np.float32)


# ---- the reference's own caller of the plugin (GR = mmdet3d/models/necks/MVSGaussian/lib/gaussian_renderer/__init__.py) ----
GR_PATH = "/root/reference/mmdet3d/models/necks/MVSGaussian/lib/gaussian_renderer/__init__.py"


def load_reference_render(torch_module=None):
    """`render` of the UNMODIFIED reference file, imported where it lies against this repository's
    `diff_gaussian_rasterization` (None when the reference tree is not mounted, e.g. on the GPU box).
    `torch_module`: stand-in for the file's global `torch` (the CPU suite maps device="cuda" to the CPU)."""
    import importlib.util
    import os
    if not os.path.exists(GR_PATH):
        return None
    spec = importlib.util.spec_from_file_location("_reference_gaussian_renderer", GR_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # executes `from diff_gaussian_rasterization import ...`: our import-name shim
    if torch_module is not None:
        mod.torch = torch_module
    return mod.render


def replay_render(data, idx, pts_xyz, pts_rgb, rotations, scales, opacity, bg_color, torch=torch):
    """The call sequence of GR:17-75 restated for boxes without the reference tree (tests/test_abi_and_host.py proves,
    where the tree is mounted, that it reaches the plugin boundary with exactly the arguments of the unmodified file)."""
    import math
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    bg_color = torch.tensor(bg_color, dtype=torch.float32, device="cuda")
    screenspace_points = torch.zeros_like(pts_xyz, dtype=torch.float32, requires_grad=True, device="cuda") + 0
    try:
        screenspace_points.retain_grad()
    except Exception:  # noqa: BLE001
        pass
    raster_settings = GaussianRasterizationSettings(
        image_height=int(data["height"]), image_width=int(data["width"]), tanfovx=math.tan(data["FovX"] * 0.5),
        tanfovy=math.tan(data["FovY"] * 0.5), bg=bg_color, scale_modifier=1.0, viewmatrix=data["world_view_transform"],
        projmatrix=data["full_proj_transform"], sh_degree=3, campos=data["camera_center"], prefiltered=False)
    rendered_image, _, rendered_depth = GaussianRasterizer(raster_settings=raster_settings)(
        means3D=pts_xyz, means2D=screenspace_points, shs=None, colors_precomp=pts_rgb, opacities=opacity, scales=scales,
        rotations=rotations, cov3D_precomp=None)
    return rendered_image, rendered_depth
