"""Seeded synthetic nuScenes-shaped scenes for tests and benches.

Gaussians sit on OcRFDet's voxel grid (/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:651-673:
128 x 128 pillars over +-51.2 m, 13 heights) with the parameter distributions the OcRF MLP heads
produce at initialisation (VT:1128-1133: sigmoid opacity / colour, softplus scale, unit quaternion).
Generated with numpy on the host so the oracle and the GPU path see bit-identical inputs.
"""
import numpy as np

from .cameras import ego_ring_cameras, make_camera


def voxel_centres(bev=128, heights=13, pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)):
    Z = 8
    zs = np.concatenate([np.linspace(3, Z - 1, 5), np.linspace(0.5, Z - 0.5, heights - 5)]) / Z
    xs = (np.arange(bev) + 0.5) / bev
    ys = (np.arange(bev) + 0.5) / bev
    zz, yy, xx = np.meshgrid(zs, ys, xs, indexing="ij")
    pts = np.stack([xx, yy, zz], -1).reshape(-1, 3)
    lo, hi = np.array(pc_range[:3]), np.array(pc_range[3:])
    return (pts * (hi - lo) + lo).astype(np.float32)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gaussians_on_grid(P, seed, channels=3, bev=128, scale_mean=-1.0):
    rng = np.random.default_rng(seed)
    grid = voxel_centres(bev=bev)
    idx = rng.choice(grid.shape[0], size=P, replace=P > grid.shape[0])
    idx.sort()
    means = grid[idx] + rng.uniform(-0.4, 0.4, size=(P, 3)).astype(np.float32)
    scales = np.log1p(np.exp(rng.normal(scale_mean, 1.0, size=(P, 3)))).astype(np.float32)
    q = rng.normal(size=(P, 4))
    rots = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    opac = _sigmoid(rng.normal(0.0, 1.5, size=(P, 1))).astype(np.float32)
    colors = _sigmoid(rng.normal(0.0, 1.0, size=(P, channels))).astype(np.float32)
    return dict(means3D=means.astype(np.float32), scales=scales, rotations=rots, opacities=opac, colors=colors)


def frustum_scene(P=10000, seed=0, width=704, height=256, channels=3):
    """Config 1: one camera at the origin looking down +z, P Gaussians placed inside its frustum."""
    rng = np.random.default_rng(seed)
    fx = 557.2 * width / 704.0
    K = np.array([[fx, 0, width / 2.0], [0, fx, height / 2.0], [0, 0, 1]], dtype=np.float32)
    cam = make_camera(K, np.eye(3), np.zeros(3), width, height)
    z = rng.uniform(1.0, 60.0, size=P)
    x = z * cam["tanfovx"] * rng.uniform(-1.1, 1.1, size=P)
    y = z * cam["tanfovy"] * rng.uniform(-1.1, 1.1, size=P)
    g = gaussians_on_grid(P, seed + 1, channels=channels)
    g["means3D"] = np.stack([x, y, z], -1).astype(np.float32)
    return g, [cam]


def ring_scene(P=100000, seed=0, width=704, height=256, channels=3, n_views=6, bev=128):
    """Configs 2-5: P voxel-grid Gaussians seen by the six ego-ring cameras."""
    g = gaussians_on_grid(P, seed, channels=channels, bev=bev)
    cams = ego_ring_cameras(width, height)[:n_views]
    return g, cams
