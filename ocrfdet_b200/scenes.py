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


def bev_pool_case(B=1, N=6, D=88, H=16, W=44, C=80, bev=128, seed=0, pc=51.2, depth_range=(1.0, 45.0)):
    """Rank / interval tensors of an LSS lift as `voxel_pooling_prepare_v2` builds them
    (view_transformer_ocrf.py:697-748): every (camera, depth bin, pixel) frustum point is dropped into a
    bev x bev grid over +-pc metres (Z collapsed), points outside are filtered, the rest are sorted by BEV cell.
    Six ring cameras with a 70 degree horizontal field of view.  Returns numpy arrays:
    depth [B,N,D,H,W], feat [B,N,H,W,C], ranks_depth, ranks_feat, ranks_bev (int32, sorted by ranks_bev),
    interval_starts, interval_lengths, n_bev = B*bev*bev."""
    rng = np.random.default_rng(seed)
    d = depth_range[0] + (depth_range[1] - depth_range[0]) * (np.arange(D) + 0.5) / D
    half = np.tan(np.deg2rad(35.0))
    u = ((np.arange(W) + 0.5) / W * 2 - 1) * half
    v = ((np.arange(H) + 0.5) / H * 2 - 1) * half * H / W
    yaw = np.deg2rad(np.array([0, 55, 110, 180, -110, -55])[np.arange(N) % 6] + 360.0 * (np.arange(N) // 6) / max(N, 1))
    dd, vv, uu = np.meshgrid(d, v, u, indexing="ij")          # [D,H,W]
    xc, zc = uu * dd, dd                                          # camera frame: x right, z forward
    ranks_depth, ranks_feat, ranks_bev = [], [], []
    for b in range(B):
        shift = rng.uniform(-2, 2, size=2)                        # per-sample ego jitter (augmentation)
        for n in range(N):
            ex = np.cos(yaw[n]) * zc + np.sin(yaw[n]) * xc + shift[0]
            ey = np.sin(yaw[n]) * zc - np.cos(yaw[n]) * xc + shift[1]
            ix = np.floor((ex + pc) / (2 * pc) * bev).astype(np.int64)
            iy = np.floor((ey + pc) / (2 * pc) * bev).astype(np.int64)
            kept = (ix >= 0) & (ix < bev) & (iy >= 0) & (iy < bev)
            di, hi, wi = np.nonzero(kept)
            bn = b * N + n
            ranks_depth.append(((bn * D + di) * H + hi) * W + wi)
            ranks_feat.append((bn * H + hi) * W + wi)
            ranks_bev.append((b * bev + iy[kept]) * bev + ix[kept])
    rd, rf, rb = (np.concatenate(a) for a in (ranks_depth, ranks_feat, ranks_bev))
    order = np.argsort(rb, kind="stable")
    rd, rf, rb = rd[order].astype(np.int32), rf[order].astype(np.int32), rb[order].astype(np.int32)
    kept = np.ones(len(rb), bool)
    kept[1:] = rb[1:] != rb[:-1]
    starts = np.nonzero(kept)[0].astype(np.int32)
    lengths = np.diff(np.append(starts, len(rb))).astype(np.int32)
    depth = rng.random((B, N, D, H, W), dtype=np.float32)
    depth /= depth.sum(2, keepdims=True)                          # a softmax-like depth distribution
    feat = rng.normal(size=(B, N, H, W, C)).astype(np.float32)
    return dict(depth=depth, feat=feat, ranks_depth=rd, ranks_feat=rf, ranks_bev=rb, interval_starts=starts,
                interval_lengths=lengths, n_bev=B * bev * bev, bev_feat_shape=(B, 1, bev, bev, C))
