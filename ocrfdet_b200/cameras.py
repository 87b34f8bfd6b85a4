"""Camera-matrix conventions of the render path's caller (the input contract).

Restates, for tests / benches / callers that batch views, what OcRFDet does per sample on the host:
  getWorld2View2       /root/reference/mmdet3d/models/necks/MVSGaussian/lib/utils/data_utils.py:703-714
  getProjectionMatrix  /root/reference/mmdet3d/models/necks/MVSGaussian/lib/utils/data_utils.py:716-734
  assembly + transposes /root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:1143-1149
All matrices handed to the rasterizer are TRANSPOSED (row-vector convention), i.e. the kernels read
element (r, c) at m[4c + r].
"""
import math

import numpy as np

ZNEAR, ZFAR = 0.01, 999.9  # view_transformer_ocrf.py:629-630


def world_to_view(R, t, translate=(0.0, 0.0, 0.0), scale=1.0):
    """R: camera-to-world rotation [3,3]; t: world-to-camera translation [3] (3DGS convention)."""
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = np.asarray(R, dtype=np.float64).T
    Rt[:3, 3] = np.asarray(t, dtype=np.float64)
    Rt[3, 3] = 1.0
    c2w = np.linalg.inv(Rt)
    c2w[:3, 3] = (c2w[:3, 3] + np.asarray(translate, dtype=np.float64)) * scale
    return np.linalg.inv(c2w).astype(np.float32)


def projection_matrix(znear, zfar, K, h, w):
    K = np.asarray(K, dtype=np.float32)
    near_fx, near_fy = np.float32(znear) / K[0, 0], np.float32(znear) / K[1, 1]
    left, right = -(w - K[0, 2]) * near_fx, K[0, 2] * near_fx
    bottom, top = (K[1, 2] - h) * near_fy, K[1, 2] * near_fy
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(K, R_c2w, t_w2c, width, height, znear=ZNEAR, zfar=ZFAR):
    """Everything `GaussianRasterizationSettings` needs for one view, as float32 numpy arrays."""
    K = np.asarray(K, dtype=np.float32)
    view_t = world_to_view(R_c2w, t_w2c).T.copy()                      # world_view_transform (transposed)
    proj_t = projection_matrix(znear, zfar, K, height, width).T.copy()  # projection_matrix (transposed)
    full_t = (view_t @ proj_t).astype(np.float32)                       # full_proj_transform
    campos = np.linalg.inv(view_t)[3, :3].astype(np.float32)            # camera_center
    fovx = 2.0 * math.atan(width / (2.0 * float(K[0, 0])))
    fovy = 2.0 * math.atan(height / (2.0 * float(K[1, 1])))
    return dict(viewmatrix=view_t, projmatrix=full_t, campos=campos, tanfovx=math.tan(fovx * 0.5),
                tanfovy=math.tan(fovy * 0.5), width=int(width), height=int(height))


def ego_ring_cameras(width=704, height=256, yaws_deg=(0.0, 55.0, -55.0, 110.0, -110.0, 180.0), cam_height=1.5,
                     fx=557.2):
    """Six nuScenes-like cameras around the ego origin: pinhole fx = fy = 557.2 at
    704x256 (scaled with the width), principal point at the image centre, yaw about ego +z."""
    f = fx * width / 704.0
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1]], dtype=np.float32)
    base = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], dtype=np.float64)  # camera axes (x right, y down, z fwd) in ego
    cams = []
    for yaw in yaws_deg:
        a = math.radians(yaw)
        Rz = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
        R_c2w = Rz @ base
        centre = np.array([0.0, 0.0, cam_height])
        t_w2c = -R_c2w.T @ centre
        cams.append(make_camera(K, R_c2w, t_w2c, width, height))
    return cams
