"""Golden vectors for voxel colouring / retain_valid_pixels, produced by the REFERENCE's own methods.

Run in the build container (needs /root/reference; CPU only):   python tests/golden/make_golden_voxel_color.py
`lidar_points_to_image_values`, `color_voxels` and `retain_valid_pixels` are lifted out of the view-transformer class
in view_transformer_ocrf.py with `ast` (the module needs mmcv) and executed unmodified under torch with a dummy `self`.
Inputs are regenerated from the seed stored in each file (voxel_color_case below); outputs go to
tests/golden/voxelcolor_*.npz.
"""
import ast
import os
import sys

import numpy as np
import torch

REF = "/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py"
WANTED = ("lidar_points_to_image_values", "color_voxels", "retain_valid_pixels")
CASES = [dict(name="voxelcolor_b1_n6", B=1, N=6, P=5, Q=96, H=24, W=40, seed=21),
         dict(name="voxelcolor_b2_n3", B=2, N=3, P=3, Q=50, H=17, W=31, seed=22)]


def voxel_color_case(B, N, P, Q, H, W, seed, name=None):
    """Projected voxel centres in pixel units, partly outside the image, a visibility mask, 0..255 images."""
    rng = np.random.default_rng(seed)
    pillars = np.stack([rng.uniform(-0.15 * W, 1.15 * W, size=(B, N, P, Q)),
                        rng.uniform(-0.15 * H, 1.15 * H, size=(B, N, P, Q))], -1).astype(np.float32)
    inside = (pillars[..., 0] >= 0) & (pillars[..., 0] <= W - 1) & (pillars[..., 1] >= 0) & (pillars[..., 1] <= H - 1)
    mask = (inside & (rng.random((B, N, P, Q)) < 0.7))[..., None]
    imgs = rng.integers(0, 256, size=(B, N, 3, H, W)).astype(np.float32)
    return pillars, imgs, mask


def reference_methods():
    tree = ast.parse(open(REF).read())
    fns = [n for cls in tree.body if isinstance(cls, ast.ClassDef) for n in cls.body
           if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert sorted(f.name for f in fns) == sorted(WANTED), [f.name for f in fns]
    ns = {"torch": torch, "F": torch.nn.functional}
    exec(compile(ast.Module(body=fns, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    ns = reference_methods()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for case in CASES:
        pillars, imgs, mask = voxel_color_case(**case)
        tp, ti, tm = torch.from_numpy(pillars), torch.from_numpy(imgs), torch.from_numpy(mask)
        B, N, P, Q, _ = pillars.shape
        img_values = ns["lidar_points_to_image_values"](None, tp, ti, tm)                      # :1068
        voxels = torch.zeros(B, P, Q, 3)
        colored, avg, valid = ns["color_voxels"](None, voxels, img_values, tm)                  # :1070
        sparse = ns["retain_valid_pixels"](None, ti, tp.view(B, N, P, Q, 1, 2), tm.view(B, N, P, Q, 1, 1))  # :1077
        np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"), case=str(case), avg=avg.numpy(),
                            valid=valid.numpy(), colored=colored.numpy(), sparse=sparse.numpy())
        print("wrote", case["name"], float(avg.abs().max()), int(valid.sum()), int((sparse != 255).sum()))


if __name__ == "__main__":
    sys.exit(main())
