"""One sample of BASELINE config 5 (6 views 512x1408, 1 M Gaussians): two forward + backward steps, for the ncu launch list."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrfdet_b200 import rasterizer as R  # noqa: E402
from ocrfdet_b200.scenes import ring_scene  # noqa: E402

S, P, V, W, H, C = 1, 1_000_000, 6, 1408, 512, 3
g, cams = ring_scene(P=P, seed=4321, width=W, height=H, channels=C, n_views=V, bev=277)
names = ("means3D", "scales", "rotations", "opacities", "colors")
dev = {k: torch.from_numpy(g[k][None]).cuda().requires_grad_(True) for k in names}
cam_t = R.pack_camera_dicts(cams, "cuda")
bg = torch.zeros(C, device="cuda")
gcol, gop = torch.randn(V, C, H, W, device="cuda"), torch.randn(V, 1, H, W, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for k in names:
        dev[k].grad = None
    c, r, d, o = R.render_batch(dev["means3D"], dev["opacities"], cam_t, H, W, bg, colors_precomp=dev["colors"],
                                scales=dev["scales"], rotations=dev["rotations"])
    torch.autograd.backward([c, o], [gcol, gop])
torch.cuda.synchronize()
