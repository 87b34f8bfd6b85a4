"""One forward + backward of BASELINE config 4 (2 frames x 6 views, 80 feature channels), for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrfdet_b200 import rasterizer as R  # noqa: E402
from ocrfdet_b200.scenes import ring_scene  # noqa: E402

S, P, V, W, H, C = 2, 100000, 6, 704, 256, 80
gs = [ring_scene(P=P, seed=1234 + s, width=W, height=H, channels=C, n_views=V) for s in range(S)]
cams = sum([g[1] for g in gs], [])
names = ("means3D", "scales", "rotations", "opacities", "colors")
dev = {k: torch.from_numpy(np.stack([g[0][k] for g in gs])).cuda().requires_grad_(True) for k in names}
cam_t = R.pack_camera_dicts(cams, "cuda")
bg = torch.zeros(C, device="cuda")
gcol, gop = torch.randn(S * V, C, H, W, device="cuda"), torch.randn(S * V, 1, H, W, device="cuda")
for _ in range(3):
    for k in names:
        dev[k].grad = None
    c, r, d, o = R.render_batch(dev["means3D"], dev["opacities"], cam_t, H, W, bg, colors_precomp=dev["colors"],
                                scales=dev["scales"], rotations=dev["rotations"])
    torch.autograd.backward([c, o], [gcol, gop])
torch.cuda.synchronize()
