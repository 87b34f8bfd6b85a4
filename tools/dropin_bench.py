"""The pure drop-in number: OcRFDet's own call pattern -- one GaussianRasterizer call per camera view through the
`diff_gaussian_rasterization` import name, forward + backward, exact sizing (one 8-byte read-back per call) -- on the
bench workload (6 views 256x704, 100k Gaussians).  No batching, no capacity mode: what a user gets by swapping the
package and changing nothing else.  python tools/dropin_bench.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from ocrfdet_b200.scenes import ring_scene  # noqa: E402

W, H, P, V = 704, 256, 100000, 6
g, cams = ring_scene(P=P, seed=1234, width=W, height=H, n_views=V)
names = ("means3D", "scales", "rotations", "opacities", "colors")
dev = {k: torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in names}
bg = torch.zeros(3, device="cuda")
gcol = torch.randn(V, 3, H, W, device="cuda")
settings = [GaussianRasterizationSettings(
    image_height=H, image_width=W, tanfovx=c["tanfovx"], tanfovy=c["tanfovy"], bg=bg, scale_modifier=1.0,
    viewmatrix=torch.from_numpy(c["viewmatrix"]).cuda(), projmatrix=torch.from_numpy(c["projmatrix"]).cuda(), sh_degree=3,
    campos=torch.from_numpy(c["campos"]).cuda(), prefiltered=False) for c in cams]


def step():
    for k in names:
        dev[k].grad = None
    for v in range(V):
        means2D = torch.zeros_like(dev["means3D"], requires_grad=True)
        image, radii, depth = GaussianRasterizer(raster_settings=settings[v])(
            means3D=dev["means3D"], means2D=means2D, shs=None, colors_precomp=dev["colors"], opacities=dev["opacities"],
            scales=dev["scales"], rotations=dev["rotations"], cov3D_precomp=None)
        image.backward(gcol[v])


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(30):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.mean(ts))
print(json.dumps({"dropin_per_view_calls_ms_per_step": ms, "views_per_s": V / ms * 1e3, "views": V}))
