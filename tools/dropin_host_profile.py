"""Host-side profile of the unmodified-caller pattern: one GaussianRasterizer call + backward per view (cProfile)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from ocrfdet_b200.scenes import ring_scene  # noqa: E402

W, H, P, V = 704, 256, 100000, 6
g, cams = ring_scene(P=P, seed=1234, width=W, height=H, n_views=V)
names = ("means3D", "scales", "rotations", "opacities", "colors")
flat = {k: torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in names}
bg = torch.zeros(3, device="cuda")
gcol = torch.randn(V, 3, H, W, device="cuda")
settings = [GaussianRasterizationSettings(
    image_height=H, image_width=W, tanfovx=c["tanfovx"], tanfovy=c["tanfovy"], bg=bg, scale_modifier=1.0,
    viewmatrix=torch.from_numpy(c["viewmatrix"]).cuda(), projmatrix=torch.from_numpy(c["projmatrix"]).cuda(),
    sh_degree=3, campos=torch.from_numpy(c["campos"]).cuda(), prefiltered=False) for c in cams]


def step():
    for k in names:
        flat[k].grad = None
    for v in range(V):
        means2D = torch.zeros_like(flat["means3D"], requires_grad=True)
        image, _radii, _depth = GaussianRasterizer(raster_settings=settings[v])(
            means3D=flat["means3D"], means2D=means2D, shs=None, colors_precomp=flat["colors"],
            opacities=flat["opacities"], scales=flat["scales"], rotations=flat["rotations"], cov3D_precomp=None)
        image.backward(gcol[v])


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    step()
torch.cuda.synchronize()
print("ms per 6-view step: %.3f" % ((time.perf_counter() - t0) / 30 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
