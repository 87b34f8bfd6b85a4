"""One forward + backward of the three upstream ops at their full shapes, for ncu (profiles/r1_ops_ncu_summary.csv)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrfdet_b200.bev_pool import bev_pool_v2  # noqa: E402
from ocrfdet_b200.gaussian_heads import GaussianHeads  # noqa: E402
from ocrfdet_b200.scenes import bev_pool_case  # noqa: E402
from ocrfdet_b200.voxel_color import color_voxels_from_images, retain_valid_pixels  # noqa: E402
from tests.golden.make_golden_voxel_color import voxel_color_case  # noqa: E402

torch.manual_seed(0)
n, F = 212992, 80
m = GaussianHeads(F).cuda()
feat = torch.randn(n, F, device="cuda", requires_grad=True)
rgb = torch.rand(n, 3, device="cuda")
for _ in range(2):
    outs = m(feat, rgb)
    sum(o.sum() for o in outs).backward()
c = bev_pool_case(B=8, N=6, D=88, H=16, W=44, C=80, bev=128, seed=3)
t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items() if isinstance(v, np.ndarray)}
d, f = t["depth"].requires_grad_(True), t["feat"].requires_grad_(True)
out = bev_pool_v2(d, f, t["ranks_depth"], t["ranks_feat"], t["ranks_bev"], c["bev_feat_shape"], t["interval_starts"],
                  t["interval_lengths"])
out.sum().backward()
pillars, imgs, mask = voxel_color_case(8, 6, 13, 16384, 256, 704, seed=1)
tp, ti, tm = (torch.from_numpy(a).cuda() for a in (pillars, imgs, mask))
color_voxels_from_images(tp, ti, tm, divisor=255.0)
retain_valid_pixels(ti, tp, tm)
torch.cuda.synchronize()
