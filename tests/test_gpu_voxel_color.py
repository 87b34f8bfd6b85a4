"""Scope row f-4: fused voxel colouring and retain_valid_pixels against the C oracle, the golden vectors recorded from
the reference's own methods, and torch's grid_sample at the full OcRF shape."""
import ast
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle
from tests.golden.make_golden_voxel_color import voxel_color_case
from tests.test_oracle_golden import GOLDEN_VC

pytestmark = pytest.mark.gpu


def cuda(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


@pytest.mark.parametrize("path", GOLDEN_VC, ids=[os.path.basename(p) for p in GOLDEN_VC])
def test_matches_reference_golden(path):
    from ocrfdet_b200.voxel_color import color_voxels_from_images, retain_valid_pixels
    z = np.load(path)
    pillars, imgs, mask = voxel_color_case(**ast.literal_eval(str(z["case"])))
    tp, ti, tm = cuda(pillars, imgs, mask)
    avg, valid = color_voxels_from_images(tp, ti, tm)
    assert np.array_equal(valid.cpu().numpy(), z["valid"])
    assert np.abs(avg.cpu().numpy() - z["avg"]).max() <= 1e-5 * 255
    B, N, P, Q, _ = pillars.shape
    sparse = retain_valid_pixels(ti, tp.view(B, N, P, Q, 1, 2), tm.view(B, N, P, Q, 1, 1))
    assert np.array_equal(sparse.cpu().numpy(), z["sparse"])


@pytest.mark.parametrize("case", [dict(B=1, N=6, P=13, Q=1024, H=64, W=176, seed=3), dict(B=2, N=1, P=1, Q=7, H=2, W=2, seed=4),
                                  dict(B=1, N=2, P=2, Q=300, H=9, W=5, seed=5)])
def test_matches_oracle_bit_for_bit(case):
    """Same float operations in the same order as the oracle: exact equality, including points outside the image."""
    from ocrfdet_b200.voxel_color import color_voxels_from_images, retain_valid_pixels
    pillars, imgs, mask = voxel_color_case(**case)
    pillars[0, 0, 0, :3] = [[-1.0, 3.0], [0.0, 0.0], [case["W"] - 1, case["H"] - 1]]   # the -1 marker and both corners
    mask[0, 0, 0, :3] = True
    tp, ti, tm = cuda(pillars, imgs, mask)
    avg, valid = color_voxels_from_images(tp, ti, tm, divisor=255.0)
    wavg, wvalid = oracle.color_voxels(pillars, imgs, mask, divisor=255.0)
    assert np.array_equal(valid.cpu().numpy(), wvalid)
    assert np.array_equal(avg.cpu().numpy(), wavg)
    sparse = retain_valid_pixels(ti, tp, tm)
    assert np.array_equal(sparse.cpu().numpy(), oracle.retain_valid_pixels(imgs, pillars, mask))


def test_full_shape_against_grid_sample():
    """One OcRF sample: 6 cameras, 13 x 16384 voxels, 256x704 images; torch's own grid_sample as the comparator."""
    from ocrfdet_b200.voxel_color import color_voxels_from_images, retain_valid_pixels
    B, N, P, Q, H, W = 1, 6, 13, 16384, 256, 704
    pillars, imgs, mask = voxel_color_case(B, N, P, Q, H, W, seed=8)
    tp, ti, tm = cuda(pillars, imgs, mask)
    avg, valid = color_voxels_from_images(tp, ti, tm)
    grid = tp / torch.tensor([W - 1, H - 1], device="cuda") * 2 - 1
    vals = F.grid_sample(ti.view(B * N, 3, H, W), grid.view(B * N, 1, P * Q, 2), align_corners=True)
    vals = vals.view(B, N, 3, P, Q).permute(0, 1, 3, 4, 2) * tm.float()
    want = vals.sum(1) / tm.float().sum(1).clamp(min=1)
    assert float((avg - want).abs().max()) <= 1e-4 * 255   # torch's kernel contracts to fma; ours follows the CPU order
    assert torch.equal(valid, tm.squeeze(-1).any(1))
    sparse = retain_valid_pixels(ti, tp, tm)
    hit = torch.zeros(B * N, H * W, dtype=torch.bool, device="cuda")
    m = tm.view(B * N, -1)
    xy = tp.view(B * N, -1, 2).long().clamp(0, max(W, H) - 1)
    for v in range(B * N):
        sel = xy[v][m[v]]
        hit[v, sel[:, 1] * W + sel[:, 0]] = True
    want_sparse = torch.where(hit.view(B, N, 1, H, W), ti, torch.full_like(ti, 255.0))
    assert torch.equal(sparse, want_sparse)


def test_edge_cases_and_errors():
    from ocrfdet_b200.voxel_color import color_voxels_from_images, retain_valid_pixels
    imgs = torch.rand(1, 2, 3, 8, 8, device="cuda") * 255
    pillars = torch.zeros(1, 2, 1, 0, 2, device="cuda")
    mask = torch.zeros(1, 2, 1, 0, 1, dtype=torch.bool, device="cuda")
    avg, valid = color_voxels_from_images(pillars, imgs, mask)
    assert tuple(avg.shape) == (1, 1, 0, 3) and tuple(valid.shape) == (1, 1, 0)
    out = retain_valid_pixels(imgs, pillars, mask)
    assert float((out - 255).abs().max()) == 0.0
    with pytest.raises(ValueError):
        color_voxels_from_images(torch.zeros(1, 2, 1, 4, 3, device="cuda"), imgs, torch.zeros(1, 2, 1, 4, 1, device="cuda"))
    with pytest.raises(Exception):
        color_voxels_from_images(pillars.cpu(), imgs.cpu(), mask.cpu())
