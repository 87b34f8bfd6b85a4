"""Voxel colouring + retain_valid_pixels at the OcRF training shape (B samples x 6 cameras, 13 x 128 x 128 voxels,
256x704 images): the fused ops against a torch formulation structured like the reference's
(view_transformer_ocrf.py:924-971: grid_sample + masked mean over cameras; :1004-1022: a Python loop over
B x 6 x 13 boolean-mask scatters).  python tools/voxel_color_bench.py"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrfdet_b200.voxel_color import color_voxels_from_images, retain_valid_pixels  # noqa: E402
from tests.golden.make_golden_voxel_color import voxel_color_case  # noqa: E402


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def torch_color(pillars, imgs, mask):
    B, N, P, Q, _ = pillars.shape
    _, _, C, H, W = imgs.shape
    grid = pillars.clone()
    grid[..., 0] = pillars[..., 0] / (W - 1) * 2 - 1
    grid[..., 1] = pillars[..., 1] / (H - 1) * 2 - 1
    vals = F.grid_sample(imgs.view(B * N, C, H, W), grid.view(B * N, 1, P * Q, 2), align_corners=True)
    vals = vals.view(B, N, C, P, Q).permute(0, 1, 3, 4, 2) * mask.float()
    m = mask.squeeze(-1)
    masked = torch.where(m[..., None], vals, torch.zeros_like(vals))
    cnt = m.sum(1, keepdim=True).float().unsqueeze(-1)
    cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
    return (masked.sum(1, keepdim=True) / cnt).squeeze(1), m.any(1)


def torch_retain(imgs, cloud, mask):
    B, N, C, H, W = imgs.shape
    out = torch.ones_like(imgs) * 255
    valid = torch.where(mask.expand(-1, -1, -1, -1, 2), cloud, torch.tensor(-1.0, device=cloud.device))
    for b in range(B):
        for n in range(N):
            for z in range(cloud.shape[2]):
                c = valid[b, n, z]
                sel = c[c[..., 0] != -1].long().clamp(0, max(W, H) - 1)
                out[b, n, :, sel[:, 1], sel[:, 0]] = imgs[b, n, :, sel[:, 1], sel[:, 0]]
    return out


def main():
    B = int(os.environ.get("VC_B", "8"))
    pillars, imgs, mask = voxel_color_case(B, 6, 13, 16384, 256, 704, seed=1)
    tp, ti, tm = (torch.from_numpy(a).cuda() for a in (pillars, imgs, mask))
    res = {"B": B, "voxels_per_sample": 13 * 16384, "cameras": 6}
    res["ours_color_ms"] = timeit(lambda: color_voxels_from_images(tp, ti, tm, divisor=255.0))
    res["torch_color_ms"] = timeit(lambda: torch_color(tp, ti, tm)[0] / 255.0)
    res["ours_retain_ms"] = timeit(lambda: retain_valid_pixels(ti, tp, tm))
    res["torch_retain_ms"] = timeit(lambda: torch_retain(ti, tp, tm), reps=3, warm=1)
    # algorithmic bytes of the colouring: coords 8 + mask 1 per (camera, voxel), 12 B written per voxel
    n = B * 13 * 16384
    res["color_stream_GBps"] = (n * 6 * 9 + n * 12) / res["ours_color_ms"] / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
