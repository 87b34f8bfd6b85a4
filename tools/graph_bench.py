"""Eager vs CUDA-graph replay of the bench step (BASELINE config 2): host issue time and device time per step."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrfdet_b200 import rasterizer as R  # noqa: E402
from ocrfdet_b200.graphs import GraphedRenderStep, INPUTS  # noqa: E402
from ocrfdet_b200.scenes import ring_scene  # noqa: E402

W, H, P, V = 704, 256, 100000, 6
g, cams = ring_scene(P=P, seed=1234, width=W, height=H, n_views=V)
dev = {k: torch.from_numpy(g[k]).unsqueeze(0).cuda().requires_grad_(True) for k in INPUTS}
cam_t = R.pack_camera_dicts(cams, "cuda")
bg = torch.zeros(3, device="cuda")
gcol, gop = torch.randn(V, 3, H, W, device="cuda"), torch.randn(V, 1, H, W, device="cuda")
cap = 5_000_000


def eager():
    for k in INPUTS:
        dev[k].grad = None
    c, r, d, o = R.render_batch(dev["means3D"], dev["opacities"], cam_t, H, W, bg, colors_precomp=dev["colors"],
                                scales=dev["scales"], rotations=dev["rotations"], pair_capacity=cap)
    torch.autograd.backward([c, o], [gcol, gop])


step = GraphedRenderStep(S=1, P=P, cams=cam_t, height=H, width=W, channels=3, pair_capacity=cap)
step.capture(grad_color=gcol, grad_opacity=gop, **{k: dev[k].detach() for k in INPUTS})


def measure(fn, n=100):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3, a.elapsed_time(b) / n


res = {}
res["eager_host_ms"], res["eager_device_ms"] = measure(eager)
res["graph_host_ms"], res["graph_device_ms"] = measure(step.graph.replay)
R.check_overflow()
print(json.dumps(res))
