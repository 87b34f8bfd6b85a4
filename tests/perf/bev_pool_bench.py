"""BEV pooling v2 at the OcRFDet training shape (8 samples x 6 cameras, D=88, 16x44 features, C=80, 128x128 BEV):
ours (C ABI through ocrfdet_b200.bev_pool) against the reference's own CUDA kernels (oracle/_ref/libbevpool_ref.so,
with the reference's Python regrouping before its backward).  python tests/perf/bev_pool_bench.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ocrfdet_b200.bev_pool import QuickCumsumCuda  # noqa: E402
from ocrfdet_b200.scenes import bev_pool_case  # noqa: E402
from oracle import ref  # noqa: E402


def timeit(fn, flush, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(os.environ.get("BEV_B", "8"))
    c = bev_pool_case(B=B, N=6, D=88, H=16, W=44, C=80, bev=128, seed=3)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items() if isinstance(v, np.ndarray)}
    depth, feat = t["depth"].requires_grad_(True), t["feat"].requires_grad_(True)
    og = torch.randn(c["bev_feat_shape"], device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    args = (t["ranks_depth"], t["ranks_feat"], t["ranks_bev"], c["bev_feat_shape"], t["interval_starts"], t["interval_lengths"])
    res = {"B": B, "points": int(t["ranks_bev"].numel()), "intervals": int(t["interval_starts"].numel()), "C": 80}

    def ours_fwd():
        with torch.no_grad():
            return QuickCumsumCuda.apply(depth, feat, *args)

    def ours_both():
        out = QuickCumsumCuda.apply(depth, feat, *args)
        out.backward(og)
        depth.grad = feat.grad = None

    res["ours_fwd_ms"] = timeit(ours_fwd, flush)
    res["ours_fwd_bwd_ms"] = timeit(ours_both, flush)
    if ref.bev_available():
        d0, f0 = depth.detach(), feat.detach()
        og2 = og.reshape(-1, 80)

        def ref_fwd():
            return ref.bev_pool_forward(d0, f0, t["ranks_depth"], t["ranks_feat"], t["ranks_bev"], c["n_bev"],
                                        t["interval_starts"], t["interval_lengths"])

        def ref_both():
            ref_fwd()
            ref.bev_pool_backward(og2, d0, f0, t["ranks_depth"], t["ranks_feat"], t["ranks_bev"], stable=False)

        res["ref_fwd_ms"] = timeit(ref_fwd, flush)
        res["ref_fwd_bwd_ms"] = timeit(ref_both, flush)
    # algorithmic bytes: every point gathers one feat row (fwd) / one out_grad row (bwd) + 12 B of ranks + 4 B depth
    n, C = res["points"], 80
    res["fwd_gather_GBps"] = n * (4 * C + 16) / res["ours_fwd_ms"] / 1e6
    res["bwd_gather_GBps"] = n * (4 * C + 16 + 12) / max(res["ours_fwd_bwd_ms"] - res["ours_fwd_ms"], 1e-6) / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
