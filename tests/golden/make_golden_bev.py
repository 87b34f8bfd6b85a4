"""Golden vectors for BEV pooling, recorded from the REFERENCE's own CUDA kernels
(oracle/_ref/libbevpool_ref.so = /root/reference/mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu + a C-ABI harness).

Run on a GPU box:   python tests/golden/make_golden_bev.py gpurun_out/golden
then copy the bevpool_*.npz files into tests/golden/.  Inputs are regenerated from the parameters stored in each file
(ocrfdet_b200.scenes.bev_pool_case), the fixtures hold the reference's outputs only.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ocrfdet_b200.scenes import bev_pool_case  # noqa: E402
from oracle import ref  # noqa: E402

CASES = [dict(name="bevpool_b1_n2_c80", B=1, N=2, D=24, H=8, W=22, C=80, bev=64, seed=11),
         dict(name="bevpool_b2_n3_c19", B=2, N=3, D=16, H=6, W=10, C=19, bev=32, seed=12)]


def case_inputs(case):
    kw = {k: v for k, v in case.items() if k != "name"}
    c = bev_pool_case(**kw)
    rng = np.random.default_rng(case["seed"] + 1000)
    c["out_grad"] = rng.normal(size=(c["n_bev"], case["C"])).astype(np.float32)
    return c


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for case in CASES:
        c = case_inputs(case)
        t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items() if isinstance(v, np.ndarray)}
        out = ref.bev_pool_forward(t["depth"], t["feat"], t["ranks_depth"], t["ranks_feat"], t["ranks_bev"], c["n_bev"],
                                   t["interval_starts"], t["interval_lengths"])
        dg, fg = ref.bev_pool_backward(t["out_grad"], t["depth"], t["feat"], t["ranks_depth"], t["ranks_feat"],
                                       t["ranks_bev"])
        np.savez_compressed(os.path.join(outdir, case["name"] + ".npz"), case=str(case), out=out.cpu().numpy(),
                            depth_grad=dg.cpu().numpy(), feat_grad=fg.cpu().numpy())
        print("wrote", case["name"], int(t["ranks_bev"].numel()), "points")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
