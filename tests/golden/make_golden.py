"""Records golden vectors from the REFERENCE's own CUDA rasterizer (oracle/_ref/libinria_ref.so).

Run on a GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy the .npz files into tests/golden/.  Inputs are regenerated from the seeds stored in each
file (ocrfdet_b200.scenes), so the fixtures hold OUTPUTS only and stay small.
The CPU suite (tests/test_oracle_golden.py) checks the C oracle against them: this is what pins the
oracle to the reference's behaviour (the reference ships no test vectors of its own).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ocrfdet_b200.scenes import frustum_scene, ring_scene  # noqa: E402
from oracle import ref  # noqa: E402

CASES = [
    dict(name="frustum_p2000_160x96", kind="frustum", P=2000, seed=101, W=160, H=96, view=0, bg=[0.1, 0.2, 0.3]),
    dict(name="ring_p6000_176x64", kind="ring", P=6000, seed=102, W=176, H=64, view=1, bg=[0.0, 0.0, 0.0]),
    dict(name="frustum_p300_50x37", kind="frustum", P=300, seed=103, W=50, H=37, view=0, bg=[0.5, 0.25, 0.75]),
]


def scene(case):
    if case["kind"] == "frustum":
        g, cams = frustum_scene(P=case["P"], seed=case["seed"], width=case["W"], height=case["H"])
    else:
        g, cams = ring_scene(P=case["P"], seed=case["seed"], width=case["W"], height=case["H"], n_views=6)
    return g, cams[case["view"]]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for case in CASES:
        g, cam = scene(case)
        W, H = case["W"], case["H"]
        t = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
        cu = lambda a: torch.from_numpy(np.asarray(a, np.float32)).cuda()  # noqa: E731
        rr = ref.RefRasterizer()
        col, radii, n = rr.forward(t["means3D"], t["opacities"], t["colors"], cu(cam["viewmatrix"]), cu(cam["projmatrix"]),
                                   cu(cam["campos"]), W, H, cam["tanfovx"], cam["tanfovy"], cu(case["bg"]),
                                   scales=t["scales"], rotations=t["rotations"])
        st = rr.state()
        gcol = np.random.default_rng(case["seed"] + 7).normal(size=(3, H, W)).astype(np.float32)
        gr = rr.backward(t["means3D"], t["colors"], cu(cam["viewmatrix"]), cu(cam["projmatrix"]), cu(cam["campos"]),
                         cam["tanfovx"], cam["tanfovy"], cu(case["bg"]), radii, torch.from_numpy(gcol).cuda(),
                         scales=t["scales"], rotations=t["rotations"])
        torch.cuda.synchronize()
        c = lambda x: x.cpu().numpy()  # noqa: E731
        np.savez_compressed(
            os.path.join(outdir, case["name"] + ".npz"), case=np.array(repr(case)), num_rendered=np.int64(n),
            radii=c(radii), depths=c(st["depths"]), xy=c(st["xy"]), conic_opacity=c(st["conic_opacity"]),
            tiles_touched=c(st["tiles_touched"]), offsets=c(st["offsets"]), keys=c(st["keys"]), point_list=c(st["point_list"]),
            ranges=c(st["ranges"]), final_T=c(st["final_T"]), n_contrib=c(st["n_contrib"]), color=c(col),
            g_means3D=c(gr["means3D"]), g_scales=c(gr["scales"]), g_rotations=c(gr["rotations"]),
            g_opacities=c(gr["opacities"]), g_colors=c(gr["colors"]), g_means2D=c(gr["means2D"]), g_conic=c(gr["conic"]),
            gpu=np.array(torch.cuda.get_device_name(0)))
        print(case["name"], "num_rendered", n)
        rr.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
