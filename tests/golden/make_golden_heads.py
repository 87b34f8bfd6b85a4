"""Golden vectors for the Gaussian-construction heads, produced by the REFERENCE's own module code.

Run in the build container (needs /root/reference; CPU only):   python tests/golden/make_golden_heads.py
The four nn.Module classes are lifted out of view_transformer_ocrf.py with `ast` (the file itself cannot be
imported without mmcv/mmdet3d) and executed unmodified under torch; inputs, parameters, outputs and gradients
go to tests/golden/heads_*.npz.  tests/test_oracle_golden.py checks the C oracle against these files.
"""
import ast
import os
import sys

import numpy as np
import torch

REF = "/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py"
CLASSES = {"S_MLP": ("ScaleFactorMLP", 3), "R_MLP": ("RotationFactorMLP", 4), "A_MLP": ("OpacityFactorMLP", 1),
           "C_MLP": ("ColorFactorMLPGaussian", 3)}
CASES = [dict(name="heads_n257_f80", n=257, F=80, seed=7, scale=1.0),
         dict(name="heads_n64_f17", n=64, F=17, seed=8, scale=6.0)]  # large scale -> softplus threshold, dead ReLUs


def reference_classes():
    tree = ast.parse(open(REF).read())
    wanted = {c for c, _ in CLASSES.values()}
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in wanted]
    assert len(body) == len(wanted)
    ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    ns = reference_classes()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for case in CASES:
        torch.manual_seed(case["seed"])
        n, Fd = case["n"], case["F"]
        mods = {k: ns[cls](Fd, 4, out) for k, (cls, out) in CLASSES.items()}
        for m in mods.values():
            for p in m.parameters():
                p.data.mul_(case["scale"])
        feat = (torch.randn(n, Fd) * case["scale"]).requires_grad_(True)
        rgb = torch.rand(n, 3)
        # the call sequence of view_transformer_ocrf.py:1130-1133
        opacity = mods["A_MLP"](feat)
        scaling = mods["S_MLP"](feat)
        rotation = mods["R_MLP"](feat)
        color = mods["C_MLP"](torch.cat((feat, rgb), dim=-1))
        gs = [torch.randn_like(t) for t in (opacity, scaling, rotation, color)]
        (opacity * gs[0]).sum().add((scaling * gs[1]).sum()).add((rotation * gs[2]).sum()).add((color * gs[3]).sum()).backward()
        rec = dict(feat=feat.detach().numpy(), rgb=rgb.numpy(), opacity=opacity.detach().numpy(),
                   scaling=scaling.detach().numpy(), rotation=rotation.detach().numpy(), color=color.detach().numpy(),
                   g_opacity=gs[0].numpy(), g_scaling=gs[1].numpy(), g_rotation=gs[2].numpy(), g_color=gs[3].numpy(),
                   g_feat=feat.grad.numpy())
        for k, m in mods.items():
            for pn, p in m.named_parameters():
                rec["%s.%s" % (k, pn)] = p.detach().numpy()
                rec["g.%s.%s" % (k, pn)] = p.grad.numpy()
        np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"), **rec)
        print("wrote", case["name"], {k: v.shape for k, v in rec.items() if not k.startswith("g.")})


if __name__ == "__main__":
    sys.exit(main())
