"""Times the fused Gaussian-construction heads against the reference's torch formulation (4 small MLPs).
One sample of the OcRF voxel grid: feat [212992, 80] + rgb [212992, 3].  python tools/heads_bench.py"""
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrfdet_b200.gaussian_heads import GaussianHeads  # noqa: E402


class RefHead(nn.Module):
    def __init__(self, fin, out, act):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(fin, 4), nn.Linear(4, out), act

    def forward(self, x):
        return self.act(self.fc2(F.relu(self.fc1(x))))


def timeit(fn, flush, reps=30, warm=5):
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
    n, Fd = 212992, 80
    torch.manual_seed(0)
    dev = "cuda"
    acts = {"S_MLP": F.softplus, "R_MLP": lambda x: F.normalize(x, dim=-1), "A_MLP": torch.sigmoid, "C_MLP": torch.sigmoid}
    outs = {"S_MLP": 3, "R_MLP": 4, "A_MLP": 1, "C_MLP": 3}
    ref = nn.ModuleDict({h: RefHead(Fd + 3 if h == "C_MLP" else Fd, o, acts[h]) for h, o in outs.items()}).to(dev)
    m = GaussianHeads(Fd).to(dev)
    m.load_state_dict({k.replace(".act", ""): v for k, v in ref.state_dict().items()})
    m.packed.grad = None
    feat = torch.randn(n, Fd, device=dev, requires_grad=True)
    rgb = torch.rand(n, 3, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gs = [torch.randn(n, w, device=dev) for w in (1, 3, 4, 3)]

    def ref_fwd():
        return (ref["A_MLP"](feat), ref["S_MLP"](feat), ref["R_MLP"](feat), ref["C_MLP"](torch.cat((feat, rgb), -1)))

    def both(f):
        def run():
            o = f()
            torch.autograd.backward(o, gs)
            feat.grad = None
        return run

    def fused_fwd():
        return m(feat, rgb)

    res = {}
    with torch.no_grad():
        res["ref_fwd_ms"] = timeit(ref_fwd, flush)
        res["fused_fwd_ms"] = timeit(fused_fwd, flush)
    res["ref_fwd_bwd_ms"] = timeit(both(ref_fwd), flush)
    res["fused_fwd_bwd_ms"] = timeit(both(fused_fwd), flush)
    fwd_bytes = n * (Fd * 4 + 12 + 11 * 4 + 64)
    bwd_bytes = n * (Fd * 4 + 12 + 64 + 11 * 4 + Fd * 4)
    res["fused_fwd_GBps"] = fwd_bytes / res["fused_fwd_ms"] / 1e6
    res["fused_bwd_GBps"] = bwd_bytes / max(res["fused_fwd_bwd_ms"] - res["fused_fwd_ms"], 1e-6) / 1e6
    res["n"], res["F"] = n, Fd

    # kernels alone through the C ABI: 4 samples per call (273 MB of features, larger than L2), no flush needed
    from ocrfdet_b200 import _lib
    from ocrfdet_b200.gaussian_heads import _views
    L = _lib.lib()
    n4 = 4 * n
    f4, r4 = torch.randn(n4, Fd, device=dev), torch.rand(n4, 3, device=dev)
    v = _views(m.packed.data, Fd)
    outs = [torch.empty(n4, w, device=dev) for w in (1, 3, 4, 3, 16)]
    g4 = [torch.randn(n4, w, device=dev) for w in (1, 3, 4, 3)]
    gf = torch.empty_like(f4)
    gp = torch.zeros_like(m.packed.data)
    gv = _views(gp, Fd)
    P = _lib.ptr

    def k_fwd():
        _lib.check(L.ocrf_gaussian_heads_forward(_lib.current_stream(), n4, Fd, P(f4), P(r4), P(v["w1t"]), P(v["b1"]),
                                                 P(v["w2"]), P(v["b2"]), P(outs[0]), P(outs[1]), P(outs[2]), P(outs[3]),
                                                 P(outs[4])), "fwd")

    def k_bwd():
        _lib.check(L.ocrf_gaussian_heads_backward(_lib.current_stream(), n4, Fd, P(f4), P(r4), P(v["w1t"]), P(v["w2"]),
                                                  P(v["b2"]), P(outs[4]), P(g4[0]), P(g4[1]), P(g4[2]), P(g4[3]), P(gf),
                                                  P(gv["w1t"]), P(gv["b1"]), P(gv["w2"]), P(gv["b2"]), P(ws4)), "bwd")

    ws4 = torch.empty(L.ocrf_gaussian_heads_backward_workspace_bytes(n4), dtype=torch.uint8, device=dev)
    tiny = torch.empty(1, device=dev)
    res["kernel_fwd_ms_4samples"] = timeit(k_fwd, tiny)
    res["kernel_bwd_ms_4samples"] = timeit(k_bwd, tiny)
    res["kernel_fwd_GBps"] = 4 * fwd_bytes / res["kernel_fwd_ms_4samples"] / 1e6
    res["kernel_bwd_GBps"] = 4 * bwd_bytes / res["kernel_bwd_ms_4samples"] / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
