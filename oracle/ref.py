"""ctypes front end of oracle/_ref/libinria_ref.so -- the reference's own CUDA rasterizer, built
unmodified from the vendored sources by `make -C oracle ref` (needs /root/reference; the .so travels).

TEST / BENCH INFRASTRUCTURE ONLY.  Used (a) as the bit-exact comparator for keys / sort order /
tile ranges / colour / gradients on the GPU, (b) to record the golden vectors under tests/golden/,
(c) as `bench.py --impl reference`.  It has no depth or opacity output (the w-depth fork's source is
absent from the reference tree).  Runs on the legacy default stream, like the reference.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libinria_ref.so")
_lib = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        L.ref_ctx_create.restype = C.c_void_p
        L.ref_ctx_destroy.argtypes = [C.c_void_p]
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        L.ref_forward.argtypes = [vp, i, i, i, vp, i, i, vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, i, vp, vp]
        L.ref_backward.argtypes = [vp, i, i, i, i, vp, i, i, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, vp, vp, vp, vp,
                                   vp, vp, vp, vp, vp, vp, vp]
        L.ref_mark_visible.argtypes = [i, vp, vp, vp, vp]
        L.ref_get_state.argtypes = [vp] + [vp] * 13
        _lib = L
    return _lib


def _p(t):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


class RefRasterizer:
    """One context = the three grow-only buffers of one forward/backward pair."""

    def __init__(self):
        self.h = C.c_void_p(lib().ref_ctx_create())
        self.meta = None

    def close(self):
        if self.h:
            lib().ref_ctx_destroy(self.h)
            self.h = None

    def forward(self, means3D, opacities, colors, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, bg,
                scales=None, rotations=None, scale_modifier=1.0, cov3D_precomp=None, shs=None, sh_degree=0,
                prefiltered=False):
        P = means3D.shape[0]
        dev = means3D.device
        M = 0 if shs is None else shs.shape[1]
        out = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        n = lib().ref_forward(self.h, P, sh_degree, M, _p(bg), W, H, _p(means3D), _p(shs), _p(colors), _p(opacities),
                              _p(scales), scale_modifier, _p(rotations), _p(cov3D_precomp), _p(viewmatrix),
                              _p(projmatrix), _p(campos), tanfovx, tanfovy, int(prefiltered), _p(out), _p(radii))
        if n < 0:
            raise RuntimeError("reference forward failed")
        self.meta = dict(P=P, W=W, H=H, R=n, M=M, D=sh_degree)
        return out, radii, n

    def backward(self, means3D, colors, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg, radii, dL_dpix,
                 scales=None, rotations=None, scale_modifier=1.0, cov3D_precomp=None, shs=None):
        m = self.meta
        P, dev = m["P"], means3D.device
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        g = dict(means2D=z(P, 3), conic=z(P, 2, 2), opacities=z(P, 1), colors=z(P, 3), means3D=z(P, 3), cov3D=z(P, 6),
                 shs=z(P, max(m["M"], 1), 3), scales=z(P, 3), rotations=z(P, 4))
        rc = lib().ref_backward(self.h, P, m["D"], m["M"], m["R"], _p(bg), m["W"], m["H"], _p(means3D), _p(shs),
                                _p(colors), _p(scales), scale_modifier, _p(rotations), _p(cov3D_precomp),
                                _p(viewmatrix), _p(projmatrix), _p(campos), tanfovx, tanfovy, _p(radii),
                                _p(dL_dpix.contiguous()), _p(g["means2D"]), _p(g["conic"]), _p(g["opacities"]),
                                _p(g["colors"]), _p(g["means3D"]), _p(g["cov3D"]), _p(g["shs"]), _p(g["scales"]),
                                _p(g["rotations"]))
        if rc != 0:
            raise RuntimeError("reference backward failed")
        return g

    # ---- allocation-free variants for the "kernel-only" timing of bench.py --impl reference ----
    def alloc_io(self, P, W, H, device, M=1):
        """Outputs of one forward/backward pair, allocated ONCE: (out_color, radii, flat gradient buffer, views)."""
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)  # noqa: E731
        sizes = dict(means2D=(P, 3), conic=(P, 2, 2), opacities=(P, 1), colors=(P, 3), means3D=(P, 3), cov3D=(P, 6),
                     shs=(P, max(M, 1), 3), scales=(P, 3), rotations=(P, 4))
        flat = z(sum(int(torch.tensor(v).prod()) for v in sizes.values()))
        g, off = {}, 0
        for k, shp in sizes.items():
            n = int(torch.tensor(shp).prod())
            g[k] = flat[off:off + n].view(shp)
            off += n
        return z(3, H, W), torch.zeros((P,), dtype=torch.int32, device=device), flat, g

    def forward_into(self, out, radii, means3D, opacities, colors, viewmatrix, projmatrix, campos, W, H, tanfovx,
                     tanfovy, bg, scales, rotations):
        P = means3D.shape[0]
        n = lib().ref_forward(self.h, P, 0, 0, _p(bg), W, H, _p(means3D), None, _p(colors), _p(opacities), _p(scales),
                              1.0, _p(rotations), None, _p(viewmatrix), _p(projmatrix), _p(campos), tanfovx, tanfovy, 0,
                              _p(out), _p(radii))
        if n < 0:
            raise RuntimeError("reference forward failed")
        self.meta = dict(P=P, W=W, H=H, R=n, M=0, D=0)
        return n

    def backward_into(self, flat, g, means3D, colors, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg, radii,
                      dL_dpix, scales, rotations):
        m = self.meta
        flat.zero_()  # ONE fill for the nine accumulators (the binding issues nine torch::zeros, rasterize_points.cu:151-159)
        rc = lib().ref_backward(self.h, m["P"], 0, 0, m["R"], _p(bg), m["W"], m["H"], _p(means3D), None, _p(colors),
                                _p(scales), 1.0, _p(rotations), None, _p(viewmatrix), _p(projmatrix), _p(campos),
                                tanfovx, tanfovy, _p(radii), _p(dL_dpix), _p(g["means2D"]), _p(g["conic"]),
                                _p(g["opacities"]), _p(g["colors"]), _p(g["means3D"]), _p(g["cov3D"]), _p(g["shs"]),
                                _p(g["scales"]), _p(g["rotations"]))
        if rc != 0:
            raise RuntimeError("reference backward failed")

    def state(self):
        """Internal buffers of the last forward: keys, point list, ranges, ... (device tensors)."""
        m = self.meta
        P, R, W, H = m["P"], m["R"], m["W"], m["H"]
        dev = "cuda"
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        s = dict(depths=torch.zeros(P, dtype=torch.float32, device=dev),
                 xy=torch.zeros((P, 2), dtype=torch.float32, device=dev),
                 cov3D=torch.zeros((P, 6), dtype=torch.float32, device=dev),
                 conic_opacity=torch.zeros((P, 4), dtype=torch.float32, device=dev),
                 tiles_touched=torch.zeros(P, dtype=torch.int32, device=dev),
                 offsets=torch.zeros(P, dtype=torch.int32, device=dev),
                 keys_unsorted=torch.zeros(R, dtype=torch.int64, device=dev),
                 values_unsorted=torch.zeros(R, dtype=torch.int32, device=dev),
                 keys=torch.zeros(R, dtype=torch.int64, device=dev),
                 point_list=torch.zeros(R, dtype=torch.int32, device=dev),
                 ranges=torch.zeros((tiles, 2), dtype=torch.int32, device=dev),
                 final_T=torch.zeros((H, W), dtype=torch.float32, device=dev),
                 n_contrib=torch.zeros((H, W), dtype=torch.int32, device=dev))
        order = ["depths", "xy", "cov3D", "conic_opacity", "tiles_touched", "offsets", "keys_unsorted",
                 "values_unsorted", "keys", "point_list", "ranges", "final_T", "n_contrib"]
        rc = lib().ref_get_state(self.h, *[_p(s[k]) for k in order])
        if rc != 0:
            raise RuntimeError("reference state read-back failed")
        torch.cuda.synchronize()
        return s


def mark_visible(means3D, viewmatrix, projmatrix):
    P = means3D.shape[0]
    present = torch.zeros(P, dtype=torch.bool, device=means3D.device)
    lib().ref_mark_visible(P, _p(means3D), _p(viewmatrix), _p(projmatrix), _p(present))
    return present


# ---- the reference's bev_pool_v2 kernels (oracle/_ref/libbevpool_ref.so, `make -C oracle refbev`) ----
SO_BEV = os.path.join(_HERE, "_ref", "libbevpool_ref.so")
_lib_bev = None


def bev_available():
    return os.path.exists(SO_BEV)


def _bev():
    global _lib_bev
    if _lib_bev is None:
        L = C.CDLL(SO_BEV)
        L.ref_bev_pool_forward.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 8
        L.ref_bev_pool_backward.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 10
        _lib_bev = L
    return _lib_bev


def bev_pool_forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, n_bev, interval_starts, interval_lengths):
    """CUDA tensors in, out [n_bev, c] (bev_pool.py:18-44 without the permute).  Legacy default stream."""
    c = feat.shape[-1]
    out = feat.new_zeros((n_bev, c))
    torch.cuda.synchronize()
    rc = _bev().ref_bev_pool_forward(c, interval_starts.numel(), _p(depth), _p(feat), _p(ranks_depth), _p(ranks_feat),
                                     _p(ranks_bev), _p(interval_starts), _p(interval_lengths), _p(out))
    torch.cuda.synchronize()
    assert rc == 0, rc
    return out


def bev_pool_backward(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, stable=True):
    """bev_pool.py:46-79: regroup by ranks_feat (stable order here, see oracle.bev_pool_regroup), then the kernel."""
    order = torch.sort(ranks_feat.long(), stable=stable)[1]
    rf, rd, rb = ranks_feat[order].contiguous(), ranks_depth[order].contiguous(), ranks_bev[order].contiguous()
    kept = torch.ones(rb.shape[0], device=rb.device, dtype=torch.bool)
    kept[1:] = rf[1:] != rf[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = rb.shape[0] - starts[-1]
    depth_grad, feat_grad = torch.zeros_like(depth), torch.zeros_like(feat)
    c = feat.shape[-1]
    torch.cuda.synchronize()
    rc = _bev().ref_bev_pool_backward(c, starts.numel(), _p(out_grad.contiguous()), _p(depth), _p(feat), _p(rd), _p(rf),
                                      _p(rb), _p(starts.contiguous()), _p(lengths.contiguous()), _p(depth_grad),
                                      _p(feat_grad))
    torch.cuda.synchronize()
    assert rc == 0, rc
    return depth_grad, feat_grad
