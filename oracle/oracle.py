"""ctypes/numpy front end of the CPU oracle (oracle/ocrf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under ocrfdet_b200/ imports this module.

Every function mirrors one stage of the reference render path; see the header of
ocrf_oracle.c for the reference file:line each stage restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libocrf_oracle.so")
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C restatement with gcc (seconds)."""
    src = os.path.join(_HERE, "ocrf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ocrf_oracle_scan.restype = C.c_uint64
        _lib.ocrf_oracle_higher_msb.restype = C.c_uint32
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def higher_msb(n):
    return int(lib().ocrf_oracle_higher_msb(C.c_uint32(n)))


def tile_grid(W, H):
    return (W + 15) // 16, (H + 15) // 16


def preprocess(means, opacities, view, proj, W, H, tanfovx, tanfovy, scales=None, rots=None, scale_modifier=1.0,
               cov3D_precomp=None, shs=None, sh_degree=0, campos=None):
    means = _f32(means)
    P = means.shape[0]
    scales, rots, cov3D_precomp, shs = _f32(scales), _f32(rots), _f32(cov3D_precomp), _f32(shs)
    opacities = _f32(opacities).reshape(-1)
    view, proj = _f32(view).reshape(-1), _f32(proj).reshape(-1)
    campos = _f32(campos if campos is not None else np.zeros(3))
    out = dict(
        radii=np.zeros(P, np.int32), xy=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
        cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
        tiles_touched=np.zeros(P, np.uint32), rgb=np.zeros((P, 3), np.float32), clamped=np.zeros((P, 3), np.uint8))
    sh_M = 0 if shs is None else shs.shape[1]
    lib().ocrf_oracle_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(sh_M), _ptr(means), _ptr(scales), C.c_float(scale_modifier), _ptr(rots),
        _ptr(opacities), _ptr(shs), _ptr(cov3D_precomp), _ptr(view), _ptr(proj), _ptr(campos), C.c_int(W), C.c_int(H),
        C.c_float(tanfovx), C.c_float(tanfovy), _ptr(out["radii"]), _ptr(out["xy"]), _ptr(out["depths"]),
        _ptr(out["cov3D"]), _ptr(out["conic_opacity"]), _ptr(out["tiles_touched"]), _ptr(out["rgb"]),
        _ptr(out["clamped"]))
    if cov3D_precomp is not None:
        out["cov3D"] = cov3D_precomp.reshape(P, 6).copy()
    return out


def mark_visible(means, view, proj):
    means = _f32(means)
    P = means.shape[0]
    present = np.zeros(P, np.uint8)
    lib().ocrf_oracle_mark_visible(C.c_int(P), _ptr(means), _ptr(_f32(view).reshape(-1)), _ptr(_f32(proj).reshape(-1)),
                                   _ptr(present))
    return present.astype(bool)


def sort_pairs(keys, values, end_bit):
    keys = np.ascontiguousarray(keys, np.uint64)
    values = np.ascontiguousarray(values, np.uint32)
    ko, vo = np.zeros_like(keys), np.zeros_like(values)
    lib().ocrf_oracle_sort_pairs(C.c_uint64(keys.shape[0]), _ptr(keys), _ptr(values), _ptr(ko), _ptr(vo),
                                 C.c_int(end_bit))
    return ko, vo


def bin_tiles(xy, depths, radii, tiles_touched, W, H):
    """Scan + duplicateWithKeys + stable sort + identifyTileRanges."""
    P = radii.shape[0]
    gx, gy = tile_grid(W, H)
    offsets = np.zeros(P, np.uint32)
    N = int(lib().ocrf_oracle_scan(C.c_int(P), _ptr(np.ascontiguousarray(tiles_touched, np.uint32)), _ptr(offsets)))
    keys_u, vals_u = np.zeros(N, np.uint64), np.zeros(N, np.uint32)
    lib().ocrf_oracle_duplicate(C.c_int(P), _ptr(_f32(xy)), _ptr(_f32(depths)), _ptr(np.ascontiguousarray(radii, np.int32)),
                                _ptr(offsets), C.c_int(W), C.c_int(H), _ptr(keys_u), _ptr(vals_u))
    end_bit = 32 + higher_msb(gx * gy)
    keys, vals = sort_pairs(keys_u, vals_u, end_bit)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    lib().ocrf_oracle_tile_ranges(C.c_uint64(N), _ptr(keys), C.c_int(gx * gy), _ptr(ranges))
    return dict(N=N, offsets=offsets, keys_unsorted=keys_u, values_unsorted=vals_u, keys=keys, point_list=vals,
                ranges=ranges, end_bit=end_bit)


def tile_ranges(keys_sorted, tiles):
    keys_sorted = np.ascontiguousarray(keys_sorted, np.uint64)
    ranges = np.zeros((tiles, 2), np.uint32)
    lib().ocrf_oracle_tile_ranges(C.c_uint64(keys_sorted.shape[0]), _ptr(keys_sorted), C.c_int(tiles), _ptr(ranges))
    return ranges


def render_forward(W, H, ranges, point_list, xy, depths, conic_opacity, colors, bg, amb_eps=8e-6):
    colors = _f32(colors)
    Cc = colors.shape[1]
    out = dict(color=np.zeros((Cc, H, W), np.float32), depth=np.zeros((1, H, W), np.float32),
               opacity=np.zeros((1, H, W), np.float32), final_T=np.zeros((H, W), np.float32),
               n_contrib=np.zeros((H, W), np.uint32), ambiguous=np.zeros((H, W), np.uint8))
    lib().ocrf_oracle_render_forward(
        C.c_int(W), C.c_int(H), C.c_int(Cc), _ptr(np.ascontiguousarray(ranges, np.uint32)),
        _ptr(np.ascontiguousarray(point_list, np.uint32)), _ptr(_f32(xy)), _ptr(_f32(depths)), _ptr(_f32(conic_opacity)),
        _ptr(colors), _ptr(_f32(bg)), _ptr(out["color"]), _ptr(out["depth"]), _ptr(out["opacity"]), _ptr(out["final_T"]),
        _ptr(out["n_contrib"]), _ptr(out["ambiguous"]), C.c_float(amb_eps))
    return out


def render_backward(P, W, H, ranges, point_list, xy, conic_opacity, colors, bg, final_T, n_contrib, dL_dpix,
                    dL_dopacity_map=None):
    colors = _f32(colors)
    Cc = colors.shape[1]
    g = dict(mean2D=np.zeros((P, 2), np.float64), conic=np.zeros((P, 3), np.float64), opacity=np.zeros(P, np.float64),
             colors=np.zeros((P, Cc), np.float64))
    lib().ocrf_oracle_render_backward(
        C.c_int(P), C.c_int(W), C.c_int(H), C.c_int(Cc), _ptr(np.ascontiguousarray(ranges, np.uint32)),
        _ptr(np.ascontiguousarray(point_list, np.uint32)), _ptr(_f32(xy)), _ptr(_f32(conic_opacity)), _ptr(colors),
        _ptr(_f32(bg)), _ptr(_f32(final_T)), _ptr(np.ascontiguousarray(n_contrib, np.uint32)), _ptr(_f32(dL_dpix)),
        _ptr(_f32(dL_dopacity_map)), _ptr(g["mean2D"]), _ptr(g["conic"]), _ptr(g["opacity"]), _ptr(g["colors"]))
    return g


def preprocess_backward(means, radii, cov3D, view, proj, W, H, tanfovx, tanfovy, dL_dmean2D, dL_dconic, scales=None,
                        rots=None, scale_modifier=1.0, shs=None, sh_degree=0, clamped=None, campos=None,
                        dL_dcolor=None, f64=False):
    """f64=False: float32 arithmetic like the reference.  f64=True: the same chain in double from double
    screen-space gradients (the well-conditioned value both float32 implementations are judged against)."""
    means = _f32(means)
    P = means.shape[0]
    scales, rots, shs = _f32(scales), _f32(rots), _f32(shs)
    sh_M = 0 if shs is None else shs.shape[1]
    campos = _f32(campos if campos is not None else np.zeros(3))
    rt = np.float64 if f64 else np.float32
    _f32g = (lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=rt)))
    g = dict(means=np.zeros((P, 3), rt), cov3D=np.zeros((P, 6), rt),
             scales=np.zeros((P, 3), rt), rots=np.zeros((P, 4), rt),
             shs=None if shs is None else np.zeros(shs.shape, rt))
    fn = lib().ocrf_oracle_preprocess_backward_f64 if f64 else lib().ocrf_oracle_preprocess_backward
    fn(
        C.c_int(P), C.c_int(sh_degree), C.c_int(sh_M), _ptr(means), _ptr(np.ascontiguousarray(radii, np.int32)), _ptr(shs),
        _ptr(None if clamped is None else np.ascontiguousarray(clamped, np.uint8)), _ptr(scales),
        C.c_float(scale_modifier), _ptr(rots), _ptr(_f32(cov3D)), _ptr(_f32(view).reshape(-1)),
        _ptr(_f32(proj).reshape(-1)), _ptr(campos), C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy),
        _ptr(_f32g(dL_dmean2D)), _ptr(_f32g(dL_dconic)), _ptr(_f32g(dL_dcolor)), _ptr(g["means"]), _ptr(g["cov3D"]),
        _ptr(g["scales"] if scales is not None else None), _ptr(g["rots"] if scales is not None else None),
        _ptr(g["shs"]))
    return g


def rasterize(means, opacities, colors, view, proj, W, H, tanfovx, tanfovy, bg, scales=None, rots=None,
              scale_modifier=1.0, cov3D_precomp=None, shs=None, sh_degree=0, campos=None, amb_eps=8e-6):
    """Whole forward path: preprocess -> bin -> blend.  Returns (outputs, state)."""
    pre = preprocess(means, opacities, view, proj, W, H, tanfovx, tanfovy, scales=scales, rots=rots,
                     scale_modifier=scale_modifier, cov3D_precomp=cov3D_precomp, shs=shs, sh_degree=sh_degree,
                     campos=campos)
    b = bin_tiles(pre["xy"], pre["depths"], pre["radii"], pre["tiles_touched"], W, H)
    feats = pre["rgb"] if shs is not None else _f32(colors)
    out = render_forward(W, H, b["ranges"], b["point_list"], pre["xy"], pre["depths"], pre["conic_opacity"], feats, bg,
                         amb_eps=amb_eps)
    return out, dict(pre=pre, bin=b, feats=feats)


def rasterize_backward(state, means, view, proj, W, H, tanfovx, tanfovy, bg, out, dL_dcolor, dL_dopacity_map=None,
                       scales=None, rots=None, scale_modifier=1.0, shs=None, sh_degree=0, campos=None, f64=True):
    """Whole backward path.  Returns dict of gradients as the plugin returns them.

    f64=True (default) evaluates the per-Gaussian chain in double from the double-accumulated
    screen-space sums; f64=False rounds those sums to float32 and uses float32 arithmetic exactly like
    the reference's two preprocess-backward kernels."""
    pre, b, feats = state["pre"], state["bin"], state["feats"]
    P = pre["radii"].shape[0]
    g = render_backward(P, W, H, b["ranges"], b["point_list"], pre["xy"], pre["conic_opacity"], feats, bg,
                        out["final_T"], out["n_contrib"], dL_dcolor, dL_dopacity_map)
    pb = preprocess_backward(means, pre["radii"], pre["cov3D"], view, proj, W, H, tanfovx, tanfovy,
                             g["mean2D"], g["conic"], scales=scales, rots=rots,
                             scale_modifier=scale_modifier, shs=shs, sh_degree=sh_degree, clamped=pre["clamped"],
                             campos=campos, dL_dcolor=g["colors"], f64=f64)
    return dict(means3D=pb["means"], means2D=g["mean2D"], colors=g["colors"], opacities=g["opacity"],
                scales=pb["scales"], rotations=pb["rots"], cov3D=pb["cov3D"], shs=pb["shs"], conic=g["conic"])


def opacity_mask_forward(x, w, opacity_bev):
    x = _f32(x)
    B, Cc, H, W = x.shape
    w = _f32(w).reshape(2, -1)
    K = int(round(np.sqrt(w.shape[1])))
    out, mask, stats = np.zeros_like(x), np.zeros((B, 1, H, W), np.float32), np.zeros((B, 2, H, W), np.float32)
    lib().ocrf_oracle_opacity_mask_forward(C.c_int(B), C.c_int(Cc), C.c_int(H), C.c_int(W), C.c_int(K), _ptr(x), _ptr(w),
                                           _ptr(_f32(opacity_bev)), _ptr(out), _ptr(mask), _ptr(stats))
    return out, mask, stats


def opacity_mask_backward(x, w, mask, stats, g_out):
    x = _f32(x)
    B, Cc, H, W = x.shape
    w = _f32(w).reshape(2, -1)
    K = int(round(np.sqrt(w.shape[1])))
    gx, gw, gop = np.zeros_like(x), np.zeros((2, K, K), np.float32), np.zeros((B, 1, H, W), np.float32)
    lib().ocrf_oracle_opacity_mask_backward(C.c_int(B), C.c_int(Cc), C.c_int(H), C.c_int(W), C.c_int(K), _ptr(x), _ptr(w),
                                            _ptr(_f32(mask)), _ptr(_f32(stats)), _ptr(_f32(g_out)), _ptr(gx), _ptr(gw),
                                            _ptr(gop))
    return gx, gw, gop


HEAD_ORDER = ("S_MLP", "R_MLP", "A_MLP", "C_MLP")
HEAD_OUTS = (3, 4, 1, 3)


def _head_ptrs(params, key, shapes=None):
    arrs = [_f32(params[h][key]) for h in HEAD_ORDER]
    return arrs, (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])


def gaussian_heads_forward(feat, rgb, params):
    """params: {"S_MLP": {"fc1.weight": [4,F], "fc1.bias": [4], "fc2.weight": [3,4], "fc2.bias": [3]}, "R_MLP": ..,
    "A_MLP": .., "C_MLP": {"fc1.weight": [4,F+3], ..}} in nn.Linear layout (view_transformer_ocrf.py:272-320)."""
    feat, rgb = _f32(feat), _f32(rgb)
    n, F = feat.shape
    keep = [_head_ptrs(params, k) for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")]
    op, sc = np.zeros((n, 1), np.float32), np.zeros((n, 3), np.float32)
    rot, col = np.zeros((n, 4), np.float32), np.zeros((n, 3), np.float32)
    lib().ocrf_oracle_gaussian_heads_forward(C.c_longlong(n), C.c_int(F), _ptr(feat), _ptr(rgb), keep[0][1], keep[1][1],
                                             keep[2][1], keep[3][1], _ptr(op), _ptr(sc), _ptr(rot), _ptr(col))
    return op, sc, rot, col


def gaussian_heads_backward(feat, rgb, params, g_opacity, g_scales, g_rotations, g_colors):
    feat, rgb = _f32(feat), _f32(rgb)
    n, F = feat.shape
    keep = [_head_ptrs(params, k) for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")]
    grads = {h: {"fc1.weight": np.zeros((4, F + 3 if h == "C_MLP" else F), np.float32), "fc1.bias": np.zeros(4, np.float32),
                 "fc2.weight": np.zeros((o, 4), np.float32), "fc2.bias": np.zeros(o, np.float32)}
             for h, o in zip(HEAD_ORDER, HEAD_OUTS)}
    gp = [(C.c_void_p * 4)(*[grads[h][k].ctypes.data for h in HEAD_ORDER])
          for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")]
    g_feat = np.zeros((n, F), np.float32)
    gs = [_f32(g_opacity), _f32(g_scales), _f32(g_rotations), _f32(g_colors)]
    lib().ocrf_oracle_gaussian_heads_backward(C.c_longlong(n), C.c_int(F), _ptr(feat), _ptr(rgb), keep[0][1], keep[1][1],
                                              keep[2][1], keep[3][1], _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]),
                                              _ptr(g_feat), gp[0], gp[1], gp[2], gp[3])
    return g_feat, grads


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def bev_pool_forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, n_bev, interval_starts, interval_lengths):
    """out [n_bev, c] (channels-last, before the permute of bev_pool.py:86)."""
    depth, feat = _f32(depth).reshape(-1), _f32(feat)
    c = feat.shape[-1]
    feat = feat.reshape(-1, c)
    rd, rf, rb, st, ln = (_i32(a) for a in (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths))
    out = np.zeros((n_bev, c), np.float32)
    lib().ocrf_oracle_bev_pool_forward(C.c_int(c), C.c_int(len(st)), _ptr(depth), _ptr(feat), _ptr(rd), _ptr(rf), _ptr(rb),
                                       _ptr(st), _ptr(ln), _ptr(out))
    return out


def bev_pool_regroup(ranks_depth, ranks_feat, ranks_bev):
    """bev_pool.py:47-60 with a STABLE argsort (torch's default argsort leaves the order of equal keys open)."""
    rf = _i32(ranks_feat)
    order = np.argsort(rf, kind="stable")
    rf, rd, rb = rf[order], _i32(ranks_depth)[order], _i32(ranks_bev)[order]
    kept = np.ones(len(rf), bool)
    kept[1:] = rf[1:] != rf[:-1]
    starts = np.nonzero(kept)[0].astype(np.int32)
    lengths = np.diff(np.append(starts, len(rf))).astype(np.int32)
    return rd, rf, rb, starts, lengths


def bev_pool_backward(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev):
    """out_grad [n_bev, c] -> (depth_grad like depth, feat_grad like feat)."""
    depth_shape, feat_shape = np.shape(depth), np.shape(feat)
    depth, feat = _f32(depth).reshape(-1), _f32(feat)
    c = feat.shape[-1]
    feat = feat.reshape(-1, c)
    og = _f32(out_grad).reshape(-1, c)
    depth_grad, feat_grad = np.zeros_like(depth), np.zeros_like(feat)
    if len(ranks_feat):
        rd, rf, rb, st, ln = bev_pool_regroup(ranks_depth, ranks_feat, ranks_bev)
        lib().ocrf_oracle_bev_pool_backward(C.c_int(c), C.c_int(len(st)), _ptr(og), _ptr(depth), _ptr(feat), _ptr(rd),
                                            _ptr(rf), _ptr(rb), _ptr(st), _ptr(ln), _ptr(depth_grad), _ptr(feat_grad))
    return depth_grad.reshape(depth_shape), feat_grad.reshape(feat_shape)


def color_voxels(pillars, imgs, mask, divisor=1.0):
    """pillars [B,N,P,Q,2], imgs [B,N,C,H,W], mask [B,N,P,Q,1] -> avg [B,P,Q,C], valid [B,P,Q] (bool)."""
    pillars, imgs = _f32(pillars), _f32(imgs)
    B, N, P, Q, _ = pillars.shape
    _, _, Cc, H, W = imgs.shape
    m8 = np.ascontiguousarray(np.asarray(mask).reshape(B, N, P * Q) != 0, dtype=np.uint8)
    avg, valid = np.zeros((B, P, Q, Cc), np.float32), np.zeros((B, P, Q), np.uint8)
    lib().ocrf_oracle_color_voxels(C.c_int(B), C.c_int(N), C.c_longlong(P * Q), C.c_int(Cc), C.c_int(H), C.c_int(W),
                                   _ptr(pillars), _ptr(m8), _ptr(imgs), C.c_float(divisor), _ptr(avg), _ptr(valid))
    return avg, valid.astype(bool)


def retain_valid_pixels(image_matrix, cloud, mask, fill=255.0):
    img, cloud = _f32(image_matrix), _f32(cloud)
    B, N, Cc, H, W = img.shape
    M = int(np.prod(cloud.shape[2:-1]))
    m8 = np.ascontiguousarray(np.asarray(mask).reshape(B * N, M) != 0, dtype=np.uint8)
    out = np.zeros_like(img)
    lib().ocrf_oracle_retain_valid_pixels(C.c_int(B * N), C.c_longlong(M), C.c_int(Cc), C.c_int(H), C.c_int(W),
                                          _ptr(cloud), _ptr(m8), _ptr(img), C.c_float(fill), _ptr(out))
    return out
