"""CPU oracle (numpy, float64) of the height-aware opacity (HOA) lift -- stage 5 of the path.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke, bench cpu legs).  Nothing under ocrfdet_b200/ imports it.

Restates, with closed-form backward passes (checked against the reference modules' autograd through the golden
vectors of tests/golden/make_golden_hoa.py):

  lift        /root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:1159-1161
              opacity_up = interpolate(opacity, 128 -> 21, bilinear, align_corners=True), same for alpha_lidar;
              opacity_alpha = interpolate(DeformableAttention2D(opacity_up, alpha_up), 21 -> 128) + opacity
  attention   /root/reference/mmdet3d/ops/cross_attention_2d.py:93-220 in OcRFDet's configuration
              (view_transformer_ocrf.py:639-648: dim 13, dim_head 8, heads 1, one offset group, downsample 4,
              offset kernel 6, offset scale 4; dropout is an explicit keep-mask here)
  converter   OpacityVoxelToBEVConverter + HeightAttention, view_transformer_ocrf.py:421-518, batch norm in
              training mode (batch statistics, biased variance) or eval mode (running statistics)

Parameter dictionaries use the reference's state_dict names.
"""
import numpy as np

F64 = np.float64


# ------------------------------------------------------------------------------------------------------------------
# bilinear resize, align_corners=True (ATen upsample_bilinear2d: src = dst * (in - 1) / (out - 1))
# ------------------------------------------------------------------------------------------------------------------
def _ac_taps(n_in, n_out):
    """Source taps and weights exactly as ATen forms them for float32 tensors (UpSample.h: the scale, the source index
    and both lambdas are float32 expressions); at index ~127 one float32 ulp of the source position is 8e-6, so a
    float64 restatement of the same formula would sit 4e-6 away from the reference."""
    f32 = np.float32
    scale = f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0.0)
    src = (scale * np.arange(n_out).astype(f32)).astype(f32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    w1 = (src - i0.astype(f32)).astype(f32)
    w0 = (f32(1.0) - w1).astype(f32)
    return i0, i1, w0.astype(F64), w1.astype(F64)


def resize_ac(x, out_h, out_w):
    """x [..., H, W] -> [..., out_h, out_w]."""
    y0, y1, wy0, wy1 = _ac_taps(x.shape[-2], out_h)
    x0, x1, wx0, wx1 = _ac_taps(x.shape[-1], out_w)
    rows = x[..., y0, :] * wy0[:, None] + x[..., y1, :] * wy1[:, None]
    return rows[..., :, x0] * wx0 + rows[..., :, x1] * wx1


def resize_ac_backward(g, in_h, in_w):
    """Adjoint of resize_ac: g [..., out_h, out_w] -> [..., in_h, in_w]."""
    out_h, out_w = g.shape[-2:]
    y0, y1, wy0, wy1 = _ac_taps(in_h, out_h)
    x0, x1, wx0, wx1 = _ac_taps(in_w, out_w)
    lead = g.shape[:-2]
    rows = np.zeros(lead + (out_h, in_w), F64)
    np.add.at(rows, (Ellipsis, x0), g * wx0)
    np.add.at(rows, (Ellipsis, x1), g * wx1)
    out = np.zeros(lead + (in_h, in_w), F64)
    np.add.at(out, (Ellipsis, y0, slice(None)), rows * wy0[:, None])
    np.add.at(out, (Ellipsis, y1, slice(None)), rows * wy1[:, None])
    return out


# ------------------------------------------------------------------------------------------------------------------
# DeformableAttention2D (heads = 1, one offset group)
# ------------------------------------------------------------------------------------------------------------------
def _gelu(x):
    from math import sqrt
    from scipy.special import erf
    return 0.5 * x * (1.0 + erf(x / sqrt(2.0)))


def _gelu_grad(x):
    from math import pi, sqrt
    from scipy.special import erf
    return 0.5 * (1.0 + erf(x / sqrt(2.0))) + x * np.exp(-0.5 * x * x) / sqrt(2.0 * pi)


def _p(params, name):
    return np.asarray(params[name], F64)


def attention_forward(params, xq, xkv, keep=None, downsample=4, offset_scale=4.0, ksize=6):
    """xq, xkv [B, dim, h, w] -> y [B, dim, h, w] and a cache for the backward.
    keep: optional dropout keep-mask [B, h*w, hk*wk], already divided by (1 - p)."""
    xq, xkv = np.asarray(xq, F64), np.asarray(xkv, F64)
    B, dim, h, w = xq.shape
    Wq = _p(params, "to_q.weight")[:, :, 0, 0]
    Wk = _p(params, "to_k.weight")[:, :, 0, 0]
    Wv = _p(params, "to_v.weight")[:, :, 0, 0]
    Wo, bo = _p(params, "to_out.weight")[:, :, 0, 0], _p(params, "to_out.bias")
    wdw, bdw = _p(params, "to_offsets.0.weight")[:, 0], _p(params, "to_offsets.0.bias")
    wpw = _p(params, "to_offsets.2.weight")[:, :, 0, 0]
    inner = Wq.shape[0]
    scale = inner ** -0.5  # dim_head ** -0.5 with one head
    pad = (ksize - downsample) // 2
    hk, wk = (h + 2 * pad - ksize) // downsample + 1, (w + 2 * pad - ksize) // downsample + 1

    q = np.einsum("oc,bchw->bohw", Wq, xq)  # cross_attention_2d.py:155
    # to_offsets (:133-139): depthwise k x k stride-4 conv, GELU, 1x1 conv to 2, tanh, * offset_scale
    qp = np.pad(q, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    dw = np.zeros((B, inner, hk, wk), F64)
    for ky in range(ksize):
        for kx in range(ksize):
            dw += qp[:, :, ky:ky + downsample * hk:downsample, kx:kx + downsample * wk:downsample] * wdw[None, :, ky, kx, None, None]
    dw += bdw[None, :, None, None]
    ge = _gelu(dw)
    o2 = np.einsum("oc,bchw->bohw", wpw, ge)
    th = np.tanh(o2)
    off = th * offset_scale
    gx, gy = np.meshgrid(np.arange(wk, dtype=F64), np.arange(hk, dtype=F64))  # create_grid_like (:21-30): [0] = x, [1] = y
    vgrid = np.stack([gx, gy])[None] + off
    # normalize_grid (:32-41) with dim = 1: channel 0 is divided by (h - 1), channel 1 by (w - 1) OF THE OFFSET MAP
    vn = np.stack([2.0 * vgrid[:, 0] / max(hk - 1, 1) - 1.0, 2.0 * vgrid[:, 1] / max(wk - 1, 1) - 1.0], -1)  # [B,hk,wk,2]
    # grid_sample(x_kv, vn, bilinear, zeros, align_corners=False) (:176-179)
    ix = ((vn[..., 0] + 1.0) * w - 1.0) / 2.0
    iy = ((vn[..., 1] + 1.0) * h - 1.0) / 2.0
    x0, y0 = np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)
    fx, fy = ix - x0, iy - y0
    kvf = np.zeros((B, dim, hk, wk), F64)
    taps = []
    for dy_, wy in ((0, 1.0 - fy), (1, fy)):
        for dx_, wx in ((0, 1.0 - fx), (1, fx)):
            yy, xx = y0 + dy_, x0 + dx_
            ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
            yc, xc = np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)
            val = xkv[np.arange(B)[:, None, None], :, yc, xc]  # [B,hk,wk,dim]
            val = np.where(ok[..., None], val, 0.0)
            kvf += np.moveaxis(val * (wy * wx)[..., None], -1, 1)
            taps.append((yc, xc, ok, wy, wx, dy_, dx_))
    k = np.einsum("oc,bchw->bohw", Wk, kvf).reshape(B, inner, hk * wk)
    v = np.einsum("oc,bchw->bohw", Wv, kvf).reshape(B, inner, hk * wk)
    qs = (q * scale).reshape(B, inner, h * w)
    sim = np.einsum("bdi,bdj->bij", qs, k)
    # CPB (:52-88): queries on the normalised x_kv grid, keys at the normalised sampling positions
    qx, qy = np.meshgrid(np.arange(w, dtype=F64), np.arange(h, dtype=F64))
    # normalize_grid(grid, dim=0) unbinds (x, y) and divides x by (h - 1), y by (w - 1); h == w in OcRFDet and the
    # oracle keeps the reference's pairing:
    gq = np.stack([2.0 * qx / max(h - 1, 1) - 1.0, 2.0 * qy / max(w - 1, 1) - 1.0], -1).reshape(h * w, 2)
    gk = vn.reshape(B, hk * wk, 2)
    pos = gq[None, :, None, :] - gk[:, None, :, :]  # [B, i, j, 2]
    bb = np.sign(pos) * np.log(np.abs(pos) + 1.0)
    W1, b1 = _p(params, "rel_pos_bias.mlp.0.0.weight"), _p(params, "rel_pos_bias.mlp.0.0.bias")
    W2, b2 = _p(params, "rel_pos_bias.mlp.1.0.weight"), _p(params, "rel_pos_bias.mlp.1.0.bias")
    W3, b3 = _p(params, "rel_pos_bias.mlp.2.weight"), _p(params, "rel_pos_bias.mlp.2.bias")
    z1 = bb @ W1.T + b1
    h1 = np.maximum(z1, 0.0)
    z2 = h1 @ W2.T + b2
    h2 = np.maximum(z2, 0.0)
    bias = (h2 @ W3.T + b3)[..., 0]
    s = sim + bias
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    attn = e / e.sum(-1, keepdims=True)
    attn_d = attn if keep is None else attn * np.asarray(keep, F64)
    out = np.einsum("bij,bdj->bdi", attn_d, v).reshape(B, inner, h, w)
    y = np.einsum("oc,bchw->bohw", Wo, out) + bo[None, :, None, None]
    cache = dict(xq=xq, xkv=xkv, q=q, qp_shape=qp.shape, dw=dw, ge=ge, th=th, vgrid=vgrid, vn=vn, taps=taps, fx=fx, fy=fy,
                 x0=x0, y0=y0, kvf=kvf, k=k, v=v, qs=qs, pos=pos, bb=bb, z1=z1, h1=h1, z2=z2, h2=h2, attn=attn, attn_d=attn_d,
                 keep=keep, out=out, dims=(B, dim, h, w, hk, wk, inner, scale, pad, downsample, offset_scale, ksize))
    return y, cache


def attention_backward(params, cache, g_y):
    """-> (g_xq, g_xkv, parameter gradients by state_dict name)."""
    c = cache
    B, dim, h, w, hk, wk, inner, scale, pad, ds, offset_scale, ksize = c["dims"]
    g_y = np.asarray(g_y, F64)
    Wq = _p(params, "to_q.weight")[:, :, 0, 0]
    Wk = _p(params, "to_k.weight")[:, :, 0, 0]
    Wv = _p(params, "to_v.weight")[:, :, 0, 0]
    Wo = _p(params, "to_out.weight")[:, :, 0, 0]
    wdw = _p(params, "to_offsets.0.weight")[:, 0]
    wpw = _p(params, "to_offsets.2.weight")[:, :, 0, 0]
    W1 = _p(params, "rel_pos_bias.mlp.0.0.weight")
    W2 = _p(params, "rel_pos_bias.mlp.1.0.weight")
    W3 = _p(params, "rel_pos_bias.mlp.2.weight")
    G = {}
    G["to_out.bias"] = g_y.sum((0, 2, 3))
    G["to_out.weight"] = np.einsum("bohw,bchw->oc", g_y, c["out"])[:, :, None, None]
    g_out = np.einsum("oc,bohw->bchw", Wo, g_y).reshape(B, inner, h * w)
    g_attn_d = np.einsum("bdi,bdj->bij", g_out, c["v"])
    g_v = np.einsum("bij,bdi->bdj", c["attn_d"], g_out)
    g_attn = g_attn_d if c["keep"] is None else g_attn_d * np.asarray(c["keep"], F64)
    attn = c["attn"]
    g_s = attn * (g_attn - (attn * g_attn).sum(-1, keepdims=True))
    g_qs = np.einsum("bij,bdj->bdi", g_s, c["k"])
    g_k = np.einsum("bij,bdi->bdj", g_s, c["qs"])
    # CPB
    g_bias = g_s[..., None]
    G["rel_pos_bias.mlp.2.bias"] = g_bias.sum((0, 1, 2))
    G["rel_pos_bias.mlp.2.weight"] = np.einsum("bijo,bijk->ok", g_bias, c["h2"])
    g_h2 = g_bias @ W3
    g_z2 = g_h2 * (c["z2"] > 0)
    G["rel_pos_bias.mlp.1.0.bias"] = g_z2.sum((0, 1, 2))
    G["rel_pos_bias.mlp.1.0.weight"] = np.einsum("bijo,bijk->ok", g_z2, c["h1"])
    g_h1 = g_z2 @ W2
    g_z1 = g_h1 * (c["z1"] > 0)
    G["rel_pos_bias.mlp.0.0.bias"] = g_z1.sum((0, 1, 2))
    G["rel_pos_bias.mlp.0.0.weight"] = np.einsum("bijo,bijk->ok", g_z1, c["bb"])
    g_bb = g_z1 @ W1
    g_pos = g_bb / (np.abs(c["pos"]) + 1.0)  # d/dpos of sign(pos) log(|pos| + 1)
    g_vn = -g_pos.sum(1).reshape(B, hk, wk, 2)
    # k, v projections
    kvf2 = c["kvf"].reshape(B, dim, hk * wk)
    G["to_k.weight"] = np.einsum("boj,bcj->oc", g_k, kvf2)[:, :, None, None]
    G["to_v.weight"] = np.einsum("boj,bcj->oc", g_v, kvf2)[:, :, None, None]
    g_kvf = (np.einsum("oc,boj->bcj", Wk, g_k) + np.einsum("oc,boj->bcj", Wv, g_v)).reshape(B, dim, hk, wk)
    # grid_sample backward
    g_xkv = np.zeros((B, dim, h, w), F64)
    g_ix = np.zeros((B, hk, wk), F64)
    g_iy = np.zeros((B, hk, wk), F64)
    gk_last = np.moveaxis(g_kvf, 1, -1)  # [B,hk,wk,dim]
    bidx = np.arange(B)[:, None, None]
    fx, fy = c["fx"], c["fy"]
    for (yc, xc, ok, wy, wx, dy_, dx_) in c["taps"]:
        contrib = np.where(ok[..., None], gk_last * (wy * wx)[..., None], 0.0)
        np.add.at(g_xkv, (bidx, slice(None), yc, xc), contrib)
        val = np.where(ok[..., None], c["xkv"][bidx, :, yc, xc], 0.0)
        dot = (val * gk_last).sum(-1)
        g_ix += dot * wy * (1.0 if dx_ else -1.0)
        g_iy += dot * wx * (1.0 if dy_ else -1.0)
    g_vn[..., 0] += g_ix * (w / 2.0)
    g_vn[..., 1] += g_iy * (h / 2.0)
    g_vgrid = np.stack([g_vn[..., 0] * 2.0 / max(hk - 1, 1), g_vn[..., 1] * 2.0 / max(wk - 1, 1)], 1)
    g_o2 = g_vgrid * offset_scale * (1.0 - c["th"] ** 2)
    G["to_offsets.2.weight"] = np.einsum("bohw,bchw->oc", g_o2, c["ge"])[:, :, None, None]
    g_ge = np.einsum("oc,bohw->bchw", wpw, g_o2)
    g_dw = g_ge * _gelu_grad(c["dw"])
    G["to_offsets.0.bias"] = g_dw.sum((0, 2, 3))
    qp = np.pad(c["q"], ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    g_qp = np.zeros_like(qp)
    g_wdw = np.zeros_like(wdw)
    for ky in range(ksize):
        for kx in range(ksize):
            sl = (slice(None), slice(None), slice(ky, ky + ds * hk, ds), slice(kx, kx + ds * wk, ds))
            g_wdw[:, ky, kx] = (qp[sl] * g_dw).sum((0, 2, 3))
            g_qp[sl] += g_dw * wdw[None, :, ky, kx, None, None]
    G["to_offsets.0.weight"] = g_wdw[:, None]
    g_q = g_qs.reshape(B, inner, h, w) * scale + g_qp[:, :, pad:pad + h, pad:pad + w]
    G["to_q.weight"] = np.einsum("bohw,bchw->oc", g_q, c["xq"])[:, :, None, None]
    g_xq = np.einsum("oc,bohw->bchw", Wq, g_q)
    return g_xq, g_xkv, G


def lift_forward(params, opacity, alpha, keep=None):
    """view_transformer_ocrf.py:1159-1161: -> (opacity_alpha [B,Hh,S,S], cache)."""
    opacity, alpha = np.asarray(opacity, F64), np.asarray(alpha, F64)
    S_h, S_w = opacity.shape[-2:]
    ch, cw = int(S_h / 6), int(S_w / 6)
    o_up, a_up = resize_ac(opacity, ch, cw), resize_ac(alpha, ch, cw)
    y, cache = attention_forward(params, o_up, a_up, keep=keep)
    out = resize_ac(y, S_h, S_w) + opacity
    cache.update(full=(S_h, S_w), coarse=(ch, cw), opacity_up=o_up, alpha_up=a_up, att=y)
    return out, cache


def lift_backward(params, cache, g):
    g = np.asarray(g, F64)
    S_h, S_w = cache["full"]
    ch, cw = cache["coarse"]
    g_y = resize_ac_backward(g, ch, cw)
    g_xq, g_xkv, G = attention_backward(params, cache, g_y)
    g_opacity = g + resize_ac_backward(g_xq, S_h, S_w)
    g_alpha = resize_ac_backward(g_xkv, S_h, S_w)
    return g_opacity, g_alpha, G


# ------------------------------------------------------------------------------------------------------------------
# OpacityVoxelToBEVConverter (+ HeightAttention)
# ------------------------------------------------------------------------------------------------------------------
BN_EPS = 1e-5
BLOCKS = ("encoder1", "encoder2", "bottleneck", "decoder2", "decoder1")
GATES = {"encoder1": "ca1", "encoder2": "ca2", "bottleneck": "ca_bottleneck", "decoder2": "ca_dec2", "decoder1": "ca_dec1"}


def _dw3(x, wgt, b):
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    H, W = x.shape[-2:]
    out = np.zeros_like(x)
    for ky in range(3):
        for kx in range(3):
            out += xp[:, :, ky:ky + H, kx:kx + W] * wgt[None, :, 0, ky, kx, None, None]
    return out + b[None, :, None, None]


def _dw3_backward(x, wgt, g):
    H, W = x.shape[-2:]
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    gp = np.zeros_like(xp)
    gw = np.zeros_like(wgt)
    for ky in range(3):
        for kx in range(3):
            gw[:, 0, ky, kx] = (xp[:, :, ky:ky + H, kx:kx + W] * g).sum((0, 2, 3))
            gp[:, :, ky:ky + H, kx:kx + W] += g * wgt[None, :, 0, ky, kx, None, None]
    return gp[:, :, 1:1 + H, 1:1 + W], gw, g.sum((0, 2, 3))


def _block_forward(P, name, x, train, buffers):
    """conv_block (:483-489): depthwise 3x3, 1x1, batch norm, ReLU."""
    wd, bd = _p(P, name + ".0.weight"), _p(P, name + ".0.bias")
    wp, bp = _p(P, name + ".1.weight")[:, :, 0, 0], _p(P, name + ".1.bias")
    gam, bet = _p(P, name + ".2.weight"), _p(P, name + ".2.bias")
    d = _dw3(x, wd, bd)
    t = np.einsum("oc,bchw->bohw", wp, d) + bp[None, :, None, None]
    if train:
        mean, var = t.mean((0, 2, 3)), t.var((0, 2, 3))
    else:
        mean, var = np.asarray(buffers[name + ".2.running_mean"], F64), np.asarray(buffers[name + ".2.running_var"], F64)
    inv = 1.0 / np.sqrt(var + BN_EPS)
    xhat = (t - mean[None, :, None, None]) * inv[None, :, None, None]
    z = xhat * gam[None, :, None, None] + bet[None, :, None, None]
    return np.maximum(z, 0.0), dict(x=x, d=d, t=t, xhat=xhat, inv=inv, z=z, mean=mean, var=var, train=train)


def _block_backward(P, name, c, g_out, G):
    wd = _p(P, name + ".0.weight")
    wp = _p(P, name + ".1.weight")[:, :, 0, 0]
    gam = _p(P, name + ".2.weight")
    g_z = g_out * (c["z"] > 0)
    G[name + ".2.bias"] = g_z.sum((0, 2, 3))
    G[name + ".2.weight"] = (g_z * c["xhat"]).sum((0, 2, 3))
    g_xhat = g_z * gam[None, :, None, None]
    if c["train"]:
        n = g_z.shape[0] * g_z.shape[2] * g_z.shape[3]
        s1 = g_xhat.sum((0, 2, 3))[None, :, None, None]
        s2 = (g_xhat * c["xhat"]).sum((0, 2, 3))[None, :, None, None]
        g_t = (g_xhat - s1 / n - c["xhat"] * s2 / n) * c["inv"][None, :, None, None]
    else:
        g_t = g_xhat * c["inv"][None, :, None, None]
    G[name + ".1.bias"] = g_t.sum((0, 2, 3))
    G[name + ".1.weight"] = np.einsum("bohw,bchw->oc", g_t, c["d"])[:, :, None, None]
    g_d = np.einsum("oc,bohw->bchw", wp, g_t)
    g_x, gw, gb = _dw3_backward(c["x"], wd, g_d)
    G[name + ".0.weight"], G[name + ".0.bias"] = gw, gb
    return g_x


def _gate_forward(P, name, e):
    """HeightAttention (:421-461) and its application `ca(x) * x`: four channel slices, each global max -> 1x1 conv
    (no bias) -> ReLU -> 1x1 conv (no bias) -> sigmoid."""
    B, C = e.shape[:2]
    cs = C // 4
    flat = e.reshape(B, C, -1)
    arg = flat.argmax(-1)  # first maximum, as ATen's adaptive max pool
    mx = np.take_along_axis(flat, arg[..., None], -1)[..., 0]
    gate = np.zeros((B, C), F64)
    cache = dict(e=e, arg=arg, mx=mx, hid=[], pre=[])
    for s in range(4):
        w1 = _p(P, "%s.conv%d.0.weight" % (name, s + 1))[:, :, 0, 0]
        w2 = _p(P, "%s.conv%d.2.weight" % (name, s + 1))[:, :, 0, 0]
        m = mx[:, s * cs:(s + 1) * cs]
        pre = m @ w1.T
        hid = np.maximum(pre, 0.0)
        gate[:, s * cs:(s + 1) * cs] = 1.0 / (1.0 + np.exp(-(hid @ w2.T)))
        cache["hid"].append(hid)
        cache["pre"].append(pre)
    cache["gate"] = gate
    return e * gate[:, :, None, None], cache


def _gate_backward(P, name, c, g_out, G):
    e, gate = c["e"], c["gate"]
    B, C = e.shape[:2]
    cs = C // 4
    g_e = g_out * gate[:, :, None, None]
    g_gate = (g_out * e).sum((2, 3))
    g_o = g_gate * gate * (1.0 - gate)
    g_mx = np.zeros((B, C), F64)
    for s in range(4):
        w1 = _p(P, "%s.conv%d.0.weight" % (name, s + 1))[:, :, 0, 0]
        w2 = _p(P, "%s.conv%d.2.weight" % (name, s + 1))[:, :, 0, 0]
        go = g_o[:, s * cs:(s + 1) * cs]
        G["%s.conv%d.2.weight" % (name, s + 1)] = (go.T @ c["hid"][s])[:, :, None, None]
        g_hid = go @ w2
        g_pre = g_hid * (c["pre"][s] > 0)
        G["%s.conv%d.0.weight" % (name, s + 1)] = (g_pre.T @ c["mx"][:, s * cs:(s + 1) * cs])[:, :, None, None]
        g_mx[:, s * cs:(s + 1) * cs] = g_pre @ w1
    flat = g_e.reshape(B, C, -1)
    np.add.at(flat, (np.arange(B)[:, None], np.arange(C)[None, :], c["arg"]), g_mx)
    return flat.reshape(e.shape)


def _pool(x):
    B, C, H, W = x.shape
    t = x.reshape(B, C, H // 2, 2, W // 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(B, C, H // 2, W // 2, 4)
    arg = t.argmax(-1)
    return np.take_along_axis(t, arg[..., None], -1)[..., 0], arg


def _pool_backward(g, arg):
    B, C, h, w = g.shape
    t = np.zeros((B, C, h, w, 4), F64)
    np.put_along_axis(t, arg[..., None], g[..., None], -1)
    return t.reshape(B, C, h, w, 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(B, C, 2 * h, 2 * w)


def _upconv(x, wgt, b):
    """ConvTranspose2d(k=2, s=2): out[b,o,2y+i,2x+j] = sum_c x[b,c,y,x] W[c,o,i,j] + b[o]."""
    B, C, h, w = x.shape
    t = np.einsum("bchw,coij->bohiwj", x, wgt)
    return t.reshape(B, wgt.shape[1], 2 * h, 2 * w) + b[None, :, None, None]


def _upconv_backward(x, wgt, g):
    B, C, h, w = x.shape
    g6 = g.reshape(B, wgt.shape[1], h, 2, w, 2)
    return np.einsum("bohiwj,coij->bchw", g6, wgt), np.einsum("bohiwj,bchw->coij", g6, x), g.sum((0, 2, 3))


def converter_forward(P, x, position, train=True, buffers=None):
    """OpacityVoxelToBEVConverter.forward (:494-518): x [B,13,S,S], position [1 or B,4,S,S] -> [B,1,S,S]."""
    x, position = np.asarray(x, F64), np.asarray(position, F64)
    C = {}
    e1, C["b1"] = _block_forward(P, "encoder1", x, train, buffers)
    enc1, C["g1"] = _gate_forward(P, "ca1", e1 + position)
    p1, C["p1"] = _pool(enc1)
    e2, C["b2"] = _block_forward(P, "encoder2", p1, train, buffers)
    enc2, C["g2"] = _gate_forward(P, "ca2", e2)
    p2, C["p2"] = _pool(enc2)
    e3, C["b3"] = _block_forward(P, "bottleneck", p2, train, buffers)
    bott, C["g3"] = _gate_forward(P, "ca_bottleneck", e3)
    u2 = _upconv(bott, _p(P, "upconv2.weight"), _p(P, "upconv2.bias"))
    e4, C["b4"] = _block_forward(P, "decoder2", np.concatenate([u2, enc2], 1), train, buffers)
    dec2, C["g4"] = _gate_forward(P, "ca_dec2", e4)
    u1 = _upconv(dec2, _p(P, "upconv1.weight"), _p(P, "upconv1.bias"))
    e5, C["b5"] = _block_forward(P, "decoder1", np.concatenate([u1, enc1], 1), train, buffers)
    dec1, C["g5"] = _gate_forward(P, "ca_dec1", e5)
    wo, bo = _p(P, "output_conv.weight")[:, :, 0, 0], _p(P, "output_conv.bias")
    out = np.einsum("oc,bchw->bohw", wo, dec1) + bo[None, :, None, None]
    C.update(bott=bott, dec2=dec2, dec1=dec1, pos_shape=position.shape)
    stats = {n: (C[k]["mean"], C[k]["var"]) for n, k in zip(BLOCKS, ("b1", "b2", "b3", "b4", "b5"))}
    return out, C, stats


def converter_backward(P, C, g_out):
    g_out = np.asarray(g_out, F64)
    G = {}
    wo = _p(P, "output_conv.weight")[:, :, 0, 0]
    G["output_conv.bias"] = g_out.sum((0, 2, 3))
    G["output_conv.weight"] = np.einsum("bohw,bchw->oc", g_out, C["dec1"])[:, :, None, None]
    g = np.einsum("oc,bohw->bchw", wo, g_out)
    g = _gate_backward(P, "ca_dec1", C["g5"], g, G)
    g_cat = _block_backward(P, "decoder1", C["b5"], g, G)
    g_u1, g_enc1 = g_cat[:, :4], g_cat[:, 4:]
    g_dec2, G["upconv1.weight"], G["upconv1.bias"] = _upconv_backward(C["dec2"], _p(P, "upconv1.weight"), g_u1)
    g = _gate_backward(P, "ca_dec2", C["g4"], g_dec2, G)
    g_cat = _block_backward(P, "decoder2", C["b4"], g, G)
    g_u2, g_enc2 = g_cat[:, :8], g_cat[:, 8:]
    g_bott, G["upconv2.weight"], G["upconv2.bias"] = _upconv_backward(C["bott"], _p(P, "upconv2.weight"), g_u2)
    g = _gate_backward(P, "ca_bottleneck", C["g3"], g_bott, G)
    g = _block_backward(P, "bottleneck", C["b3"], g, G)
    g_enc2 = g_enc2 + _pool_backward(g, C["p2"])
    g = _gate_backward(P, "ca2", C["g2"], g_enc2, G)
    g = _block_backward(P, "encoder2", C["b2"], g, G)
    g_enc1 = g_enc1 + _pool_backward(g, C["p1"])
    g = _gate_backward(P, "ca1", C["g1"], g_enc1, G)  # gradient of (e1 + position)
    g_pos = g.sum(0, keepdims=True) if C["pos_shape"][0] == 1 else g
    g_x = _block_backward(P, "encoder1", C["b1"], g, G)
    return g_x, g_pos, G
