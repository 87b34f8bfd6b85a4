"""Dense PyTorch compositing oracle of the render path (CPU, float32 or float64, autograd backward).

TEST / BASELINE INFRASTRUCTURE ONLY (imported by tests/ and by bench.py's cpu_baseline leg).

An independent second restatement of the same math as ocrf_oracle.c, written the "obvious" dense way
-- every pixel against every Gaussian, transmittance by cumulative product -- so that
  (a) the C oracle's closed-form backward can be checked against autograd, and
  (b) there is a CPU baseline for a path whose reference implementation is CUDA-only
      (BASELINE.json north_star: "a dense PyTorch compositing oracle of the same math").

Semantics reproduced: near cull z <= 0.2, det == 0 skip, tile-rectangle mask,
order by depth bits then index, power > 0 skip, alpha = min(0.99, o*exp(power)), alpha < 1/255 skip,
stop-before-blend at T(1-alpha) < 1e-4, out = C + T*bg, median depth (default 15), opacity = 1 - T.
Reference lines: cuda_rasterizer/forward.cu:74-152,155-256,261-374; auxiliary.h:41-77,139-164.
"""
import torch

TILE = 16


def _cam_tensors(cam, dtype):
    view = torch.as_tensor(cam["viewmatrix"], dtype=dtype).reshape(4, 4)   # transposed: row-vector convention
    proj = torch.as_tensor(cam["projmatrix"], dtype=dtype).reshape(4, 4)
    return view, proj


def preprocess(means3D, scales, rotations, opacities, cam, W, H, scale_modifier=1.0, cov3D_precomp=None):
    """Differentiable stage 1.  Returns dict with per-Gaussian xy, depth, conic (A,B,C), opacity, radii, rect, valid."""
    dtype = means3D.dtype
    view, proj = _cam_tensors(cam, dtype)
    P = means3D.shape[0]
    hom = torch.cat([means3D, torch.ones(P, 1, dtype=dtype)], 1)
    t = hom @ view          # [P,4] view space (row vector times transposed matrix)
    h = hom @ proj
    pw = 1.0 / (h[:, 3] + 1e-7)
    ndc = h[:, :2] * pw[:, None]
    tz = t[:, 2]
    valid = tz > 0.2
    tanx, tany = cam["tanfovx"], cam["tanfovy"]
    fx, fy = W / (2.0 * tanx), H / (2.0 * tany)
    if cov3D_precomp is None:
        r, x, y, z = rotations.unbind(1)
        R = torch.stack([
            1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
        Mm = R * (scales * scale_modifier)[:, None, :]      # R diag(s)
        Sigma = Mm @ Mm.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], 1).reshape(P, 3, 3)
    safe_tz = torch.where(valid, tz, torch.ones_like(tz))
    limx, limy = 1.3 * tanx, 1.3 * tany
    cx = torch.clamp(t[:, 0] / safe_tz, -limx, limx) * safe_tz
    cy = torch.clamp(t[:, 1] / safe_tz, -limy, limy) * safe_tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / safe_tz, zero, -(fx * cx) / (safe_tz * safe_tz),
                     zero, fy / safe_tz, -(fy * cy) / (safe_tz * safe_tz)], 1).reshape(P, 2, 3)
    Rv = view[:3, :3].T      # rotation part of world->view as a column-vector matrix
    M2 = J @ Rv              # [P,2,3]
    cov = M2 @ Sigma @ M2.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    valid = valid & (det != 0)
    sdet = torch.where(det != 0, det, torch.ones_like(det))
    conic = torch.stack([c / sdet, -b / sdet, a / sdet], 1)
    mid = 0.5 * (a + c)
    disc = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    lam = torch.maximum(mid + disc, mid - disc)
    radii = torch.ceil(3.0 * torch.sqrt(lam.detach().clamp(min=0))).to(torch.int64)
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    rf = radii.to(dtype)
    pxd, pyd = px.detach(), py.detach()
    x0 = torch.clamp(torch.trunc((pxd - rf) / TILE), 0, gx).to(torch.int64)
    y0 = torch.clamp(torch.trunc((pyd - rf) / TILE), 0, gy).to(torch.int64)
    x1 = torch.clamp(torch.trunc((pxd + rf + TILE - 1) / TILE), 0, gx).to(torch.int64)
    y1 = torch.clamp(torch.trunc((pyd + rf + TILE - 1) / TILE), 0, gy).to(torch.int64)
    valid = valid & ((x1 - x0) * (y1 - y0) > 0)
    radii = torch.where(valid, radii, torch.zeros_like(radii))
    return dict(xy=torch.stack([px, py], 1), depth=tz, conic=conic, opacity=opacities.reshape(-1), radii=radii,
                rect=(x0, y0, x1, y1), valid=valid)


def render(means3D, scales, rotations, opacities, colors, cam, W, H, bg, scale_modifier=1.0, cov3D_precomp=None,
           pixel_chunk=4096):
    """Dense forward.  Returns (color [C,H,W], depth [1,H,W], opacity [1,H,W], radii [P])."""
    dtype = means3D.dtype
    pre = preprocess(means3D, scales, rotations, opacities, cam, W, H, scale_modifier, cov3D_precomp)
    idx = torch.nonzero(pre["valid"]).reshape(-1)
    # order: depth bit pattern (positive floats order like their bits), ties by index -> stable sort of float32 depth
    d32 = pre["depth"].detach()[idx].to(torch.float32)
    order = idx[torch.argsort(d32, stable=True)]
    xy, conic, op, depth = pre["xy"][order], pre["conic"][order], pre["opacity"][order], pre["depth"][order]
    col = colors[order]
    x0, y0, x1, y1 = (r[order] for r in pre["rect"])
    Cc = colors.shape[1]
    bg = torch.as_tensor(bg, dtype=dtype)
    out_c = torch.zeros(H * W, Cc, dtype=dtype)
    out_d = torch.full((H * W,), 15.0, dtype=dtype)
    out_o = torch.zeros(H * W, dtype=dtype)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    cols, deps, ops = [], [], []
    for s in range(0, H * W, pixel_chunk):
        pxs, pys = xs[s:s + pixel_chunk], ys[s:s + pixel_chunk]
        txs, tys = (pxs // TILE)[:, None], (pys // TILE)[:, None]
        in_rect = (txs >= x0[None]) & (txs < x1[None]) & (tys >= y0[None]) & (tys < y1[None])
        dx = xy[None, :, 0] - pxs[:, None].to(dtype)
        dy = xy[None, :, 1] - pys[:, None].to(dtype)
        power = -0.5 * (conic[None, :, 0] * dx * dx + conic[None, :, 2] * dy * dy) - conic[None, :, 1] * dx * dy
        alpha = torch.clamp(op[None] * torch.exp(power), max=0.99)
        active = in_rect & (power <= 0) & (alpha >= 1.0 / 255.0)
        a_eff = torch.where(active, alpha, torch.zeros_like(alpha))
        T_incl = torch.cumprod(1.0 - a_eff, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], 1)
        # stop-before-blend: the first active Gaussian that would push T below 1e-4, and everything after it
        stop = active & (T_incl < 1e-4)
        alive = torch.cumsum(stop.to(torch.int32), dim=1) == 0
        w = torch.where(alive, a_eff * T_excl, torch.zeros_like(a_eff))
        cols.append(w @ col)
        # final T = product over blended Gaussians
        T_fin = torch.prod(torch.where(alive, 1.0 - a_eff, torch.ones_like(a_eff)), dim=1)
        ops.append(T_fin)
        med = alive & active & (T_excl > 0.5) & (T_incl < 0.5)
        has = med.any(dim=1)
        first = torch.argmax(med.to(torch.int32), dim=1)
        deps.append(torch.where(has, depth.detach()[first], torch.full_like(T_fin, 15.0)))
    T_fin = torch.cat(ops)
    out_c = torch.cat(cols) + T_fin[:, None] * bg[None]
    out_d = torch.cat(deps)
    return (out_c.T.reshape(Cc, H, W), out_d.reshape(1, H, W), (1.0 - T_fin).reshape(1, H, W), pre["radii"])
