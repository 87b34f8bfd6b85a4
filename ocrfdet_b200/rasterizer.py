"""Host-side mirror of the reference plugin `diff_gaussian_rasterization` on top of the C ABI.

Same names, argument meaning and error behaviour as
/root/reference/mmdet3d/models/necks/MVSGaussian/lib/submodules/diff-gaussian-rasterization/
diff_gaussian_rasterization/__init__.py  (cited below as PKG:line), with the live w-depth fork's
3-tuple return `(color, radii, depth)` that OcRFDet unpacks
(/root/reference/mmdet3d/models/necks/MVSGaussian/lib/gaussian_renderer/__init__.py:62).

Extensions (all opt-in, the reference call keeps working unmodified):
  * `debug` defaults to False, so the fork's 11-field settings tuple and the vendored 12-field one
    both construct.
  * `GaussianRasterizer(settings, return_opacity=True)` appends the accumulated opacity map
    (1 - final_T) with a backward.
  * any channel count C for `colors_precomp` (the reference is compiled for C = 3).
  * `render_batch`: every (sample, view) pair of a rank in ONE launch sequence.

PyTorch is plumbing here (device memory, streams, autograd glue); all compute is in
libocrf_raster.so and there is no CPU / eager fallback.
"""
import ctypes as C
import os
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import OcrfBinLayout, OcrfGeomLayout, OcrfImageLayout, OcrfShape, check, current_stream, ptr


class GaussianRasterizationSettings(NamedTuple):
    """PKG:157-169 (the live fork omits `debug`: GR:39-57)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool = False


def pack_cameras(viewmatrix, projmatrix, campos, tanfovx, tanfovy):
    """Camera records [V, 40] for the C ABI: view[16] proj[16] campos[3] tanfovx tanfovy pad[3].

    viewmatrix / projmatrix are the reference's transposed 4x4 matrices ([4,4] or [V,4,4]);
    tanfovx / tanfovy are floats or [V] tensors.  Built on the device of `viewmatrix`.
    """
    vm = viewmatrix.reshape(-1, 16).float()
    V = vm.shape[0]
    pm = projmatrix.reshape(-1, 16).float()
    cp = campos.reshape(-1, 3).float()
    dev = vm.device

    def col(t):
        if torch.is_tensor(t):
            return t.reshape(-1, 1).float().to(dev).expand(V, 1)
        return torch.full((V, 1), float(t), dtype=torch.float32, device=dev)

    pad = torch.zeros((V, 3), dtype=torch.float32, device=dev)
    return torch.cat([vm, pm, cp, col(tanfovx), col(tanfovy), pad], dim=1).contiguous()


def pack_camera_dicts(cams, device="cuda"):
    """`pack_cameras` for a list of `ocrfdet_b200.cameras.make_camera` dicts (numpy) -> [V,40] on `device`."""
    import numpy as np
    vm = torch.from_numpy(np.stack([c["viewmatrix"] for c in cams])).to(device)
    pm = torch.from_numpy(np.stack([c["projmatrix"] for c in cams])).to(device)
    cp = torch.from_numpy(np.stack([c["campos"] for c in cams])).to(device)
    tx = torch.tensor([c["tanfovx"] for c in cams], dtype=torch.float32).to(device)
    ty = torch.tensor([c["tanfovy"] for c in cams], dtype=torch.float32).to(device)
    return pack_cameras(vm, pm, cp, tx, ty)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise _lib.OcrfError("%s must be a CUDA tensor: the render path has no CPU implementation" % name)


class _Status:
    """Per-device sticky status words of the render path (`sticky_status` of ocrf_bin_forward).

    The library ORs the error flags of every call into `dev[0]` and keeps the largest pair count in `dev[1]`; it
    never clears them.  Capacity-mode calls (no host read-back) queue an 8-byte copy into the pinned mirror `host`
    after their forward, and every later entry into the path (`render_batch`, the autograd backward,
    `check_overflow`) looks at the mirror: an overflow is reported at the next call without any synchronisation --
    a training loop cannot keep rendering background images unnoticed.  Device words, not Python state, carry the
    flag, so concurrent streams and CUDA-graph replays all land in the same place.
    """
    _per_device = {}

    def __init__(self, device):
        self.dev = torch.zeros(2, dtype=torch.int32, device=device)
        self.host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.host_np = self.host.numpy()  # same memory: polling costs a numpy scalar read, not a tensor op
        self.muted = False  # GraphedRenderStep: no exception out of the middle of a capture (replays never poll)

    @classmethod
    def get(cls, device):
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        st = cls._per_device.get(device.index)
        if st is None:
            st = cls._per_device[device.index] = cls(device)
        return st

    def mirror(self):
        """Queue the 8-byte copy of the status words into the pinned mirror (asynchronous, capturable)."""
        self.host.copy_(self.dev, non_blocking=True)

    def _raise(self, flags, pairs):
        self.dev.zero_()
        self.host_np[:] = 0
        if flags & 1:
            raise _lib.OcrfError("binning workspace overflow: %d (tile, Gaussian) pairs exceed pair_capacity; that "
                                 "call rendered background only" % (pairs & 0xFFFFFFFF))
        raise _lib.OcrfError("Point is filtered although prefiltered is set. This shouldn't happen!")

    def poll(self):
        """Non-blocking: raise if a mirrored status word shows an error of an earlier call."""
        if self.muted:
            return
        flags = int(self.host_np[0])
        if flags & 3:
            self._raise(flags, int(self.host_np[1]))

    def check(self):
        """Blocking: read the device words (8 bytes) and raise if any call since the last check failed."""
        h = self.dev.cpu()
        flags = int(h[0])
        if flags & 3:
            self._raise(flags, int(h[1]))


class _Workspaces:
    """Byte layouts of the three caller-owned workspaces (one query each, no GPU work)."""

    def __init__(self, shape, use_sh):
        L = _lib.lib()
        self.geom = OcrfGeomLayout()
        self.image = OcrfImageLayout()
        check(L.ocrf_geom_layout(C.byref(shape), int(use_sh), C.byref(self.geom)), "ocrf_geom_layout")
        check(L.ocrf_image_layout(C.byref(shape), C.byref(self.image)), "ocrf_image_layout")

    @staticmethod
    def bin_layout(shape, n):
        b = OcrfBinLayout()
        check(_lib.lib().ocrf_bin_layout(C.byref(shape), C.c_uint64(n), C.byref(b)), "ocrf_bin_layout")
        return b


class _RasterizeBatch(torch.autograd.Function):
    """The autograd node (PKG:44-155), batched over V views of S samples."""

    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors, opacities, scales, rotations, cov3D_precomp, cams, bg, cfg):
        _Status.get(means3D.device).poll()  # an overflow of an earlier capacity-mode call surfaces here, sync-free
        args = (ctx, means3D, means2D, shs, colors, opacities, scales, rotations, cov3D_precomp, cams, bg, cfg)
        if cfg.get("debug"):
            return _debug_guard("forward", "snapshot_fw.dump", args[1:], lambda: _RasterizeBatch._forward_impl(*args))
        return _RasterizeBatch._forward_impl(*args)

    @staticmethod
    def _forward_impl(ctx, means3D, means2D, shs, colors, opacities, scales, rotations, cov3D_precomp, cams, bg, cfg):
        L = _lib.lib()
        S, P = means3D.shape[0], means3D.shape[1]
        V = cams.shape[0]
        use_sh = shs is not None
        Cc = 3 if use_sh else colors.shape[-1]
        W, H = cfg["W"], cfg["H"]
        shape = OcrfShape(S, P, V, V // S, W, H, Cc, cfg["sh_degree"], shs.shape[2] if use_sh else 0)
        dev = means3D.device
        stream = current_stream()
        status = _Status.get(dev)
        debug = bool(cfg.get("debug"))
        if use_sh and cfg.get("colors_ready") is not None:
            torch.cuda.current_stream().wait_event(cfg["colors_ready"])
        ws = _Workspaces(shape, use_sh)
        radii = torch.empty((V, P), dtype=torch.int32, device=dev)
        geom = torch.empty(ws.geom.total, dtype=torch.uint8, device=dev)
        image = torch.empty(ws.image.total, dtype=torch.uint8, device=dev)
        color = torch.empty((V, Cc, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((V, 1, H, W), dtype=torch.float32, device=dev)
        opac = torch.empty((V, 1, H, W), dtype=torch.float32, device=dev)

        check(L.ocrf_preprocess_forward_filtered(stream, C.byref(shape), ptr(means3D), ptr(scales), ptr(rotations),
                                                 ptr(cov3D_precomp), ptr(opacities), ptr(shs), ptr(cams),
                                                 C.c_float(cfg["scale_modifier"]), int(cfg["prefiltered"]),
                                                 C.c_float(cfg.get("min_opacity", 0.0)), ptr(radii), ptr(geom)),
              "ocrf_preprocess_forward")
        _stage("preprocess")
        if debug:
            _debug_sync("preprocess")
        if cfg.get("colors_ready") is not None and not use_sh:   # SH coefficients are read by the preprocess itself
            torch.cuda.current_stream().wait_event(cfg["colors_ready"])
        capacity = cfg.get("pair_capacity")
        if capacity is None:
            # exact sizing: one 8-byte read-back per BATCH (the reference syncs once per view,
            # rasterizer_impl.cu:281)
            hdr = geom[ws.geom.header:ws.geom.header + 8].view(torch.int32).cpu()
            num_pairs, err = int(hdr[0]) & 0xFFFFFFFF, int(hdr[1])
            if err & 2:  # the reference traps on the device here (CR/auxiliary.h:156-160)
                status.dev.zero_()
                raise _lib.OcrfError("Point is filtered although prefiltered is set. This shouldn't happen!")
            capacity = num_pairs
        else:
            num_pairs = None
            capacity = max(int(capacity), 1)
        binl = _Workspaces.bin_layout(shape, capacity)
        # the default multi-split mode touches only the first `split_total` bytes of the layout (56 B per pair)
        split_mode = cfg.get("binning") not in ("pairsort", "depthfirst") and ((W + 15) // 16) * ((H + 15) // 16) <= 4096
        binning = torch.empty(binl.split_total if split_mode else binl.total, dtype=torch.uint8, device=dev)
        check(L.ocrf_bin_forward(stream, C.byref(shape), C.c_uint64(capacity), ptr(radii), ptr(colors), int(use_sh),
                                 C.c_uint32({"pairsort": _lib.OCRF_BIN_PAIR_SORT, "depthfirst": _lib.OCRF_BIN_DEPTH_FIRST}
                                            .get(cfg.get("binning"), 0)),
                                 ptr(geom), ptr(binning), ptr(image), ptr(status.dev)), "ocrf_bin_forward")
        _stage("binning")
        if debug:
            _debug_sync("binning")
            status.check()
        check(L.ocrf_render_forward(stream, C.byref(shape), C.c_uint64(capacity), ptr(colors), int(use_sh), ptr(bg),
                                    ptr(geom), ptr(binning), ptr(image), ptr(color), ptr(depth), ptr(opac)),
              "ocrf_render_forward")
        _stage("render_forward")
        if debug:
            _debug_sync("render_forward")
        if num_pairs is None:
            status.mirror()  # capacity mode: the next entry into the path sees this call's flags without a sync

        global _LAST_STATE
        if KEEP_STATE:
            _LAST_STATE = dict(shape=shape, layouts=(ws.geom, binl, ws.image), geom=geom, binning=binning, image=image,
                               radii=radii, capacity=capacity, colors=colors, use_sh=use_sh,
                               binning_mode=cfg.get("binning") if cfg.get("binning") in ("pairsort", "depthfirst")
                               else "split")
        ctx.shape, ctx.cfg, ctx.capacity, ctx.use_sh = shape, cfg, capacity, use_sh
        ctx.num_rendered = num_pairs
        ctx.layouts = (ws.geom, binl, ws.image)
        ctx.save_for_backward(means3D, shs, colors, scales, rotations, cov3D_precomp, cams, bg, radii, geom, binning,
                              image)
        ctx.mark_non_differentiable(radii, depth)  # no depth backward in the fork (its README:13)
        ctx.set_materialize_grads(False)  # no zero-filled gradient tensors for radii / depth / unused outputs
        return color, radii, depth, opac

    @staticmethod
    def backward(ctx, g_color, _g_radii, _g_depth, g_opac):
        _Status.get(ctx.saved_tensors[0].device).poll()
        if ctx.cfg.get("debug"):
            return _debug_guard("backward", "snapshot_bw.dump", tuple(ctx.saved_tensors) + (g_color, g_opac, ctx.cfg),
                                lambda: _RasterizeBatch._backward_impl(ctx, g_color, g_opac))
        return _RasterizeBatch._backward_impl(ctx, g_color, g_opac)

    @staticmethod
    def _backward_impl(ctx, g_color, g_opac):
        L = _lib.lib()
        means3D, shs, colors, scales, rotations, cov3D_precomp, cams, bg, radii, geom, binning, image = ctx.saved_tensors
        shape, cfg, use_sh = ctx.shape, ctx.cfg, ctx.use_sh
        S, P, V, Cc = shape.S, shape.P, shape.V, shape.C
        dev = means3D.device
        stream = current_stream()
        debug = bool(cfg.get("debug"))
        _stage("backward_begin")
        if g_color is None:
            g_color = torch.zeros((V, Cc, shape.H, shape.W), dtype=torch.float32, device=dev)
        g_color = g_color.contiguous()
        g_opac = g_opac.contiguous() if g_opac is not None else None
        ggrad = torch.empty((V, P, _lib.OCRF_GGRAD_STRIDE), dtype=torch.float64, device=dev)
        g_feat = torch.empty((V, P, 3) if use_sh else (S, P, Cc), dtype=torch.float32, device=dev)
        check(L.ocrf_clear_gradients(stream, C.byref(shape), int(use_sh), ptr(radii), ptr(ggrad), ptr(g_feat)),
              "ocrf_clear_gradients")
        check(L.ocrf_render_backward(stream, C.byref(shape), C.c_uint64(ctx.capacity), ptr(colors), int(use_sh),
                                     ptr(bg), ptr(geom), ptr(binning), ptr(image), ptr(g_color), ptr(g_opac),
                                     ptr(ggrad), ptr(g_feat)), "ocrf_render_backward")
        _stage("render_backward")
        if debug:
            _debug_sync("render_backward")
        g_means3D = torch.empty_like(means3D)
        # dL/dmean2D is only materialised when the caller holds a means2D leaf (the reference's screenspace_points)
        g_means2D = torch.empty((V, P, 3), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        g_opacities = torch.empty((S, P, 1), dtype=torch.float32, device=dev)
        has_cov = cov3D_precomp is not None
        g_scales = None if has_cov else torch.empty_like(scales)
        g_rots = None if has_cov else torch.empty_like(rotations)
        g_cov = torch.empty_like(cov3D_precomp) if has_cov else None
        g_shs = torch.empty_like(shs) if use_sh else None
        check(L.ocrf_preprocess_backward(stream, C.byref(shape), ptr(means3D), ptr(scales), ptr(rotations),
                                         ptr(cov3D_precomp), ptr(shs), ptr(cams), C.c_float(cfg["scale_modifier"]),
                                         ptr(radii), ptr(geom), ptr(ggrad), ptr(g_feat) if use_sh else None,
                                         ptr(g_means3D), ptr(g_means2D), ptr(g_opacities), ptr(g_scales), ptr(g_rots),
                                         ptr(g_cov), ptr(g_shs)), "ocrf_preprocess_backward")
        _stage("preprocess_backward")
        if debug:
            _debug_sync("preprocess_backward")
        return (g_means3D, g_means2D, g_shs, None if use_sh else g_feat, g_opacities, g_scales, g_rots, g_cov, None,
                None, None)


KEEP_STATE = False   # tests / bench statistics: keep the workspaces of the most recent forward alive
_LAST_STATE = None
STAGE_HOOK = None    # bench: callable(name) invoked on the launching stream after each stage's launches


def _stage(name):
    if STAGE_HOOK is not None:
        STAGE_HOOK(name)


def _debug_sync(stage):
    """`debug=True` (PKG:83-90 -> CHECK_CUDA, CR/auxiliary.h:166-173): synchronise after every stage and raise
    on the first CUDA error, naming the stage."""
    rc = _lib.lib().ocrf_debug_sync(current_stream())
    if rc != 0:
        raise _lib.OcrfError("%s failed (debug mode): %s (code %d)"
                             % (stage, _lib.lib().ocrf_error_string(rc).decode(), rc))


def _debug_guard(which, dump_name, args, fn):
    """PKG:83-90 / PKG:132-139: copy the arguments to the CPU before the call; if it raises, save them as
    `snapshot_fw.dump` / `snapshot_bw.dump` in the working directory and re-raise."""
    plain = (int, float, bool, str, type(None))
    cpu_args = tuple(a.detach().cpu().clone() if torch.is_tensor(a) else
                     ({k: v for k, v in a.items() if isinstance(v, plain)} if isinstance(a, dict) else a) for a in args)
    try:
        return fn()
    except Exception:
        torch.save(cpu_args, dump_name)
        print("\nAn error occured in %s. Please forward %s for debugging." % (which, dump_name))
        raise


def last_state(reference_lists=False):
    """Typed views into the workspaces of the most recent forward (requires KEEP_STATE = True).

    Mirrors what the reference keeps in its geomBuffer / binningBuffer / imgBuffer blobs (PKG:97).
    With `reference_lists=True` and the default (depth-first) binning, the reference's own algorithm --
    sort every (tile | depth) pair -- is additionally run on the same geometry state and its results are
    returned as `keys_ref`, `point_list_ref`, `records_ref`, `ranges_ref`, `ranges_render_ref`.
    Reading `num_pairs` synchronises.
    """
    st = _LAST_STATE
    if st is None:
        raise _lib.OcrfError("no state kept: set ocrfdet_b200.rasterizer.KEEP_STATE = True before rendering")
    shape, (g, b, im) = st["shape"], st["layouts"]
    geom, binning, image = st["geom"], st["binning"], st["image"]
    V, P, W, H = shape.V, shape.P, shape.W, shape.H
    n = V * P
    tiles = ((W + 15) // 16) * ((H + 15) // 16)

    def view(buf, off, count, dtype, *dims):
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return buf[off:off + nbytes].view(dtype).view(*dims)

    hdr = view(geom, g.header, 4, torch.int32, 4).cpu()
    N = int(hdr[0]) & 0xFFFFFFFF
    if N > st["capacity"]:
        N = 0
    out = dict(
        num_pairs=N, error=int(hdr[1]), num_visible=int(hdr[3]) & 0xFFFFFFFF, radii=st["radii"], binning=st["binning_mode"],
        depths=view(geom, g.depths, n, torch.float32, V, P), xy=view(geom, g.xy, 2 * n, torch.float32, V, P, 2),
        conic_opacity=view(geom, g.conic_opacity, 4 * n, torch.float32, V, P, 4),
        tiles_touched=view(geom, g.tiles_touched, n, torch.int32, V, P),
        offsets=view(geom, g.offsets, n, torch.int32, V * P),
        ranges=view(image, im.ranges, 2 * V * tiles, torch.int32, V, tiles, 2),
        ranges_render=view(image, im.ranges_render, 2 * V * tiles, torch.int32, V, tiles, 2),
        records=view(binning, b.records, 12 * N, torch.int32, N, 12),
        final_T=view(image, im.final_T, V * H * W, torch.float32, V, H, W),
        n_contrib=view(image, im.n_contrib, V * H * W, torch.int32, V, H, W),
        max_contrib=view(image, im.max_contrib, V * tiles, torch.int32, V, tiles))
    if st["binning_mode"] != "split":  # the multi-split path never materialises the pair lists
        out["keys"] = view(binning, b.keys, N, torch.int64, N)
        out["point_list"] = view(binning, b.point_list, N, torch.int32, N)
    if reference_lists and st["binning_mode"] != "pairsort":
        L = _lib.lib()
        bin2 = torch.empty(b.total, dtype=torch.uint8, device=binning.device)
        img2 = torch.empty_like(image)
        check(L.ocrf_bin_forward(current_stream(), C.byref(shape), C.c_uint64(st["capacity"]), ptr(st["radii"]),
                                 ptr(st["colors"]), int(st["use_sh"]), C.c_uint32(_lib.OCRF_BIN_PAIR_SORT), ptr(geom),
                                 ptr(bin2), ptr(img2), None), "ocrf_bin_forward(pair sort)")
        out["keys_ref"] = view(bin2, b.keys, N, torch.int64, N)
        out["point_list_ref"] = view(bin2, b.point_list, N, torch.int32, N)
        out["records_ref"] = view(bin2, b.records, 12 * N, torch.int32, N, 12)
        out["ranges_ref"] = view(img2, im.ranges, 2 * V * tiles, torch.int32, V, tiles, 2)
        out["ranges_render_ref"] = view(img2, im.ranges_render, 2 * V * tiles, torch.int32, V, tiles, 2)
    return out


def render_batch(means3D, opacities, cams, image_height, image_width, bg, colors_precomp=None, shs=None, scales=None,
                 rotations=None, cov3D_precomp=None, means2D=None, scale_modifier=1.0, sh_degree=0, prefiltered=False,
                 pair_capacity=None, binning: Optional[str] = None, colors_ready=None, debug: bool = False,
                 sample_chunk: Optional[int] = None, min_opacity: float = 0.0):
    """Render V = cams.shape[0] views of S = means3D.shape[0] samples in one launch sequence.

    means3D [S,P,3]; opacities [S,P,1]; colors_precomp [S,P,C] or shs [S,P,M,3]; scales [S,P,3] and
    rotations [S,P,4], or cov3D_precomp [S,P,6]; cams from `pack_cameras`, view v looks at sample
    v // (V // S).  Returns (color [V,C,H,W], radii [V,P], depth [V,1,H,W], opacity [V,1,H,W]).
    `means2D` ([V,P,3] zeros, requires_grad) receives dL/dmean2D as in the reference.
    `binning`: None / "split" = depth-sort the visible Gaussians, then ONE stable multi-split of the pair stream by
    tile writing the culled records directly (default; the pair lists are not materialised); "depthfirst" = same
    depth sort, pairs emitted in depth order, tile bits sorted (2 passes); "pairsort" = the reference's algorithm
    (sort every (tile | depth) pair).  All three give bit-identical records, range tables and images.
    `colors_ready`: optional `torch.cuda.Event`; colours / SH coefficients are first read by the binning stage, so a
    caller that uploads them on another stream can let that copy run under the preprocess and the depth sort: the
    launching stream waits for the event only after the preprocess has been queued.
    `pair_capacity`: if given, the binning workspace is sized for that many (tile, Gaussian) pairs
    and NO host synchronisation happens (CUDA-graph friendly); an overflow renders background and
    raises at the next entry into the path (`render_batch`, the backward) or `check_overflow()`, whichever comes first.
    `sample_chunk`: render at most that many samples per launch sequence (their views stay together), one after the
    other, and concatenate: bounds the transient workspaces -- the binning workspace is 56 bytes per (tile, Gaussian)
    pair -- for shapes like BASELINE config 5 (8 samples x 6 views at 512x1408 with a million Gaussians each: 134 M
    pairs per sample).  Every chunk is its own autograd node; `pair_capacity` may then be a sequence, one per chunk.
    `min_opacity`: foreground filter (OcRFDet renders every voxel of its 13 x 128 x 128 grid,
    view_transformer_ocrf.py:1130-1153).  Gaussians with opacity < min(min_opacity, 1/255) are culled in the preprocess
    like out-of-frustum ones: they could never pass the blend's alpha >= 1/255 test, so images and gradients are
    unchanged while radii (0 for them), keys and pair counts shrink.  0 (default) keeps the reference's radii.
    `debug`: the reference's debug mode (PKG:83-90,132-139; CR/auxiliary.h:166-173): synchronise and check for CUDA
    errors after every stage, and on any failure save the CPU copy of the arguments as `snapshot_fw.dump` /
    `snapshot_bw.dump` before re-raising.
    """
    _require_cuda(means3D, "means3D")
    if means3D.dim() != 3 or means3D.shape[-1] != 3:
        raise Exception("means3D must have dimensions (samples, num_points, 3)")
    S, P = means3D.shape[0], means3D.shape[1]
    if sample_chunk is not None and 0 < sample_chunk < S and cams.dim() == 2 and cams.shape[0] % S == 0:
        vps = cams.shape[0] // S
        cut = lambda t, a, b: None if t is None else t[a:b]  # noqa: E731
        outs = []
        for i, s0 in enumerate(range(0, S, sample_chunk)):
            s1 = min(S, s0 + sample_chunk)
            cap = pair_capacity[i] if isinstance(pair_capacity, (list, tuple)) else pair_capacity
            outs.append(render_batch(
                means3D[s0:s1], opacities[s0:s1], cams[s0 * vps:s1 * vps], image_height, image_width, bg,
                colors_precomp=cut(colors_precomp, s0, s1), shs=cut(shs, s0, s1), scales=cut(scales, s0, s1),
                rotations=cut(rotations, s0, s1), cov3D_precomp=cut(cov3D_precomp, s0, s1),
                means2D=cut(means2D, s0 * vps, s1 * vps), scale_modifier=scale_modifier, sh_degree=sh_degree,
                prefiltered=prefiltered, pair_capacity=cap, binning=binning, colors_ready=colors_ready, debug=debug,
                min_opacity=min_opacity))
        return tuple(torch.cat(parts, 0) for parts in zip(*outs))
    if isinstance(pair_capacity, (list, tuple)):
        pair_capacity = pair_capacity[0]
    if cams.dim() != 2 or cams.shape[1] != _lib.OCRF_CAM_STRIDE:
        raise Exception("cams must have dimensions (views, %d): build it with pack_cameras" % _lib.OCRF_CAM_STRIDE)
    V = cams.shape[0]
    if V == 0 or V % S != 0:
        raise Exception("the number of views must be a positive multiple of the number of samples")
    if (shs is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    Cc = 3 if shs is not None else colors_precomp.shape[-1]

    def expect(t, name, *tail):
        if t is not None and tuple(t.shape) != (S, P) + tail:
            raise Exception("%s must have dimensions %s, got %s" % (name, (S, P) + tail, tuple(t.shape)))

    if opacities.numel() != S * P:
        raise Exception("opacities must have dimensions (%d, %d, 1), got %s" % (S, P, tuple(opacities.shape)))
    expect(colors_precomp, "colors_precomp", Cc)
    expect(scales, "scales", 3)
    expect(rotations, "rotations", 4)
    expect(cov3D_precomp, "cov3D_precomp", 6)
    if shs is not None and (shs.dim() != 4 or tuple(shs.shape[:2]) != (S, P) or shs.shape[3] != 3):
        raise Exception("shs must have dimensions (%d, %d, M, 3), got %s" % (S, P, tuple(shs.shape)))
    if bg.numel() < Cc:  # the kernels read bg[k] for every channel k
        raise Exception("bg must hold one value per channel (%d), got %d" % (Cc, bg.numel()))
    if Cc > 96 and torch.is_grad_enabled() and any(
            t is not None and t.requires_grad for t in (means3D, opacities, colors_precomp, scales, rotations,
                                                        cov3D_precomp, means2D)):
        raise Exception("more than 96 feature channels are forward-only: the blend backward keeps a pixel's upstream "
                        "gradient in registers (C <= 96)")
    f = lambda t: None if t is None else t.float().contiguous()  # noqa: E731
    if means2D is None:  # gradient holder only, never read: no fill, and no dL/dmean2D output in backward
        means2D = torch.empty((V, P, 3), dtype=torch.float32, device=means3D.device)
    cfg = dict(W=int(image_width), H=int(image_height), scale_modifier=float(scale_modifier), sh_degree=int(sh_degree),
               prefiltered=bool(prefiltered), pair_capacity=pair_capacity, colors_ready=colors_ready, debug=bool(debug),
               min_opacity=float(min_opacity),
               binning=binning if binning is not None else os.environ.get("OCRF_BINNING", "split"))
    if P == 0:
        z = lambda c: torch.zeros((V, c, cfg["H"], cfg["W"]), dtype=torch.float32, device=means3D.device)  # noqa
        return z(Cc), torch.zeros((V, 0), dtype=torch.int32, device=means3D.device), z(1), z(1)
    return _RasterizeBatch.apply(f(means3D), means2D, f(shs), f(colors_precomp), f(opacities), f(scales), f(rotations),
                                 f(cov3D_precomp), f(cams), f(bg), cfg)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, return_opacity=False):
    """PKG:20-42: the single-view entry.  Empty tensors stand for absent inputs (PKG:197-207)."""
    none_if_empty = lambda t: None if (t is None or t.numel() == 0) else t  # noqa: E731
    channels = colors_precomp.shape[-1] if (colors_precomp is not None and colors_precomp.dim() == 2) else 3
    sh, colors_precomp = none_if_empty(sh), none_if_empty(colors_precomp)
    scales, rotations, cov3Ds_precomp = none_if_empty(scales), none_if_empty(rotations), none_if_empty(cov3Ds_precomp)
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise Exception("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:57-59
    st = raster_settings
    P = means3D.shape[0]
    H, W = int(st.image_height), int(st.image_width)
    dev = means3D.device
    if P == 0:  # rasterize_points.cu:81: zero image, nothing rendered
        Cc = channels
        out = (torch.zeros((Cc, H, W), dtype=torch.float32, device=dev), torch.zeros((0,), dtype=torch.int32, device=dev),
               torch.zeros((1, H, W), dtype=torch.float32, device=dev))
        return out + ((torch.zeros((1, H, W), dtype=torch.float32, device=dev),) if return_opacity else ())
    cams = pack_cameras(st.viewmatrix.to(dev), st.projmatrix.to(dev), st.campos.to(dev), st.tanfovx, st.tanfovy)
    u = lambda t: None if t is None else t.unsqueeze(0)  # noqa: E731
    m2d = means2D.unsqueeze(0) if means2D is not None else None
    color, radii, depth, opac = render_batch(
        u(means3D), u(opacities.reshape(P, 1)), cams, H, W, st.bg.to(dev), colors_precomp=u(colors_precomp),
        shs=u(sh), scales=u(scales), rotations=u(rotations), cov3D_precomp=u(cov3Ds_precomp), means2D=m2d,
        scale_modifier=st.scale_modifier, sh_degree=st.sh_degree, prefiltered=st.prefiltered,
        debug=bool(getattr(st, "debug", False)))
    out = (color[0], radii[0], depth[0])
    return out + ((opac[0],) if return_opacity else ())


class GaussianRasterizer(nn.Module):
    """PKG:171-220."""

    def __init__(self, raster_settings, return_opacity=False, return_depth=True):
        """`return_depth=False` gives the vendored package's 2-tuple `(color, radii)` (PKG:98, used by
        MVSGaussian's `gaussian_renderer_ft`); the default is the live w-depth fork's `(color, radii, depth)`."""
        super().__init__()
        self.raster_settings = raster_settings
        self.return_opacity = return_opacity
        self.return_depth = return_depth

    def markVisible(self, positions):
        # PKG:176-185 -> rasterizer_impl.cu:141-152
        with torch.no_grad():
            st = self.raster_settings
            _require_cuda(positions, "positions")
            positions = positions.float().contiguous()
            P = positions.shape[0]
            present = torch.zeros((P,), dtype=torch.bool, device=positions.device)
            vm = st.viewmatrix.to(positions.device).float().contiguous()
            pm = st.projmatrix.to(positions.device).float().contiguous()
            check(_lib.lib().ocrf_mark_visible(current_stream(), P, ptr(positions), ptr(vm), ptr(pm), ptr(present)),
                  "ocrf_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        out = rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                  self.raster_settings, return_opacity=self.return_opacity)
        return out if self.return_depth else out[:2] + out[3:]


def check_overflow(device=None):
    """Raise if any capacity-mode render on `device` since the last check overflowed its binning workspace (such a
    call renders background only).  Reads 8 bytes back: synchronises.  Calling it is optional -- the same error is
    raised without a synchronisation at the next `render_batch` / backward after the overflowing call."""
    if not torch.cuda.is_available():
        return
    _Status.get(device if device is not None else torch.device("cuda", torch.cuda.current_device())).check()
