"""ctypes binding of libocrf_raster.so (the C ABI declared in include/ocrf_raster.h).

There is no CPU fallback: if the library is missing or a CUDA device is absent the product path
raises.  Importing this module does not touch the GPU; only `lib()` loads the shared object.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libocrf_raster.so")

OCRF_CAM_STRIDE = 40
OCRF_RECORD_BYTES = 48
OCRF_GGRAD_STRIDE = 6
ABI_VERSION = 5
OCRF_EINVAL = -1
OCRF_ECAPACITY = -2
OCRF_BIN_PAIR_SORT = 1
OCRF_BIN_DEPTH_FIRST = 2

EXPORTS = [
    "ocrf_abi_version", "ocrf_error_string", "ocrf_debug_sync", "ocrf_geom_layout", "ocrf_bin_layout", "ocrf_image_layout",
    "ocrf_sort_end_bit", "ocrf_preprocess_forward", "ocrf_preprocess_forward_filtered", "ocrf_bin_forward", "ocrf_render_forward",
    "ocrf_render_backward", "ocrf_preprocess_backward", "ocrf_mark_visible", "ocrf_sort_workspace_bytes",
    "ocrf_sort_pairs", "ocrf_opacity_mask_forward", "ocrf_opacity_mask_backward",
    "ocrf_gaussian_heads_forward", "ocrf_gaussian_heads_backward", "ocrf_gaussian_heads_backward_workspace_bytes", "ocrf_clear_gradients",
    "ocrf_hoa_lift_workspace_floats", "ocrf_hoa_lift_forward", "ocrf_hoa_lift_backward",
    "ocrf_hoa_converter_workspace_floats", "ocrf_hoa_converter_forward", "ocrf_hoa_converter_backward",
    "ocrf_color_voxels", "ocrf_retain_valid_pixels", "ocrf_bev_pool_forward", "ocrf_bev_pool_backward_workspace_bytes", "ocrf_bev_pool_backward",
]


class OcrfShape(C.Structure):
    _fields_ = [("S", C.c_int32), ("P", C.c_int32), ("V", C.c_int32), ("views_per_sample", C.c_int32),
                ("W", C.c_int32), ("H", C.c_int32), ("C", C.c_int32), ("sh_degree", C.c_int32), ("sh_M", C.c_int32)]


class OcrfGeomLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("total", "header", "depths", "xy", "conic_opacity", "tiles_touched",
                                          "offsets", "rgb", "clamped", "scan_status", "vis_keys", "vis_vals",
                                          "vis_keys_tmp", "vis_vals_tmp", "view_start", "vis_sort_ws")]


class OcrfBinLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("total", "keys", "point_list", "keys_tmp", "vals_tmp", "keys_unsorted",
                                          "vals_unsorted", "records", "histogram", "sort_status", "split_counts",
                                          "split_tiles", "split_words", "split_total")]


class OcrfImageLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("total", "ranges", "ranges_render", "final_T", "n_contrib", "max_contrib")]


class OcrfError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libocrf_raster.so once.  Raises if it has not been built (python -m ocrfdet_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    import torch  # noqa: F401  (brings libcudart into the process before our library needs it)
    if not os.path.exists(LIB_PATH):
        raise OcrfError("libocrf_raster.so is not built: run `python -m ocrfdet_b200.build` "
                        "(there is no CPU or PyTorch fallback for the render path)")
    L = C.CDLL(LIB_PATH)
    L.ocrf_abi_version.restype = C.c_int
    L.ocrf_error_string.restype = C.c_char_p
    L.ocrf_error_string.argtypes = [C.c_int]
    L.ocrf_sort_workspace_bytes.restype = C.c_size_t
    L.ocrf_sort_workspace_bytes.argtypes = [C.c_uint64]
    vp, i32, u64, f32 = C.c_void_p, C.c_int32, C.c_uint64, C.c_float
    shp = C.POINTER(OcrfShape)
    L.ocrf_geom_layout.argtypes = [shp, C.c_int, C.POINTER(OcrfGeomLayout)]
    L.ocrf_bin_layout.argtypes = [shp, u64, C.POINTER(OcrfBinLayout)]
    L.ocrf_image_layout.argtypes = [shp, C.POINTER(OcrfImageLayout)]
    L.ocrf_sort_end_bit.argtypes = [shp]
    L.ocrf_preprocess_forward.argtypes = [vp, shp, vp, vp, vp, vp, vp, vp, vp, f32, C.c_int, vp, vp]
    L.ocrf_preprocess_forward_filtered.argtypes = [vp, shp, vp, vp, vp, vp, vp, vp, vp, f32, C.c_int, f32, vp, vp]
    L.ocrf_bin_forward.argtypes = [vp, shp, u64, vp, vp, C.c_int, C.c_uint32, vp, vp, vp, vp]
    L.ocrf_debug_sync.argtypes = [vp]
    L.ocrf_render_forward.argtypes = [vp, shp, u64, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.ocrf_render_backward.argtypes = [vp, shp, u64, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ocrf_preprocess_backward.argtypes = [vp, shp, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                           vp, vp]
    L.ocrf_mark_visible.argtypes = [vp, i32, vp, vp, vp, vp]
    L.ocrf_sort_pairs.argtypes = [vp, u64, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.ocrf_opacity_mask_forward.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ocrf_opacity_mask_backward.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ocrf_hoa_lift_workspace_floats.restype = C.c_size_t
    L.ocrf_hoa_lift_workspace_floats.argtypes = [i32, i32, i32, i32]
    L.ocrf_hoa_lift_forward.argtypes = [vp, i32, i32, i32, i32] + [vp] * 6
    L.ocrf_hoa_lift_backward.argtypes = [vp, i32, i32, i32, i32] + [vp] * 9
    L.ocrf_hoa_converter_workspace_floats.restype = C.c_size_t
    L.ocrf_hoa_converter_workspace_floats.argtypes = [i32, i32]
    L.ocrf_hoa_converter_forward.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.ocrf_hoa_converter_backward.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ocrf_color_voxels.argtypes = [vp, i32, i32, C.c_int64, i32, i32, i32, vp, vp, vp, f32, vp, vp]
    L.ocrf_retain_valid_pixels.argtypes = [vp, i32, C.c_int64, i32, i32, i32, vp, vp, vp, f32, vp]
    L.ocrf_bev_pool_forward.argtypes = [vp, i32, i32] + [vp] * 8
    L.ocrf_bev_pool_backward_workspace_bytes.restype = C.c_size_t
    L.ocrf_bev_pool_backward_workspace_bytes.argtypes = [u64]
    L.ocrf_bev_pool_backward.argtypes = [vp, i32, u64, i32] + [vp] * 9
    L.ocrf_clear_gradients.argtypes = [vp, shp, C.c_int, vp, vp, vp]
    L.ocrf_gaussian_heads_forward.argtypes = [vp, C.c_int64, i32] + [vp] * 11
    L.ocrf_gaussian_heads_backward.argtypes = [vp, C.c_int64, i32] + [vp] * 16
    L.ocrf_gaussian_heads_backward_workspace_bytes.restype = C.c_size_t
    L.ocrf_gaussian_heads_backward_workspace_bytes.argtypes = [C.c_int64]
    if L.ocrf_abi_version() != ABI_VERSION:
        raise OcrfError("libocrf_raster.so ABI %d != expected %d: rebuild" % (L.ocrf_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise OcrfError("%s failed: %s (code %d)" % (what, lib().ocrf_error_string(rc).decode(), rc))


def ptr(t):
    """Device pointer of a tensor, or NULL for None / an empty tensor (as the reference binding)."""
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
