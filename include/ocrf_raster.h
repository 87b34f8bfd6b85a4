/*
 * ocrf_raster.h -- C ABI of the B200-native OcRF Gaussian render path (libocrf_raster.so).
 *
 * This is the drop-in boundary for the reference's pybind module `diff_gaussian_rasterization._C`
 * (reference paths relative to
 *  /root/reference/mmdet3d/models/necks/MVSGaussian/lib/submodules/diff-gaussian-rasterization/):
 *
 *   _C.rasterize_gaussians           ext.cpp:16, rasterize_points.cu:36-115
 *        -> ocrf_preprocess_forward + ocrf_bin_forward + ocrf_render_forward
 *   _C.rasterize_gaussians_backward  ext.cpp:17, rasterize_points.cu:118-196
 *        -> ocrf_render_backward + ocrf_preprocess_backward
 *   _C.mark_visible                  ext.cpp:18, rasterize_points.cu:198-217
 *        -> ocrf_mark_visible
 *   ObatinOpacityMask + apply        /root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:230-242,1197-1199
 *        -> ocrf_opacity_mask_forward / ocrf_opacity_mask_backward
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, an explicit cudaStream_t passed as void*.
 *   - every entry point returns 0 on success, a positive cudaError_t on a CUDA failure, or a
 *     negative OCRF_E* code for an argument error.  No C++ exception crosses the boundary.
 *   - the caller owns ALL memory (as the reference's torch-owned buffers, rasterize_points.cu:27-33,
 *     68-78); the library never allocates.  Workspace sizes and layouts come from ocrf_*_layout.
 *   - a call renders a BATCH of V views (camera v looks at sample v / views_per_sample); V = 1,
 *     S = 1 is exactly the reference's single call.  4x4 matrices are the reference's transposed
 *     (column-major) layout, auxiliary.h:58-77.
 */
#ifndef OCRF_RASTER_H
#define OCRF_RASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCRF_ABI_VERSION 5

#define OCRF_EINVAL (-1)    /* bad argument (null pointer, non-positive size, unsupported channel count) */
#define OCRF_ECAPACITY (-2) /* workspace too small for the request */

/* floats per camera record: view[16] proj[16] campos[3] tanfovx tanfovy pad[3] */
#define OCRF_CAM_STRIDE 40
/* bytes per packed (tile, Gaussian) record consumed by the blend kernels */
#define OCRF_RECORD_BYTES 48
/* DOUBLES per (view, Gaussian) screen-space gradient record: dmean2D.xy, dconic.ABC, dopacity.
 * Per-tile partial sums are formed in fp32 (fixed order); the sum over the tiles a Gaussian touches -- the part
 * whose order is undefined because it goes through global reductions -- is accumulated in fp64, which makes the
 * gradients reproducible run to run to fp32 precision and keeps the ill-conditioned conic chain well fed. */
#define OCRF_GGRAD_STRIDE 6

typedef struct OcrfShape {
  int32_t S;                /* samples (independent Gaussian sets) */
  int32_t P;                /* Gaussians per sample */
  int32_t V;                /* views in this batch; V % views_per_sample == 0 */
  int32_t views_per_sample; /* view v renders sample v / views_per_sample */
  int32_t W, H;             /* image size in pixels (all views) */
  int32_t C;                /* feature channels per Gaussian (3 for RGB) */
  int32_t sh_degree;        /* active SH degree (only read when shs != NULL) */
  int32_t sh_M;             /* SH coefficients per Gaussian (shs is [S,P,M,3]) */
} OcrfShape;

/* Byte offsets of the arrays inside the three caller-allocated workspaces.  All offsets are
 * multiples of 128.  geom: per (view, Gaussian) state kept for backward (the reference's
 * GeometryState, rasterizer_impl.cu:155-170); binning: per (tile, Gaussian) pair (BinningState,
 * :182-194); image: per pixel / per tile (ImageState, :172-179). */
typedef struct OcrfGeomLayout {
  size_t total;
  size_t header;        /* uint32[32]: [0] num_pairs, [1] error flags, [2] block ticket, ... */
  size_t depths;        /* float  [V*P] */
  size_t xy;            /* float2 [V*P] */
  size_t conic_opacity; /* float4 [V*P] */
  size_t tiles_touched; /* uint32 [V*P] */
  size_t offsets;       /* uint32 [V*P] inclusive scan of tiles_touched over the whole batch */
  size_t rgb;           /* float  [V*P*3]  (SH path only) */
  size_t clamped;       /* uint8  [V*P*3]  (SH path only) */
  size_t scan_status;   /* uint64 [blocks] decoupled look-back state of the fused scan */
  size_t vis_keys;      /* uint64 [V*P] (view << 32 | depth bits) of the visible Gaussians, depth-sorted by ocrf_bin_forward */
  size_t vis_vals;      /* uint32 [V*P] their global index view*P + i */
  size_t vis_keys_tmp;  /* ping-pong halves of that sort */
  size_t vis_vals_tmp;
  size_t view_start;    /* uint32 [V+1] first visible slot of every view */
  size_t vis_sort_ws;   /* workspace of the visible-Gaussian sort */
} OcrfGeomLayout;

typedef struct OcrfBinLayout {
  size_t total;
  size_t keys;          /* uint64 [N] sorted */
  size_t point_list;    /* uint32 [N] sorted Gaussian ids */
  size_t keys_tmp;      /* uint64 [N] the other half of the sort's ping-pong */
  size_t vals_tmp;      /* uint32 [N] */
  size_t keys_unsorted; /* where duplicateWithKeys writes: keys_tmp when the pass count is odd, else keys */
  size_t vals_unsorted; /* (transient: overwritten by the sort) */
  size_t records;       /* OCRF_RECORD_BYTES * N: per tile, the records that can reach a pixel of the tile */
  size_t histogram;     /* uint32 [8][256] */
  size_t sort_status;   /* uint32 [passes][sort_tiles][256] + tickets */
  size_t split_counts;  /* depth-first: scanned tiles_touched; multi-split: uint32 [2][V][chunks][tiles] chunk x tile counts */
  size_t split_tiles;   /* depth-first: look-back state; multi-split: uint32 [3][V*tiles] totals (full, kept), offsets */
  size_t split_words;   /* capacity of split_counts in 32-bit words */
  size_t split_total;   /* bytes the default (multi-split) mode of ocrf_bin_forward touches: records, keys (its per-pair
                           items), split_counts, split_tiles -- laid out first; `total` covers every mode */
} OcrfBinLayout;

typedef struct OcrfImageLayout {
  size_t total;
  size_t ranges;    /* uint2  [V*tiles] [first, last+1) of the tile in the sorted list (the reference's ranges) */
  size_t ranges_render; /* uint2 [V*tiles] [first, last+1) of the tile's culled records */
  size_t final_T;   /* float  [V*H*W] */
  size_t n_contrib; /* uint32 [V*H*W] */
  size_t max_contrib; /* uint32 [V*tiles] largest n_contrib of the tile (lets backward skip the tail) */
} OcrfImageLayout;

int ocrf_abi_version(void);
const char* ocrf_error_string(int code);
/* The reference's debug mode (GaussianRasterizationSettings.debug -> CHECK_CUDA, cuda_rasterizer/auxiliary.h:166-173):
 * wait for everything queued on `stream` and return the first CUDA error, sticky or asynchronous, as a status code. */
int ocrf_debug_sync(void* stream);

int ocrf_geom_layout(const OcrfShape* shape, int use_sh, OcrfGeomLayout* out);
int ocrf_bin_layout(const OcrfShape* shape, uint64_t num_pairs, OcrfBinLayout* out);
int ocrf_image_layout(const OcrfShape* shape, OcrfImageLayout* out);
/* number of key bits the sort covers: 32 + getHigherMsb(V * tiles) (rasterizer_impl.cu:35-50,300) */
int ocrf_sort_end_bit(const OcrfShape* shape);

/* Stage 1 (forward.cu:155-256 + the InclusiveSum at rasterizer_impl.cu:277).
 * means3D [S,P,3], scales [S,P,3] / rotations [S,P,4] or cov3D_precomp [S,P,6], opacities [S,P],
 * shs [S,P,M,3] or NULL, cams [V,OCRF_CAM_STRIDE].  Writes radii [V,P] and the geom workspace
 * (header[0] = total number of (tile, Gaussian) pairs of the batch). */
int ocrf_preprocess_forward(void* stream, const OcrfShape* shape, const float* means3D, const float* scales,
                            const float* rotations, const float* cov3D_precomp, const float* opacities,
                            const float* shs, const float* cams, float scale_modifier, int prefiltered,
                            int32_t* radii, void* geom_ws);
/* The same with a foreground filter (OcRFDet turns every voxel of its 13 x 128 x 128 grid into a
 * Gaussian, view_transformer_ocrf.py:1130-1153, most of them empty space): Gaussians with opacity < min_opacity are
 * culled like out-of-frustum ones (radii 0, no pairs).  min_opacity is clamped to 1/255, below which a Gaussian can
 * never pass the blend's `alpha < 1/255 -> continue` (forward.cu:345, backward.cu:478): images and gradients are
 * unchanged, only radii / keys / pair counts shrink.  min_opacity <= 0 is ocrf_preprocess_forward. */
int ocrf_preprocess_forward_filtered(void* stream, const OcrfShape* shape, const float* means3D, const float* scales,
                                     const float* rotations, const float* cov3D_precomp, const float* opacities,
                                     const float* shs, const float* cams, float scale_modifier, int prefiltered,
                                     float min_opacity, int32_t* radii, void* geom_ws);

/* Stage 2 (rasterizer_impl.cu:70-138,289-318): duplicateWithKeys + onesweep radix sort of the
 * 64-bit (view*tiles+tile | depth) keys + identifyTileRanges + record packing.
 * `pair_capacity` is the number of pairs the binning workspace was laid out for; the true count is
 * read on the device from the geom header, and exceeding the capacity sets header[1] bit 0 and
 * renders nothing rather than overrunning.  colors [S,P,C] (ignored when the SH path produced rgb).
 * `sticky_status` (may be NULL): uint32[2] owned by the caller and never cleared by the library; [0] receives (OR) the
 * error flags of this call (bit 0 = pair overflow, bit 1 = a filtered point although `prefiltered` was set), [1]
 * (max) its pair count -- the geom header itself is re-zeroed by every forward, so callers that do not read it back
 * after each call (capacity mode, CUDA-graph replays) watch these words instead. */
int ocrf_bin_forward(void* stream, const OcrfShape* shape, uint64_t pair_capacity, const int32_t* radii,
                     const float* colors, int use_sh, uint32_t flags, void* geom_ws, void* bin_ws, void* image_ws,
                     uint32_t* sticky_status);
/* flags for ocrf_bin_forward */
#define OCRF_BIN_PAIR_SORT 1u   /* the reference's algorithm: sort all (tile | depth) pairs */
#define OCRF_BIN_DEPTH_FIRST 2u /* sort the visible Gaussians by depth, emit pairs in that order, sort the tile bits only */
/* default (0): depth-sort the visible Gaussians, then ONE stable multi-split of the pair stream by tile that writes
 * the culled records directly; the key / point lists are not materialised on that path.  All three produce
 * bit-identical records and range tables. */

/* Stage 3 (forward.cu:261-374; median depth per the w-depth fork; opacity = 1 - final_T).
 * bg [C]; out_color [V,C,H,W]; out_depth, out_opacity [V,1,H,W] (either may be NULL). */
int ocrf_render_forward(void* stream, const OcrfShape* shape, uint64_t pair_capacity, const float* colors,
                        int use_sh, const float* bg, const void* geom_ws, const void* bin_ws, void* image_ws,
                        float* out_color, float* out_depth, float* out_opacity);

/* Stage 4 (backward.cu:399-557).  dL_dcolor [V,C,H,W]; dL_dopacity_map [V,1,H,W] or NULL.
 * Accumulates into ggrad [V,P,OCRF_GGRAD_STRIDE] and dL_dcolors [S,P,C] (or [V,P,3] when use_sh);
 * both must be zeroed by the caller (as torch::zeros in rasterize_points.cu:151-159), or prepared with
 * ocrf_clear_gradients. */
int ocrf_render_backward(void* stream, const OcrfShape* shape, uint64_t pair_capacity, const float* colors,
                         int use_sh, const float* bg, const void* geom_ws, const void* bin_ws,
                         const void* image_ws, const float* dL_dcolor, const float* dL_dopacity_map,
                         double* ggrad, float* dL_dcolors);

/* Clears the accumulators of ocrf_render_backward in one launch: the ggrad rows of visible (radii > 0) pairs --
 * rows of invisible pairs are neither read nor written by the two backward stages -- and all of dL_dcolors
 * ([S,P,C], or [V,P,3] when use_sh).  Equivalent to zero-filling both buffers (rasterize_points.cu:151-159).
 * Both pointers 16-byte aligned. */
int ocrf_clear_gradients(void* stream, const OcrfShape* shape, int use_sh, const int32_t* radii, double* ggrad,
                         float* dL_dcolors);

/* backward.cu:144-396: screen-space gradients -> dL_dmeans3D [S,P,3], dL_dmeans2D [V,P,3],
 * dL_dopacities [S,P], dL_dscales [S,P,3] + dL_drotations [S,P,4] (or dL_dcov3D [S,P,6] when
 * cov3D_precomp was given), dL_dshs [S,P,M,3] (SH path; dL_dcolors_view is then the [V,P,3] buffer
 * written by ocrf_render_backward).  Sums over the views of a sample happen here. */
int ocrf_preprocess_backward(void* stream, const OcrfShape* shape, const float* means3D, const float* scales,
                             const float* rotations, const float* cov3D_precomp, const float* shs,
                             const float* cams, float scale_modifier, const int32_t* radii, const void* geom_ws,
                             const double* ggrad, const float* dL_dcolors_view, float* dL_dmeans3D,
                             float* dL_dmeans2D, float* dL_dopacities, float* dL_dscales, float* dL_drotations,
                             float* dL_dcov3D, float* dL_dshs);

/* rasterizer_impl.cu:54-66,141-152.  means3D [P,3]; view/proj column-major 4x4; present uint8 [P]. */
int ocrf_mark_visible(void* stream, int32_t P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, uint8_t* present);

/* Stand-alone pieces of stage 2, exposed for tests and for callers that bring their own keys:
 * stable ascending sort of (uint64 key, uint32 value) pairs on key bits [0, end_bit).
 * ws must hold ocrf_sort_workspace_bytes(n); keys_tmp/vals_tmp are ping-pong buffers of n elements.
 * The sorted result is left in keys_out/vals_out. */
size_t ocrf_sort_workspace_bytes(uint64_t n);
int ocrf_sort_pairs(void* stream, uint64_t n, int end_bit, const uint64_t* keys_in, const uint32_t* vals_in,
                    uint64_t* keys_out, uint32_t* vals_out, uint64_t* keys_tmp, uint32_t* vals_tmp, void* ws);

/* Stage 5: mask = sigmoid(conv KxK([mean_c x, max_c x]) + opacity_bev); out = x * mask.
 * x, out [B,C,H,W]; w [2,K,K]; opacity_bev, mask [B,1,H,W]; stats [B,2,H,W] (kept for backward). */
int ocrf_opacity_mask_forward(void* stream, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K, const float* x,
                              const float* w, const float* opacity_bev, float* out, float* mask, float* stats);
/* g_x [B,C,H,W], g_opacity_bev [B,1,H,W] are written; g_w [2,K,K] is ACCUMULATED (zero it first);
 * scratch holds B*3*H*W floats. */
int ocrf_opacity_mask_backward(void* stream, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K, const float* x,
                               const float* w, const float* mask, const float* stats, const float* g_out,
                               float* g_x, float* g_w, float* g_opacity_bev, float* scratch);

/* Stage 5a: the height-aware opacity lift proper (view_transformer_ocrf.py:1159-1161),
 *   opacity_alpha = interpolate(DeformableAttention2D(interpolate(opacity, /6), interpolate(alpha_lidar, /6)), full) + opacity
 * with DeformableAttention2D of mmdet3d/ops/cross_attention_2d.py:93-220 in OcRFDet's configuration
 * (view_transformer_ocrf.py:639-648: dim 13, one head of 8, one offset group, 6x6 stride-4 offset conv, scale 4).
 * opacity, alpha, out [B,13,H,W] (bilinear resizes with align_corners=True, coarse size (H/6, W/6));
 * params: the module's OCRF_HOA_ATTN_PARAMS parameters concatenated in named_parameters() order
 *   (to_offsets.0.weight, .0.bias, to_offsets.2.weight, rel_pos_bias.mlp.{0.0,1.0,2}.{weight,bias}, to_q, to_k, to_v,
 *    to_out.weight, to_out.bias);
 * keep (may be NULL): dropout keep-mask [B, (H/6)(W/6), keys] already divided by 1 - p (the reference's p = 0.1);
 * workspace: ocrf_hoa_lift_workspace_floats(...) floats, shared by forward and backward (no state is carried).
 * Backward writes g_opacity, g_alpha [B,13,H,W] and ACCUMULATES g_params (zero it first). */
#define OCRF_HOA_ATTN_PARAMS 766
size_t ocrf_hoa_lift_workspace_floats(int32_t B, int32_t dim, int32_t H, int32_t W);
int ocrf_hoa_lift_forward(void* stream, int32_t B, int32_t dim, int32_t H, int32_t W, const float* opacity,
                          const float* alpha, const float* params, const float* keep, float* out, float* workspace);
int ocrf_hoa_lift_backward(void* stream, int32_t B, int32_t dim, int32_t H, int32_t W, const float* opacity,
                           const float* alpha, const float* params, const float* keep, const float* g_out,
                           float* g_opacity, float* g_alpha, float* g_params, float* workspace);

/* Stage 5b: OpacityVoxelToBEVConverter + HeightAttention (view_transformer_ocrf.py:421-518, called at :1196):
 * x [B,13,S,S] (the lifted opacity), position [1 or B,4,S,S] -> out [B,1,S,S] (S % 4 == 0).
 * params: the module's OCRF_HOA_CONVERTER_PARAMS parameters concatenated in named_parameters() order.
 * train != 0: batch-norm batch statistics (biased variance); batch_stats (may be NULL) receives, per block and channel,
 *   [5][16][2] = (sum, sum of squares) over B*h*w, from which the caller updates its running statistics;
 * train == 0: running_mean / running_var [5][16] (block-major, padded to 16 channels).
 * workspace: ocrf_hoa_converter_workspace_floats(B, S) floats; the backward needs the workspace of ITS forward, untouched.
 * Backward writes g_x [B,13,S,S] and g_position, and ACCUMULATES g_params (zero it first). */
#define OCRF_HOA_CONVERTER_PARAMS 1847
size_t ocrf_hoa_converter_workspace_floats(int32_t B, int32_t S);
int ocrf_hoa_converter_forward(void* stream, int32_t B, int32_t S, int32_t train, const float* x, const float* position,
                               int32_t position_batched, const float* params, const float* running_mean,
                               const float* running_var, float* out, float* batch_stats, float* workspace);
int ocrf_hoa_converter_backward(void* stream, int32_t B, int32_t S, int32_t train, const float* x, const float* position,
                                int32_t position_batched, const float* params, const float* running_mean,
                                const float* running_var, const float* g_out, float* g_x, float* g_position,
                                float* g_params, float* workspace);

/* Stage 0 (scope row a12, "next" f-1): OcRF Gaussian construction, the four MLP heads of
 * view_transformer_ocrf.py:272-320 evaluated at :1130-1133, in one pass over the voxel features.
 * feat [n,F] (F <= 125), rgb [n,3].  Packed parameters (input-major first layer):
 *   w1t [F+3,16]: columns 0-3 S_MLP.fc1, 4-7 R_MLP.fc1, 8-11 A_MLP.fc1 (their 3 rgb rows zero), 12-15 C_MLP.fc1
 *   b1 [16];  w2 [11,4]: rows 0-2 S_MLP.fc2, 3-6 R_MLP.fc2, 7 A_MLP.fc2, 8-10 C_MLP.fc2;  b2 [11].
 * Outputs: opacity [n] (sigmoid), scales [n,3] (softplus), rotations [n,4] (unit), colors [n,3] (sigmoid),
 * hidden [n,16] (post-ReLU, kept for backward).  rotations and hidden must be 16-byte aligned; feat 16-byte
 * aligned with F % 4 == 0 takes the bulk-copy path, anything else a slower plain-load path. */
int ocrf_gaussian_heads_forward(void* stream, int64_t n, int32_t F, const float* feat, const float* rgb,
                                const float* w1t, const float* b1, const float* w2, const float* b2, float* opacity,
                                float* scales, float* rotations, float* colors, float* hidden);
/* g_feat [n,F] is written; g_w1t [F+3,16], g_b1 [16], g_w2 [11,4], g_b2 [11] are ACCUMULATED (zero them first).
 * ws: ocrf_gaussian_heads_backward_workspace_bytes(n) bytes of scratch (dL/dhidden between the two kernels of the
 * fast path); may be NULL, then (as for F % 4 != 0 or unaligned feat) a single slower kernel runs. */
size_t ocrf_gaussian_heads_backward_workspace_bytes(int64_t n);
int ocrf_gaussian_heads_backward(void* stream, int64_t n, int32_t F, const float* feat, const float* rgb,
                                 const float* w1t, const float* w2, const float* b2, const float* hidden,
                                 const float* g_opacity, const float* g_scales, const float* g_rotations,
                                 const float* g_colors, float* g_feat, float* g_w1t, float* g_b1, float* g_w2,
                                 float* g_b2, void* ws);

/* "Next" row f-3: BEV pooling v2 (mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-121, bound at
 * bev_pool_v2/src/bev_pool.cpp and called from bev_pool_v2/bev_pool.py:18-79).
 * depth: flat floats indexed by ranks_depth; feat [n_feat, c]; out [n_bev, c] zero-filled by the caller (as
 * feat.new_zeros in bev_pool.py:29); the rank / interval arrays are the reference's int32 tensors, points sorted by
 * ranks_bev.  out[ranks_bev[start], :] = sum over the interval of depth * feat row (same summation order as the
 * reference kernel). */
int ocrf_bev_pool_forward(void* stream, int32_t c, int32_t n_intervals, const float* depth, const float* feat,
                          const int32_t* ranks_depth, const int32_t* ranks_feat, const int32_t* ranks_bev,
                          const int32_t* interval_starts, const int32_t* interval_lengths, float* out);
/* Backward, INCLUDING the regrouping of the points by feature pixel that the reference does in Python before its kernel
 * (bev_pool.py:47-60): pass the forward's rank arrays as they are (n_points entries, any order).  c <= 128.
 * depth_grad (same extent as depth) must be zero-filled by the caller; feat_grad [n_feat, c] is written completely.
 * ws holds ocrf_bev_pool_backward_workspace_bytes(n_points). */
size_t ocrf_bev_pool_backward_workspace_bytes(uint64_t n_points);
int ocrf_bev_pool_backward(void* stream, int32_t c, uint64_t n_points, int32_t n_feat, const float* out_grad,
                           const float* depth, const float* feat, const int32_t* ranks_depth, const int32_t* ranks_feat,
                           const int32_t* ranks_bev, float* depth_grad, float* feat_grad, void* ws);

/* "Next" row f-4 (callers upstream of the Gaussian heads).
 * ocrf_color_voxels = lidar_points_to_image_values + color_voxels (view_transformer_ocrf.py:924-971): avg [B,M,C] =
 * mean over the cameras n with mask[b,n,m] != 0 of the bilinear sample (grid_sample, align_corners=True, zero padding)
 * of imgs [B,N,C,H,W] at pixel coordinates coords [B,N,M,2] (x,y); 0 where no camera sees the voxel; the mean is
 * divided by `divisor` when it is not 1 (the caller's / 255.0 at :1071).  valid [B,M] (may be NULL) = any(mask).
 * C <= 4; coords 8-byte aligned. */
int ocrf_color_voxels(void* stream, int32_t B, int32_t N, int64_t M, int32_t C, int32_t H, int32_t W, const float* coords,
                      const uint8_t* mask, const float* imgs, float divisor, float* avg, uint8_t* valid);
/* ocrf_retain_valid_pixels = retain_valid_pixels (view_transformer_ocrf.py:1004-1022): out [V,C,H,W] = `fill` except
 * at the pixels (trunc(x), trunc(y)), clamped to [0, max(W,H)-1], of the points with mask != 0 and x != -1, which
 * keep img's value.  coords [V,M,2], mask [V,M]; out 16-byte aligned. */
int ocrf_retain_valid_pixels(void* stream, int32_t V, int64_t M, int32_t C, int32_t H, int32_t W, const float* coords,
                             const uint8_t* mask, const float* img, float fill, float* out);

#ifdef __cplusplus
}
#endif
#endif /* OCRF_RASTER_H */
