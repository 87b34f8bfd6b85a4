// Stage 2: (tile | depth) key generation, sort, tile ranges and record packing.
//
// Replaces duplicateWithKeys (cuda_rasterizer/rasterizer_impl.cu:70-111), the CUB sort call
// (:303-308), the ranges memset (:310) and identifyTileRanges (:116-138).
//
// Batch extension: the key's upper word is view * tiles + tile, so one sort orders every view of
// the batch; for V == 1 the keys are exactly the reference's.  After the sort a single pass finds
// the tile boundaries AND gathers each pair's blend inputs (mean, conic, opacity, depth, colour)
// into a contiguous 48-byte record, so that both blend kernels stream their tile's list with bulk
// async copies instead of two dependent gathers per pair.
#include <cstdlib>

#include "common.cuh"

namespace ocrf {

__device__ __forceinline__ void tile_rect_dev(float px, float py, int radius, int gx, int gy, int& x0, int& y0,
                                              int& x1, int& y1) {
  const float r = (float)radius;
  x0 = min(gx, max(0, (int)__fmul_rn(__fsub_rn(px, r), 0.0625f)));
  y0 = min(gy, max(0, (int)__fmul_rn(__fsub_rn(py, r), 0.0625f)));
  x1 = min(gx, max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(px, r), 16.f), 1.f), 0.0625f)));
  y1 = min(gy, max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(py, r), 16.f), 1.f), 0.0625f)));
}

// One thread per (view, Gaussian); rectangles larger than a warp are emitted by the whole warp.
__global__ void __launch_bounds__(256) duplicate_with_keys_kernel(OcrfShape sh, uint64_t n_cap,
                                                                  const int32_t* __restrict__ radii,
                                                                  uint32_t* __restrict__ header,
                                                                  const float* __restrict__ depths,
                                                                  const float2* __restrict__ xy,
                                                                  const uint32_t* __restrict__ offsets,
                                                                  uint64_t* __restrict__ keys,
                                                                  uint32_t* __restrict__ vals) {
  const uint32_t total = header[HDR_NUM_PAIRS];
  if ((uint64_t)total > n_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&header[HDR_ERROR], ERR_PAIR_OVERFLOW);
    return;
  }
  const size_t n = (size_t)sh.V * sh.P;
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int gx = ceil_div(sh.W, TILE), gy = ceil_div(sh.H, TILE);
  const int lane = threadIdx.x & 31;

  int x0 = 0, y0 = 0, w = 0, cnt = 0;
  uint32_t off = 0, dbits = 0, id = 0, tile_base = 0;
  if (g < n && radii[g] > 0) {
    const int v = (int)(g / sh.P);
    id = (uint32_t)(g - (size_t)v * sh.P);
    off = g == 0 ? 0u : offsets[g - 1];
    const float2 p = xy[g];
    int x1, y1;
    tile_rect_dev(p.x, p.y, radii[g], gx, gy, x0, y0, x1, y1);
    w = x1 - x0;
    cnt = w * (y1 - y0);
    dbits = __float_as_uint(depths[g]);
    tile_base = (uint32_t)v * (uint32_t)(gx * gy);
  }
  // small rectangles: serial, row-major (the reference's emission order)
  if (cnt > 0 && cnt <= 32) {
    for (int k = 0; k < cnt; k++) {
      const int ty = y0 + k / w, tx = x0 + k - (k / w) * w;
      keys[off + k] = ((uint64_t)(tile_base + (uint32_t)(ty * gx + tx)) << 32) | dbits;
      vals[off + k] = id;
    }
  }
  // large rectangles: the warp shares the work, tile k -> lane k % 32
  uint32_t big = __ballot_sync(0xffffffffu, cnt > 32);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
    const int bw = __shfl_sync(0xffffffffu, w, src), bcnt = __shfl_sync(0xffffffffu, cnt, src);
    const uint32_t boff = __shfl_sync(0xffffffffu, off, src), bd = __shfl_sync(0xffffffffu, dbits, src);
    const uint32_t bid = __shfl_sync(0xffffffffu, id, src), bt = __shfl_sync(0xffffffffu, tile_base, src);
    for (int k = lane; k < bcnt; k += 32) {
      const int ty = by0 + k / bw, tx = bx0 + k - (k / bw) * bw;
      keys[boff + k] = ((uint64_t)(bt + (uint32_t)(ty * gx + tx)) << 32) | bd;
      vals[boff + k] = bid;
    }
  }
}

// One CTA per (view, tile): finds the tile's range in the sorted keys by binary search (replacing the
// reference's memset + identifyTileRanges pass), then gathers each pair's blend inputs into 48-byte
// records -- but only for Gaussians that can reach a pixel of THIS tile.  The reference's tile
// rectangle is the bounding box of a 3-sigma circle of the LARGEST eigenvalue; about half of the
// pairs it emits have alpha < 1/255 on every pixel of the tile.  They stay in the sorted list (keys,
// point list and ranges remain bit-identical to the reference) but are dropped from the render
// records by an exact-safe bound: the minimum of the conic's quadratic form over the tile's pixel
// rectangle (continuous relaxation, evaluated on the four edges) gives the largest possible alpha;
// a record is dropped only if that bound, with a 1e-3 relative margin for rounding, is below 1/255.
// Each record keeps its 1-based position in the full list so n_contrib is the reference's value.
// Can Gaussian (mean p, conic A/B/C, opacity co.w) reach alpha >= 1/255 on any pixel of the rectangle
// [px0,px1] x [py0,py1]?  Exact-safe: false only if the continuous minimum of the conic form over the
// rectangle bounds alpha below 1/255 with a 1e-3 relative margin (NaN anywhere answers true).
__device__ __forceinline__ bool tile_can_contribute(float2 p, float4 co, float px0, float px1, float py0, float py1) {
  const float dxl = p.x - px1, dxh = p.x - px0, dyl = p.y - py1, dyh = p.y - py0;  // d = mean - pixel
  float qmin = 0.f;
  const bool inside = dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f;
  if (!inside) {
    const float A = co.x, B = co.y, Cc = co.z;
    float q = INFINITY;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float ex = e ? dxh : dxl;
      const float sy = fminf(dyh, fmaxf(dyl, __fdividef(-B * ex, Cc)));
      q = fminf(q, A * ex * ex + 2.f * B * ex * sy + Cc * sy * sy);
      const float ey = e ? dyh : dyl;
      const float sx = fminf(dxh, fmaxf(dxl, __fdividef(-B * ey, A)));
      q = fminf(q, A * sx * sx + 2.f * B * sx * ey + Cc * ey * ey);
    }
    qmin = q;
  }
  const float alpha_max = co.w * __expf(-0.5f * qmin) * 1.001f;
  return !(alpha_max < 1.0f / 255.0f);
}

__global__ void __launch_bounds__(256) ranges_cull_pack_kernel(OcrfShape sh, uint64_t n_cap, int use_sh, int has_rgb,
                                                               const uint32_t* __restrict__ header,
                                                               const uint64_t* __restrict__ keys,
                                                               const uint32_t* __restrict__ point_list,
                                                               const float* __restrict__ depths,
                                                               const float2* __restrict__ xy,
                                                               const float4* __restrict__ conic_opacity,
                                                               const float* __restrict__ rgb,
                                                               const float* __restrict__ colors,
                                                               uint2* __restrict__ ranges,
                                                               uint2* __restrict__ ranges_render,
                                                               Record* __restrict__ records,
                                                               uint32_t* __restrict__ sticky) {
  __shared__ uint32_t s_bounds[2];
  if (sticky != nullptr && blockIdx.x == 0 && threadIdx.x == 0) publish_status(header, n_cap, sticky);
  __shared__ uint32_t s_warp[8];
  const uint32_t total = header[HDR_NUM_PAIRS];
  const uint32_t n = (uint64_t)total <= n_cap ? total : 0u;
  const uint32_t vt = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 2) {  // lower bound of vt (tid 0) and of vt + 1 (tid 1) in the keys' upper words
    const uint32_t target = vt + tid;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if ((uint32_t)(keys[mid] >> 32) < target) lo = mid + 1; else hi = mid;
    }
    s_bounds[tid] = lo;
  }
  __syncthreads();
  const uint32_t lo = s_bounds[0], hi = s_bounds[1];
  if (tid == 0) ranges[vt] = lo < hi ? make_uint2(lo, hi) : make_uint2(0u, 0u);  // empty tiles read (0,0) as after the memset

  const int gx = ceil_div(sh.W, TILE), tiles = gx * ceil_div(sh.H, TILE);
  const uint32_t v = vt / (uint32_t)tiles;
  const int t = (int)(vt - v * tiles);
  const int tx = t % gx, ty = t / gx;
  const float px0 = (float)(tx * TILE), px1 = (float)min(tx * TILE + TILE - 1, sh.W - 1);
  const float py0 = (float)(ty * TILE), py1 = (float)min(ty * TILE + TILE - 1, sh.H - 1);
  const size_t sample_base = (size_t)(v / sh.views_per_sample) * sh.P;

  uint32_t kept_total = 0;
  for (uint32_t base = lo; base < hi; base += 256) {
    const uint32_t i = base + tid;
    bool keep = false;
    uint32_t id = 0;
    float2 p = make_float2(0.f, 0.f);
    float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
    size_t g = 0;
    if (i < hi) {
      id = point_list[i];
      g = (size_t)v * sh.P + id;
      p = xy[g];
      co = conic_opacity[g];
      keep = tile_can_contribute(p, co, px0, px1, py0, py1);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = 0, chunk_total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const uint32_t c = s_warp[w];
      if (w < warp) off += c;
      chunk_total += c;
    }
    if (keep) {
      const uint32_t dst = lo + kept_total + off + __popc(bal & ((1u << lane) - 1));
      float r = 0.f, gg = 0.f, bb = 0.f;
      if (has_rgb) {
        const float* c = use_sh ? rgb + g * 3 : colors + (sample_base + id) * 3;
        r = __ldg(c); gg = __ldg(c + 1); bb = __ldg(c + 2);
      }
      float4* out = reinterpret_cast<float4*>(records + dst);
      out[0] = record_head(p, co);
      out[1] = make_float4(record_qc(co), co.w, __uint_as_float(i - lo + 1), r);
      out[2] = make_float4(gg, bb, __uint_as_float(id), depths[g]);
    }
    kept_total += chunk_total;
    __syncthreads();
  }
  if (tid == 0) ranges_render[vt] = make_uint2(lo, lo + kept_total);
}

// ------------------------------------------------------------------------------------------------
// Depth-first binning: the default stage-2 algorithm.
//
// The reference emits its (tile | depth) pairs in Gaussian-index order and needs ceil((32+bits)/8)
// = 6 radix passes over all N_dup pairs (3.85 M at the headline shape, 134 M at the 1 M-Gaussian
// stress shape).  Here the ~19 k VISIBLE Gaussians of every view are first sorted by (view | depth)
// with the same onesweep sort (ties keep Gaussian-index order, as in the stable pair sort), the pairs
// are then emitted in THAT order, and only the tile bits [32, 32+bits) remain to be sorted: 2 stable
// passes instead of 6.  A stable sort by tile of a depth-ordered sequence is exactly the reference's
// order, so keys, point list and ranges are bit-identical (tests run both algorithms).
// ------------------------------------------------------------------------------------------------
// inclusive scan of tiles_touched over the depth-sorted visible Gaussians (decoupled look-back)
__global__ void __launch_bounds__(256) scan_sorted_tiles_kernel(const uint32_t* __restrict__ header,
                                                                const uint32_t* __restrict__ vis_vals,
                                                                const uint32_t* __restrict__ tiles_touched,
                                                                uint32_t* __restrict__ sorted_offsets,
                                                                unsigned long long* __restrict__ status,
                                                                uint32_t* __restrict__ ticket) {
  pdl_enter();
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_bid, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = header[HDR_NUM_VIS];
  if (tid == 0) s_bid = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t bid = s_bid;
  if ((uint64_t)bid * 1024 >= n) return;
  uint32_t x[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t j = bid * 1024 + tid * 4 + k;
    x[k] = j < n ? tiles_touched[vis_vals[j]] : 0u;
    sum += x[k];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t warp_off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const uint32_t y = s_warp[w];
    if (w < warp) warp_off += y;
    total += y;
  }
  if (warp == 0) {
    const unsigned long long excl = lookback_warp(status, (int)bid, total);
    if (lane == 0) s_prefix = (uint32_t)excl;
  }
  __syncthreads();
  uint32_t run = s_prefix + warp_off + incl - sum;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t j = bid * 1024 + tid * 4 + k;
    run += x[k];
    if (j < n) sorted_offsets[j] = run;
  }
}

// duplicateWithKeys over the depth-sorted order: one thread per sorted Gaussian, large rectangles by the warp
__global__ void __launch_bounds__(256) duplicate_sorted_kernel(OcrfShape sh, uint64_t n_cap,
                                                               const int32_t* __restrict__ radii,
                                                               uint32_t* __restrict__ header,
                                                               const uint64_t* __restrict__ vis_keys,
                                                               const uint32_t* __restrict__ vis_vals,
                                                               const float2* __restrict__ xy,
                                                               const uint32_t* __restrict__ sorted_offsets,
                                                               uint64_t* __restrict__ keys,
                                                               uint32_t* __restrict__ vals) {
  const uint32_t total = header[HDR_NUM_PAIRS];
  if ((uint64_t)total > n_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&header[HDR_ERROR], ERR_PAIR_OVERFLOW);
    return;
  }
  const uint32_t n = header[HDR_NUM_VIS];
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const int gx = ceil_div(sh.W, TILE), gy = ceil_div(sh.H, TILE);
  const int lane = threadIdx.x & 31;
  int x0 = 0, y0 = 0, w = 0, cnt = 0;
  uint32_t off = 0, dbits = 0, id = 0, tile_base = 0;
  if (j < n) {
    const uint32_t g = vis_vals[j];
    const uint64_t vk = vis_keys[j];
    const uint32_t v = (uint32_t)(vk >> 32);
    id = g - v * (uint32_t)sh.P;
    dbits = (uint32_t)vk;
    off = j == 0 ? 0u : sorted_offsets[j - 1];
    const float2 p = xy[g];
    int x1, y1;
    tile_rect_dev(p.x, p.y, radii[g], gx, gy, x0, y0, x1, y1);
    w = x1 - x0;
    cnt = w * (y1 - y0);
    tile_base = v * (uint32_t)(gx * gy);
  }
  if (cnt > 0 && cnt <= 32) {
    for (int k = 0; k < cnt; k++) {
      const int ty = y0 + k / w, tx = x0 + k - (k / w) * w;
      keys[off + k] = ((uint64_t)(tile_base + (uint32_t)(ty * gx + tx)) << 32) | dbits;
      vals[off + k] = id;
    }
  }
  uint32_t big = __ballot_sync(0xffffffffu, cnt > 32);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
    const int bw = __shfl_sync(0xffffffffu, w, src), bcnt = __shfl_sync(0xffffffffu, cnt, src);
    const uint32_t boff = __shfl_sync(0xffffffffu, off, src), bd = __shfl_sync(0xffffffffu, dbits, src);
    const uint32_t bid = __shfl_sync(0xffffffffu, id, src), bt = __shfl_sync(0xffffffffu, tile_base, src);
    for (int k = lane; k < bcnt; k += 32) {
      const int ty = by0 + k / bw, tx = bx0 + k - (k / bw) * bw;
      keys[boff + k] = ((uint64_t)(bt + (uint32_t)(ty * gx + tx)) << 32) | bd;
      vals[boff + k] = bid;
    }
  }
}

}  // namespace ocrf

using namespace ocrf;

static int higher_msb(uint32_t n) {  // rasterizer_impl.cu:35-50
  int b = 0;
  while (b < 32 && (n >> b)) b++;
  return b == 0 ? 1 : b;
}

extern "C" int ocrf_sort_end_bit(const OcrfShape* sh) {
  if (!sh) return OCRF_EINVAL;
  return 32 + higher_msb((uint32_t)sh->V * (uint32_t)(tiles_x(*sh) * tiles_y(*sh)));
}

extern "C" int ocrf_bin_layout(const OcrfShape* sh, uint64_t num_pairs, OcrfBinLayout* out) {
  if (!sh || !out) return OCRF_EINVAL;
  if (num_pairs >= (1ull << 30)) return OCRF_ECAPACITY;
  const size_t n = num_pairs ? num_pairs : 1;
  const int passes = (ocrf_sort_end_bit(sh) + 7) / 8;
  // What the default (multi-split) binning touches comes FIRST: the records, the per-pair (tile, Gaussian) items (in the
  // `keys` array), the chunk x tile count tables and the tile arrays end at `split_total`; a caller that uses only that
  // mode may allocate just that much (56 instead of 72 bytes per pair and no sort workspace).  The key / point lists
  // of the depth-first and pair-sort modes and the sort workspace follow.
  size_t off = 0;
  out->records = off;    off = align128(off + n * OCRF_RECORD_BYTES);
  out->keys = off;       off = align128(off + n * 8);
  const size_t nvp = (size_t)sh->V * (sh->P > 0 ? sh->P : 1);
  const size_t tiles_v = (size_t)tiles_x(*sh) * tiles_y(*sh);
  const uint32_t Q = multisplit_chunk_pairs(n);
  const size_t chunks_max = (n + Q - 1) / Q + 1;
  const size_t table_words = 2 * (size_t)sh->V * chunks_max * tiles_v;
  out->split_words = table_words;
  out->split_counts = off;  // [V*P] tiles_touched scanned in depth order | multi-split chunk x tile tables
  off = align128(off + (nvp + table_words) * 4);
  out->split_tiles = off;   // look-back state + ticket of that scan | [3][V*tiles] tile totals / offsets
  off = align128(off + align128(((nvp + 1023) / 1024 + 1) * 8 + 128) + 3 * (size_t)sh->V * tiles_v * 4);
  out->split_total = off + 128;
  out->keys_tmp = off;   off = align128(off + n * 8);
  out->point_list = off; off = align128(off + n * 4);
  out->vals_tmp = off;   off = align128(off + n * 4);
  // `passes` ping-pongs must end in (keys, point_list): start in tmp when odd, in keys when even
  out->keys_unsorted = (passes & 1) ? out->keys_tmp : out->keys;
  out->vals_unsorted = (passes & 1) ? out->vals_tmp : out->point_list;
  out->histogram = off;  // start of the sort workspace
  const SortWs w = sort_ws_layout(n);
  out->sort_status = off + w.status;
  off = align128(off + w.total + 128);
  out->total = off + 128;
  return 0;
}

extern "C" int ocrf_image_layout(const OcrfShape* sh, OcrfImageLayout* out) {
  if (!sh || !out) return OCRF_EINVAL;
  const size_t tiles = (size_t)sh->V * tiles_x(*sh) * tiles_y(*sh);
  const size_t pix = (size_t)sh->V * sh->W * sh->H;
  size_t off = 0;
  out->ranges = off;      off = align128(off + tiles * 8);
  out->ranges_render = off; off = align128(off + tiles * 8);
  out->max_contrib = off; off = align128(off + tiles * 4);
  out->final_T = off;     off = align128(off + pix * 4);
  out->n_contrib = off;   off = align128(off + pix * 4);
  out->total = off + 128;
  return 0;
}

extern "C" int ocrf_bin_forward(void* stream, const OcrfShape* sh, uint64_t pair_capacity, const int32_t* radii,
                                const float* colors, int use_sh, uint32_t flags, void* geom_ws, void* bin_ws,
                                void* image_ws, uint32_t* sticky_status) {
  if (!sh || !radii || !geom_ws || !bin_ws || !image_ws) return OCRF_EINVAL;
  if (sh->C == 3 && !use_sh && !colors) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OcrfGeomLayout G;
  OcrfBinLayout B;
  OcrfImageLayout I;
  ocrf_geom_layout(sh, use_sh, &G);
  int rc = ocrf_bin_layout(sh, pair_capacity, &B);
  if (rc) return rc;
  ocrf_image_layout(sh, &I);
  const size_t tiles = (size_t)sh->V * tiles_x(*sh) * tiles_y(*sh);
  if (pair_capacity == 0) {  // nothing can be rendered: both range tables read (0,0)
    cudaMemsetAsync(at<char>(image_ws, I.ranges), 0, I.max_contrib - I.ranges, st);
    return 0;
  }

  uint32_t* header = at<uint32_t>(geom_ws, G.header);
  const size_t n = (size_t)sh->V * sh->P;
  const int gxh = tiles_x(*sh), gyh = tiles_y(*sh), tiles_v = gxh * gyh;
  const bool pairsort = (flags & OCRF_BIN_PAIR_SORT) != 0;
  const bool multisplit = !pairsort && !(flags & OCRF_BIN_DEPTH_FIRST) && tiles_v <= 4096;
  (void)gxh; (void)gyh;
  if (!pairsort) {
    // (1) depth-sort the visible Gaussians of every view and (2) scan their tile counts in that order: where every
    //     Gaussian's pairs sit in the pair stream
    const uint32_t* vvals = at<uint32_t>(geom_ws, G.vis_vals);
    uint32_t* sorted_offsets = at<uint32_t>(bin_ws, B.split_counts);
    const size_t sblocks = (n + 1023) / 1024;
    if (vis_sort_onchip()) {
      // one cluster of 8 CTAs per view: four digit passes through distributed shared memory + the scan, one launch
      rc = visible_sort(st, sh, at<uint32_t>(geom_ws, G.view_start), at<uint64_t>(geom_ws, G.vis_keys),
                        at<uint32_t>(geom_ws, G.vis_vals), at<uint64_t>(geom_ws, G.vis_keys_tmp),
                        at<uint32_t>(geom_ws, G.vis_vals_tmp), at<uint32_t>(geom_ws, G.tiles_touched),
                        at<uint32_t>(geom_ws, G.offsets), sorted_offsets);
      if (rc) return rc;
    } else {
      // round 1: the global onesweep sort over (view | depth) + a look-back scan (kept for A/B measurements)
      const int vbit = vis_sort_end_bit(sh->V);
      const int vpasses = (vbit + 7) / 8;
      uint64_t* ka = at<uint64_t>(geom_ws, (vpasses & 1) ? G.vis_keys_tmp : G.vis_keys);
      uint32_t* va = at<uint32_t>(geom_ws, (vpasses & 1) ? G.vis_vals_tmp : G.vis_vals);
      uint64_t* kb = at<uint64_t>(geom_ws, (vpasses & 1) ? G.vis_keys : G.vis_keys_tmp);
      uint32_t* vb = at<uint32_t>(geom_ws, (vpasses & 1) ? G.vis_vals : G.vis_vals_tmp);
      // The sort workspace was zeroed together with the geom header (ocrf_preprocess_forward); the look-back state of
      // the scan below is cleared by the sort's histogram kernel: no memset node between preprocess and the blend.
      unsigned long long* sstat = at<unsigned long long>(bin_ws, B.split_tiles);
      uint32_t* sticket = reinterpret_cast<uint32_t*>(sstat + sblocks + 1);
      rc = sort_pairs_device(st, header + HDR_NUM_VIS, n, 0, vbit, ka, va, kb, vb, at<char>(geom_ws, G.vis_sort_ws), true,
                             reinterpret_cast<uint32_t*>(sstat), (uint32_t)(((sblocks + 1) * 8 + 64) / 4));
      if (rc) return rc;
      OCRF_LAUNCH(scan_sorted_tiles_kernel, dim3((unsigned)sblocks), dim3(256), 0, st, header, vvals,
                  at<uint32_t>(geom_ws, G.tiles_touched), sorted_offsets, sstat, sticket);
    }
    if (multisplit) {
      // (3) one stable multi-split of the pair stream by tile, culled records written directly
      uint32_t* tile_arrays = at<uint32_t>(bin_ws, B.split_tiles + align128((sblocks + 1) * 8 + 128));
      rc = multisplit_bin(st, sh, pair_capacity, use_sh, radii, colors, header, at<uint32_t>(geom_ws, G.view_start),
                          sorted_offsets, vvals, at<float2>(geom_ws, G.xy), at<float4>(geom_ws, G.conic_opacity),
                          at<float>(geom_ws, G.depths), at<float>(geom_ws, G.rgb), sorted_offsets + n, B.split_words,
                          tile_arrays, at<uint2>(bin_ws, B.keys), at<uint2>(image_ws, I.ranges),
                          at<uint2>(image_ws, I.ranges_render), at<Record>(bin_ws, B.records), sticky_status);
      return rc;
    }
    // (3') emit the pairs in depth order, then two stable passes over the tile bits only
    const int end_bit = ocrf_sort_end_bit(sh);
    const int tpasses = (end_bit - 32 + 7) / 8;
    uint64_t* pka = at<uint64_t>(bin_ws, (tpasses & 1) ? B.keys_tmp : B.keys);
    uint32_t* pva = at<uint32_t>(bin_ws, (tpasses & 1) ? B.vals_tmp : B.point_list);
    uint64_t* pkb = at<uint64_t>(bin_ws, (tpasses & 1) ? B.keys : B.keys_tmp);
    uint32_t* pvb = at<uint32_t>(bin_ws, (tpasses & 1) ? B.point_list : B.vals_tmp);
    duplicate_sorted_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        *sh, pair_capacity, radii, header, at<uint64_t>(geom_ws, G.vis_keys), vvals, at<float2>(geom_ws, G.xy),
        sorted_offsets, pka, pva);
    rc = sort_pairs_device(st, header + HDR_NUM_PAIRS, pair_capacity, 32, end_bit, pka, pva, pkb, pvb,
                           at<char>(bin_ws, B.histogram));
    if (rc) return rc;
  } else {
  uint32_t* header = at<uint32_t>(geom_ws, G.header);
    const size_t n = (size_t)sh->V * sh->P;
    duplicate_with_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        *sh, pair_capacity, radii, header, at<float>(geom_ws, G.depths), at<float2>(geom_ws, G.xy),
        at<uint32_t>(geom_ws, G.offsets), at<uint64_t>(bin_ws, B.keys_unsorted), at<uint32_t>(bin_ws, B.vals_unsorted));

    const int end_bit = ocrf_sort_end_bit(sh);
    const int passes = (end_bit + 7) / 8;
    uint64_t* ka = at<uint64_t>(bin_ws, B.keys_unsorted);
    uint32_t* va = at<uint32_t>(bin_ws, B.vals_unsorted);
    uint64_t* kb = at<uint64_t>(bin_ws, (passes & 1) ? B.keys : B.keys_tmp);
    uint32_t* vb = at<uint32_t>(bin_ws, (passes & 1) ? B.point_list : B.vals_tmp);
    rc = sort_pairs_device(st, header + HDR_NUM_PAIRS, pair_capacity, 0, end_bit, ka, va, kb, vb,
                           at<char>(bin_ws, B.histogram));
    if (rc) return rc;
  }

  ranges_cull_pack_kernel<<<(unsigned)tiles, 256, 0, st>>>(
      *sh, pair_capacity, use_sh, sh->C == 3, header, at<uint64_t>(bin_ws, B.keys), at<uint32_t>(bin_ws, B.point_list),
      at<float>(geom_ws, G.depths), at<float2>(geom_ws, G.xy), at<float4>(geom_ws, G.conic_opacity),
      at<float>(geom_ws, G.rgb), colors, at<uint2>(image_ws, I.ranges), at<uint2>(image_ws, I.ranges_render),
      at<Record>(bin_ws, B.records), sticky_status);
  OCRF_CHECK_LAST();
  return 0;
}
