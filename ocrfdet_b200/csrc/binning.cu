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

// Boundary detection on the sorted keys (ranges zeroed beforehand) fused with the record gather.
template <bool kLite>
__global__ void __launch_bounds__(256) ranges_and_pack_kernel(OcrfShape sh, uint64_t n_cap, int use_sh,
                                                              const uint32_t* __restrict__ header,
                                                              const uint64_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ point_list,
                                                              const float* __restrict__ depths,
                                                              const float2* __restrict__ xy,
                                                              const float4* __restrict__ conic_opacity,
                                                              const float* __restrict__ rgb,
                                                              const float* __restrict__ colors,
                                                              uint2* __restrict__ ranges, void* __restrict__ records) {
  const uint32_t total = header[HDR_NUM_PAIRS];
  const uint32_t n = (uint64_t)total <= n_cap ? total : 0u;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t cur = (uint32_t)(keys[i] >> 32);
  if (i == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
    if (cur != prev) {
      ranges[prev].y = i;
      ranges[cur].x = i;
    }
  }
  if (i == n - 1) ranges[cur].y = n;

  const int tiles = ceil_div(sh.W, TILE) * ceil_div(sh.H, TILE);
  const uint32_t v = cur / (uint32_t)tiles;
  const uint32_t id = point_list[i];
  const size_t g = (size_t)v * sh.P + id;
  const float2 p = xy[g];
  const float4 co = conic_opacity[g];
  const float d = depths[g];
  if (kLite) {
    float4* out = reinterpret_cast<float4*>(records) + (size_t)i * 2;
    out[0] = make_float4(p.x, p.y, co.x, co.y);
    out[1] = make_float4(co.z, co.w, d, __uint_as_float(id));
  } else {
    const float* c = use_sh ? rgb + g * 3 : colors + ((size_t)(v / sh.views_per_sample) * sh.P + id) * 3;
    float4* out = reinterpret_cast<float4*>(records) + (size_t)i * 3;
    out[0] = make_float4(p.x, p.y, co.x, co.y);
    out[1] = make_float4(co.z, co.w, d, __ldg(c));
    out[2] = make_float4(__ldg(c + 1), __ldg(c + 2), __uint_as_float(id), 0.f);
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
  size_t off = 0;
  out->keys = off;       off = align128(off + n * 8);
  out->keys_tmp = off;   off = align128(off + n * 8);
  out->point_list = off; off = align128(off + n * 4);
  out->vals_tmp = off;   off = align128(off + n * 4);
  // `passes` ping-pongs must end in (keys, point_list): start in tmp when odd, in keys when even
  out->keys_unsorted = (passes & 1) ? out->keys_tmp : out->keys;
  out->vals_unsorted = (passes & 1) ? out->vals_tmp : out->point_list;
  out->records = off;
  off = align128(off + n * (sh->C == 3 ? OCRF_RECORD_BYTES : 32));
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
  out->max_contrib = off; off = align128(off + tiles * 4);
  out->final_T = off;     off = align128(off + pix * 4);
  out->n_contrib = off;   off = align128(off + pix * 4);
  out->total = off + 128;
  return 0;
}

extern "C" int ocrf_bin_forward(void* stream, const OcrfShape* sh, uint64_t pair_capacity, const int32_t* radii,
                                const float* colors, int use_sh, void* geom_ws, void* bin_ws, void* image_ws) {
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
  cudaMemsetAsync(at<char>(image_ws, I.ranges), 0, tiles * 8, st);
  if (pair_capacity == 0) return 0;

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
  rc = sort_pairs_device(st, header + HDR_NUM_PAIRS, pair_capacity, end_bit, ka, va, kb, vb,
                         at<char>(bin_ws, B.histogram));
  if (rc) return rc;

  const unsigned grid = (unsigned)((pair_capacity + 255) / 256);
  if (sh->C == 3) {
    ranges_and_pack_kernel<false><<<grid, 256, 0, st>>>(
        *sh, pair_capacity, use_sh, header, at<uint64_t>(bin_ws, B.keys), at<uint32_t>(bin_ws, B.point_list),
        at<float>(geom_ws, G.depths), at<float2>(geom_ws, G.xy), at<float4>(geom_ws, G.conic_opacity),
        at<float>(geom_ws, G.rgb), colors, at<uint2>(image_ws, I.ranges), at<char>(bin_ws, B.records));
  } else {
    ranges_and_pack_kernel<true><<<grid, 256, 0, st>>>(
        *sh, pair_capacity, use_sh, header, at<uint64_t>(bin_ws, B.keys), at<uint32_t>(bin_ws, B.point_list),
        at<float>(geom_ws, G.depths), at<float2>(geom_ws, G.xy), at<float4>(geom_ws, G.conic_opacity),
        at<float>(geom_ws, G.rgb), colors, at<uint2>(image_ws, I.ranges), at<char>(bin_ws, B.records));
  }
  OCRF_CHECK_LAST();
  return 0;
}
