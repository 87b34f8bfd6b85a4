// Stage 2, fastest algorithm: ONE stable multi-split of the depth-ordered pair stream by tile.
//
// After the visible Gaussians of every view are sorted by (view | depth) (binning.cu), the sequence
//     S = for g in depth order: for tile in rect(g) (row-major): (g, tile)
// only has to be split STABLY by tile to obtain the reference's per-tile lists
// (cuda_rasterizer/rasterizer_impl.cu:70-138,303-318: a stable sort by tile of a depth-ordered sequence
// is the (tile | depth) order, ties in Gaussian-index order).  No pair is ever sorted:
//   (A) ms_count:   the stream is cut into chunks of equal PAIR count (so the screen-filling Gaussians
//                   that depth order packs together do not make a serial tail); every chunk enumerates
//                   its pairs, applies the exact tile cull (binning.cu: tile_can_contribute), counts
//                   pairs / kept pairs per tile and stores (tile, keep, g) per pair;
//   (B) two small scans: over the chunks of every tile, and over the tiles -> both range tables;
//   (C) ms_scatter: every chunk ranks its pairs per tile (warp match_any multi-split, the stable ranking
//                   of the onesweep sort) and writes the KEPT pairs as 48-byte blend records straight to
//                   their final slot, with their 1-based position in the tile's full list ("orig").
// The reference's key / point lists are not materialised on this path; the records and both range tables
// are bit-identical to what the pair-sort and depth-first paths produce from the same geometry state
// (tests/test_gpu_parity.py compares them), and `last_state(reference_lists=True)` materialises the
// reference lists on demand.
#include <cstdlib>

#include "common.cuh"

namespace ocrf {

constexpr int MS_THREADS = 256;
constexpr int MS_ITEMS = 16;
constexpr int MS_ROUND = MS_THREADS * MS_ITEMS;  // pairs per round of a chunk
constexpr int MS_WARPS = MS_THREADS / 32;
constexpr uint32_t MS_KEEP = 0x80000000u;

__device__ __forceinline__ bool ms_tile_can_contribute(float2 p, float4 co, float px0, float px1, float py0, float py1) {
  // identical arithmetic to binning.cu:tile_can_contribute (kept in sync by the record-equality tests)
  const float dxl = p.x - px1, dxh = p.x - px0, dyl = p.y - py1, dyh = p.y - py0;
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

__device__ __forceinline__ void ms_tile_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1,
                                             int& y1) {
  const float r = (float)radius;
  x0 = min(gx, max(0, (int)__fmul_rn(__fsub_rn(px, r), 0.0625f)));
  y0 = min(gy, max(0, (int)__fmul_rn(__fsub_rn(py, r), 0.0625f)));
  x1 = min(gx, max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(px, r), 16.f), 1.f), 0.0625f)));
  y1 = min(gy, max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(py, r), 16.f), 1.f), 0.0625f)));
}

// first / one-past-last pair of view v in the depth-ordered stream (sorted_offsets = inclusive scan)
__device__ __forceinline__ void ms_view_pairs(int v, const uint32_t* view_start, const uint32_t* sorted_offsets,
                                              uint32_t& pb, uint32_t& pe) {
  const uint32_t j0 = view_start[v], j1 = view_start[v + 1];
  pb = j0 ? sorted_offsets[j0 - 1] : 0u;
  pe = j1 ? sorted_offsets[j1 - 1] : 0u;
}

// (A) enumerate + cull + count + store items.  grid (chunks_max, V).
__global__ void __launch_bounds__(MS_THREADS) ms_count_kernel(
    OcrfShape sh, int chunks_max, uint32_t Q, uint64_t n_cap, uint32_t* __restrict__ header,
    const uint32_t* __restrict__ view_start, const uint32_t* __restrict__ sorted_offsets,
    const uint32_t* __restrict__ vis_vals, const int32_t* __restrict__ radii, const float2* __restrict__ xy,
    const float4* __restrict__ conic_opacity, uint32_t* __restrict__ cnt_full, uint32_t* __restrict__ cnt_kept,
    uint2* __restrict__ items, uint32_t* __restrict__ chain_words, int n_chain_words) {
  pdl_enter();
  extern __shared__ uint32_t s_dyn[];  // [2][tiles] counters, then [MS_ROUND/2] packed u16 Gaussian offsets
  __shared__ uint32_t s_jf, s_jl;
  __shared__ uint32_t s_wtot[MS_WARPS];
  const int v = blockIdx.y, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (c == 0 && v == 0)  // look-back state + ticket of the scan kernel that follows (no memset node in the chain)
    for (int i = tid; i < n_chain_words; i += MS_THREADS) chain_words[i] = 0u;
  if ((uint64_t)header[HDR_NUM_PAIRS] > n_cap) {
    if (c == 0 && v == 0 && tid == 0) atomicOr(&header[HDR_ERROR], ERR_PAIR_OVERFLOW);
    return;
  }
  uint32_t pb, pe;
  ms_view_pairs(v, view_start, sorted_offsets, pb, pe);
  const uint64_t cb64 = (uint64_t)pb + (uint64_t)c * Q;
  if (cb64 >= pe) return;
  const uint32_t cb = (uint32_t)cb64, ce = (uint32_t)min((uint64_t)pe, cb64 + Q);
  const int gx = ceil_div(sh.W, TILE), gy = ceil_div(sh.H, TILE), tiles = gx * gy;
  uint32_t* s_cnt = s_dyn;
  uint16_t* s_j = reinterpret_cast<uint16_t*>(s_dyn + 2 * tiles);
  for (int t = tid; t < 2 * tiles; t += MS_THREADS) s_cnt[t] = 0;
  const uint32_t jv0 = view_start[v], jv1 = view_start[v + 1];

  for (uint32_t rb = cb; rb < ce; rb += MS_ROUND) {
    const uint32_t n_items = min((uint32_t)MS_ROUND, ce - rb);
    __syncthreads();
    if (tid < 2) {  // Gaussian holding the first / last pair of the round: smallest j with sorted_offsets[j] > target
      const uint32_t target = tid == 0 ? rb : rb + n_items - 1;
      uint32_t lo = jv0, hi = jv1;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted_offsets[mid] > target) hi = mid; else lo = mid + 1;
      }
      if (tid == 0) s_jf = lo; else s_jl = lo;
    }
    for (int i = tid; i < MS_ROUND / 2; i += MS_THREADS) reinterpret_cast<uint32_t*>(s_j)[i] = 0;
    __syncthreads();
    const uint32_t jf = s_jf, jl = s_jl;
    // flag the first pair of every Gaussian that starts inside the round, then scan -> Gaussian offset per pair
    for (uint32_t j = jf + 1 + tid; j <= jl; j += MS_THREADS) s_j[sorted_offsets[j - 1] - rb] = 1;
    __syncthreads();
    {
      uint32_t* mine = reinterpret_cast<uint32_t*>(s_j) + tid * (MS_ITEMS / 2);
      uint32_t w[MS_ITEMS / 2];
      uint32_t sum = 0;
#pragma unroll
      for (int i = 0; i < MS_ITEMS / 2; i++) {
        w[i] = mine[i];
        sum += (w[i] & 0xffffu) + (w[i] >> 16);
      }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
      }
      if (lane == 31) s_wtot[warp] = incl;
      __syncthreads();
      uint32_t run = incl - sum;
#pragma unroll
      for (int q = 0; q < MS_WARPS; q++)
        if (q < warp) run += s_wtot[q];
#pragma unroll
      for (int i = 0; i < MS_ITEMS / 2; i++) {
        const uint32_t a = run + (w[i] & 0xffffu);
        const uint32_t b = a + (w[i] >> 16);
        mine[i] = a | (b << 16);
        run = b;
      }
    }
    __syncthreads();
    // every pair: its Gaussian, its tile, the exact cull
#pragma unroll 4
    for (int i = 0; i < MS_ITEMS; i++) {
      const uint32_t pos = warp * (32 * MS_ITEMS) + i * 32 + lane;
      if (pos >= n_items) continue;
      const uint32_t j = jf + s_j[pos];
      const uint32_t excl = j ? sorted_offsets[j - 1] : 0u;
      const int k = (int)(rb + pos - excl);
      const uint32_t g = vis_vals[j];
      const float2 p = xy[g];
      int x0, y0, x1, y1;
      ms_tile_rect(p.x, p.y, radii[g], gx, gy, x0, y0, x1, y1);
      const int w = x1 - x0;
      const int ty = y0 + k / w, tx = x0 + k - (k / w) * w;
      const int t = ty * gx + tx;
      const float px0 = (float)(tx * TILE), px1 = (float)min(tx * TILE + TILE - 1, sh.W - 1);
      const float py0 = (float)(ty * TILE), py1 = (float)min(ty * TILE + TILE - 1, sh.H - 1);
      const bool keep = ms_tile_can_contribute(p, conic_opacity[g], px0, px1, py0, py1);
      atomicAdd(&s_cnt[t], 1u);
      if (keep) atomicAdd(&s_cnt[tiles + t], 1u);
      items[rb + pos] = make_uint2((uint32_t)t | (keep ? MS_KEEP : 0u), g);
    }
  }
  __syncthreads();
  const size_t row = ((size_t)v * chunks_max + c) * tiles;
  for (int t = tid; t < tiles; t += MS_THREADS) {
    cnt_full[row + t] = s_cnt[t];
    cnt_kept[row + t] = s_cnt[tiles + t];
  }
}

// (B) exclusive scan over the chunks of every tile (in place, both tables) AND the exclusive scan of the tile totals
// over the whole batch -> tile offsets and both range tables, in ONE kernel: a CTA owns eight consecutive tiles of a
// view (one warp per tile: the lanes take 32 consecutive chunks, so a tile's column is read in ceil(chunks/32) parallel
// round trips), and the CTAs are chained by a decoupled look-back over their eight-tile totals (ticketed, so a CTA only
// ever waits for CTAs that already run).  Round 1 ran the tile scan as a separate single-CTA kernel (11 us of launch
// and latency for 4 224 numbers).  The look-back words live in `chain` (zeroed by ms_count_kernel).
__global__ void __launch_bounds__(256) ms_scan_kernel(OcrfShape sh, int chunks_max, uint32_t Q, uint64_t n_cap,
                                                      const uint32_t* __restrict__ header,
                                                      const uint32_t* __restrict__ view_start,
                                                      const uint32_t* __restrict__ sorted_offsets,
                                                      uint32_t* __restrict__ cnt_full, uint32_t* __restrict__ cnt_kept,
                                                      unsigned long long* __restrict__ chain, uint32_t* __restrict__ ticket,
                                                      uint32_t* __restrict__ tile_offset, uint2* __restrict__ ranges,
                                                      uint2* __restrict__ ranges_render, uint32_t* __restrict__ sticky) {
  pdl_enter();
  __shared__ uint32_t s_bid, s_prefix;
  __shared__ uint32_t s_full[8], s_kept[8];
  const int tiles = ceil_div(sh.W, TILE) * ceil_div(sh.H, TILE);
  const int groups = ceil_div(tiles, 8);  // CTAs per view
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
  __syncthreads();
  const int bid = (int)s_bid;
  const int v = bid / groups;
  const int t = (bid - v * groups) * 8 + warp;
  const bool overflow = (uint64_t)header[HDR_NUM_PAIRS] > n_cap;
  if (sticky != nullptr && bid == 0 && threadIdx.x == 0) publish_status(header, n_cap, sticky);
  uint32_t rf = 0, rk = 0;
  // (capacity overflow: ms_count_kernel left the tables unwritten and the real pair count is not bounded by the table
  // extent -- touch nothing, every range reads (0, 0) and only the background is rendered)
  if (t < tiles && !overflow) {
    uint32_t pb, pe;
    ms_view_pairs(v, view_start, sorted_offsets, pb, pe);
    const int nchunks = min(chunks_max, (int)(((uint64_t)(pe - pb) + Q - 1) / Q));
    uint32_t* cf = cnt_full + (size_t)v * chunks_max * tiles + t;
    uint32_t* ck = cnt_kept + (size_t)v * chunks_max * tiles + t;
    constexpr int BATCH = 8;  // rounds of 32 chunks whose loads are all issued before the first in-place store
    for (int cb = 0; cb < nchunks; cb += 32 * BATCH) {
      uint32_t xf[BATCH], xk[BATCH];
#pragma unroll
      for (int r = 0; r < BATCH; r++) {
        const int c = cb + r * 32 + lane;
        xf[r] = c < nchunks ? cf[(size_t)c * tiles] : 0u;
        xk[r] = c < nchunks ? ck[(size_t)c * tiles] : 0u;
      }
#pragma unroll
      for (int r = 0; r < BATCH; r++) {
        const int c = cb + r * 32 + lane;
        if (cb + r * 32 >= nchunks) break;  // warp-uniform
        uint32_t inf = xf[r], ink = xk[r];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t yf = __shfl_up_sync(0xffffffffu, inf, d), yk = __shfl_up_sync(0xffffffffu, ink, d);
          if (lane >= d) { inf += yf; ink += yk; }
        }
        if (c < nchunks) {
          cf[(size_t)c * tiles] = rf + inf - xf[r];
          ck[(size_t)c * tiles] = rk + ink - xk[r];
        }
        rf += __shfl_sync(0xffffffffu, inf, 31);
        rk += __shfl_sync(0xffffffffu, ink, 31);
      }
    }
  }
  if (lane == 0) { s_full[warp] = rf; s_kept[warp] = rk; }
  __syncthreads();
  if (warp == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) total += s_full[w];
    const unsigned long long excl = lookback_warp(chain, bid, total);
    if (lane == 0) s_prefix = (uint32_t)excl;
  }
  __syncthreads();
  if (lane == 0 && t < tiles) {
    uint32_t run = s_prefix;
    for (int w = 0; w < warp; w++) run += s_full[w];
    const size_t vt = (size_t)v * tiles + t;
    tile_offset[vt] = run;
    ranges[vt] = rf ? make_uint2(run, run + rf) : make_uint2(0u, 0u);  // empty tiles read (0,0) like the reference
    ranges_render[vt] = overflow ? make_uint2(0u, 0u) : make_uint2(run, run + rk);
  }
}

// (C) stable ranks per tile + direct record writes.  grid (chunks_max, V).
__global__ void __launch_bounds__(MS_THREADS) ms_scatter_kernel(
    OcrfShape sh, int chunks_max, uint32_t Q, uint64_t n_cap, int use_sh, int has_rgb,
    const uint32_t* __restrict__ header, const uint32_t* __restrict__ view_start,
    const uint32_t* __restrict__ sorted_offsets, const uint2* __restrict__ items, const float2* __restrict__ xy,
    const float4* __restrict__ conic_opacity, const float* __restrict__ depths, const float* __restrict__ rgb,
    const float* __restrict__ colors, const uint32_t* __restrict__ cnt_full, const uint32_t* __restrict__ cnt_kept,
    const uint32_t* __restrict__ tile_offset, Record* __restrict__ records) {
  pdl_enter();
  extern __shared__ uint32_t s_dyn[];  // [2][tiles] running bases (full: relative to the tile list, kept: record slot)
                                       // then [MS_WARPS][2][tiles] u16 per-warp counts / prefixes of the round, [2][tiles] u16 round totals
  const int v = blockIdx.y, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((uint64_t)header[HDR_NUM_PAIRS] > n_cap) return;
  uint32_t pb, pe;
  ms_view_pairs(v, view_start, sorted_offsets, pb, pe);
  const uint64_t cb64 = (uint64_t)pb + (uint64_t)c * Q;
  if (cb64 >= pe) return;
  const uint32_t cb = (uint32_t)cb64, ce = (uint32_t)min((uint64_t)pe, cb64 + Q);
  const int tiles = ceil_div(sh.W, TILE) * ceil_div(sh.H, TILE);
  uint32_t* s_base = s_dyn;
  // per-warp counts / prefixes of the round, ONE word per (warp, tile): low half = pairs, high half = kept pairs
  // (both are updated by the same lane with one load and one store)
  uint32_t* s_wc = s_dyn + 2 * tiles;                                    // [MS_WARPS][tiles]
  uint16_t* s_rt = reinterpret_cast<uint16_t*>(s_wc + (size_t)MS_WARPS * tiles);  // [2][tiles] totals of the current round
  uint32_t* my_cnt = s_wc + (size_t)warp * tiles;
  const size_t row = ((size_t)v * chunks_max + c) * tiles;
  const uint32_t* toff = tile_offset + (size_t)v * tiles;
  for (int t = tid; t < tiles; t += MS_THREADS) {
    s_base[t] = cnt_full[row + t];
    s_base[tiles + t] = toff[t] + cnt_kept[row + t];
  }
  const uint32_t lt_mask = (1u << lane) - 1;
  const size_t sample_base = (size_t)(v / sh.views_per_sample) * sh.P;
  const uint32_t gv0 = (uint32_t)v * (uint32_t)sh.P;

  for (uint32_t rb = cb; rb < ce; rb += MS_ROUND) {
    const uint32_t n_items = min((uint32_t)MS_ROUND, ce - rb);
    __syncthreads();
    for (int i = tid; i < MS_WARPS * tiles; i += MS_THREADS) s_wc[i] = 0;
    __syncthreads();
    // rank inside the warp, one row of 32 consecutive pairs at a time (stable)
    uint32_t st[MS_ITEMS];  // tile | rank_full << 13 | rank_kept << 22 | keep << 31   (ranks < 512 per warp-round)
    uint32_t gi[MS_ITEMS];
    // groups of four rows: the four (independent) peer matches are issued together, then the four dependent
    // load-add-store steps on the warp's counters follow
#pragma unroll
    for (int i0 = 0; i0 < MS_ITEMS; i0 += 4) {
      uint32_t tt[4], peers[4], kpeers[4];
      bool vld[4], kp[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t pos = warp * (32 * MS_ITEMS) + (i0 + u) * 32 + lane;
        vld[u] = pos < n_items;
        uint2 it = make_uint2(0xffffu, 0u);
        if (vld[u]) it = items[rb + pos];
        tt[u] = it.x & 0x1fffu;
        kp[u] = vld[u] && (it.x & MS_KEEP);
        gi[i0 + u] = it.y;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        peers[u] = __match_any_sync(0xffffffffu, vld[u] ? tt[u] : 0xffff0000u | lane);
        // the kept pairs of my tile among my peers: no second match, the keep flags of the warp are one vote away
        kpeers[u] = peers[u] & __ballot_sync(0xffffffffu, kp[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        uint32_t b4 = 0;
        if (vld[u]) b4 = my_cnt[tt[u]];
        __syncwarp();
        if (vld[u] && (peers[u] & lt_mask) == 0) my_cnt[tt[u]] = b4 + __popc(peers[u]) + (__popc(kpeers[u]) << 16);
        __syncwarp();
        const uint32_t rf = (b4 & 0xffffu) + __popc(peers[u] & lt_mask), rk = (b4 >> 16) + __popc(kpeers[u] & lt_mask);
        st[i0 + u] = vld[u] ? (tt[u] | (rf << 13) | (rk << 22) | (kp[u] ? MS_KEEP : 0u)) : 0xffffffffu;
      }
    }
    __syncthreads();
    // per tile: exclusive prefix over the warps (stored back into the per-warp slots) and the round totals
    for (int t = tid; t < tiles; t += MS_THREADS) {
      uint32_t rf = 0, rk = 0;
#pragma unroll
      for (int w = 0; w < MS_WARPS; w++) {
        uint32_t* f = s_wc + (size_t)w * tiles;
        const uint32_t x = f[t];
        f[t] = rf | (rk << 16);
        rf += x & 0xffffu;
        rk += x >> 16;
      }
      s_rt[t] = (uint16_t)rf;
      s_rt[tiles + t] = (uint16_t)rk;
    }
    __syncthreads();
    // write the kept pairs of this round
#pragma unroll 4
    for (int i = 0; i < MS_ITEMS; i++) {
      const uint32_t s = st[i];
      if (s == 0xffffffffu || !(s & MS_KEEP)) continue;
      const uint32_t t = s & 0x1fffu, rf = (s >> 13) & 0x1ffu, rk = (s >> 22) & 0x1ffu;
      const uint32_t wpre = my_cnt[t];                                   // this warp's prefix inside the round
      const uint32_t orig = s_base[t] + (wpre & 0xffffu) + rf + 1;        // 1-based position in the tile's full list
      const uint32_t dst = s_base[tiles + t] + (wpre >> 16) + rk;         // record slot
      const uint32_t g = gi[i];
      const float2 p = xy[g];
      const float4 co = conic_opacity[g];
      const uint32_t id = g - gv0;
      float r = 0.f, gg = 0.f, bb = 0.f;
      if (has_rgb) {
        const float* col = use_sh ? rgb + (size_t)g * 3 : colors + (sample_base + id) * 3;
        r = __ldg(col); gg = __ldg(col + 1); bb = __ldg(col + 2);
      }
      float4* out = reinterpret_cast<float4*>(records + dst);
      out[0] = record_head(p, co);
      out[1] = make_float4(record_qc(co), co.w, __uint_as_float(orig), r);
      out[2] = make_float4(gg, bb, __uint_as_float(id), depths[g]);
    }
    __syncthreads();
    for (int t = tid; t < tiles; t += MS_THREADS) {  // advance the running bases past this round
      s_base[t] += s_rt[t];
      s_base[tiles + t] += s_rt[tiles + t];
    }
  }
}

}  // namespace ocrf

using namespace ocrf;

namespace ocrf {
// Host entry used by ocrf_bin_forward (binning.cu).  Returns 0, a cudaError_t, or OCRF_E*.
int multisplit_bin(cudaStream_t st, const OcrfShape* sh, uint64_t pair_capacity, int use_sh, const int32_t* radii,
                   const float* colors, uint32_t* header, const uint32_t* view_start, const uint32_t* sorted_offsets,
                   const uint32_t* vis_vals, const float2* xy, const float4* conic_opacity, const float* depths,
                   const float* rgb, uint32_t* tables, size_t table_words, uint32_t* tile_arrays, uint2* items,
                   uint2* ranges, uint2* ranges_render, Record* records, uint32_t* sticky) {
  const int tiles_v = tiles_x(*sh) * tiles_y(*sh);
  if (tiles_v > 8191) return OCRF_ECAPACITY;
  // chunk size: multiples of MS_ROUND pairs, at most ~2048 chunks for the whole capacity
  static_assert(MS_ROUND == 4096, "multisplit_chunk_pairs (common.cuh) sizes the chunk tables in rounds of 4096 pairs");
  const uint32_t Q = multisplit_chunk_pairs(pair_capacity);
  const int chunks_max = (int)((pair_capacity + Q - 1) / Q) + 1;
  if ((size_t)2 * sh->V * chunks_max * tiles_v > table_words) return OCRF_ECAPACITY;
  uint32_t* cnt_full = tables;
  uint32_t* cnt_kept = tables + (size_t)sh->V * chunks_max * tiles_v;
  // tile arrays: [0, V*tiles) the look-back words of the scan (uint64 per CTA = per eight tiles, then its ticket),
  // [2*V*tiles, 3*V*tiles) the tile offsets
  const int scan_blocks = sh->V * ceil_div(tiles_v, 8);
  unsigned long long* chain = reinterpret_cast<unsigned long long*>(tile_arrays);
  uint32_t* scan_ticket = tile_arrays + 2 * (size_t)scan_blocks + 2;
  const int n_chain_words = 2 * scan_blocks + 4;
  uint32_t* tile_offset = tile_arrays + 2 * (size_t)sh->V * tiles_v;
  const size_t smem_a = (size_t)2 * tiles_v * 4 + MS_ROUND * 2;
  const size_t smem_c = (size_t)2 * tiles_v * 4 + (size_t)MS_WARPS * tiles_v * 4 + (size_t)2 * tiles_v * 2;
  static unsigned long long attr_a = 0, attr_c = 0;  // per-device bit masks
  cudaError_t ae = ensure_dynamic_smem(ms_count_kernel, 100 * 1024, attr_a);
  if (ae != cudaSuccess) return (int)ae;
  ae = ensure_dynamic_smem(ms_scatter_kernel, 200 * 1024, attr_c);
  if (ae != cudaSuccess) return (int)ae;
  if (smem_a > 100 * 1024 || smem_c > 200 * 1024) return OCRF_ECAPACITY;
  const dim3 grid(chunks_max, sh->V);
#ifdef OCRF_DIAG  // timing experiments only: repeat an idempotent kernel, the stage-time delta is its in-stream cost
  static const int diag_dup = getenv("OCRF_DUP") ? atoi(getenv("OCRF_DUP")) : 0;
#define OCRF_DIAG_REP(bit) (((diag_dup >> (bit)) & 1) ? 2 : 1)
#else
#define OCRF_DIAG_REP(bit) 1
#endif
  for (int rep = 0; rep < OCRF_DIAG_REP(0); rep++)
  OCRF_LAUNCH(ms_count_kernel, dim3(grid), dim3(MS_THREADS), smem_a, st, *sh, chunks_max, Q, pair_capacity, header, view_start, sorted_offsets,
                                                    vis_vals, radii, xy, conic_opacity, cnt_full, cnt_kept, items,
                                                    tile_arrays, n_chain_words);
  OCRF_LAUNCH(ms_scan_kernel, dim3(scan_blocks), dim3(256), 0, st, *sh, chunks_max, Q, pair_capacity, header, view_start,
              sorted_offsets, cnt_full, cnt_kept, chain, scan_ticket, tile_offset, ranges, ranges_render, sticky);
  for (int rep = 0; rep < OCRF_DIAG_REP(2); rep++)
  OCRF_LAUNCH(ms_scatter_kernel, dim3(grid), dim3(MS_THREADS), smem_c, st, *sh, chunks_max, Q, pair_capacity, use_sh, sh->C == 3, header,
                                                      view_start, sorted_offsets, items, xy, conic_opacity, depths, rgb,
                                                      colors, cnt_full, cnt_kept, tile_offset, records);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
}  // namespace ocrf
