// Stage 2, first step: depth-sort the VISIBLE Gaussians of every view, on chip, and scan their tile counts.
//
// The reference sorts every (tile | depth) pair (cuda_rasterizer/rasterizer_impl.cu:303-308).  Here only the visible
// Gaussians of a view (~19 k at the headline shape, 228 KB of (key, value) pairs) are sorted by depth; the pair stream
// is then split stably by tile (multisplit.cu).  Round 1 pushed these 114 k keys through the global onesweep sort: a
// histogram kernel and five chained passes, 77 us for 1.4 MB -- a fixed chain of launches and L2 round trips.
//
// Two kernels, both with one thread-block CLUSTER of 8 CTAs per view (the views are already segments of the compacted
// key array, so the view bits need no pass) and four 8-bit passes over the depth bits:
//   visible_sort_reg_kernel  (below, second in this file) holds the keys in REGISTERS and exchanges them through
//                            distributed shared memory: up to 131 072 visible Gaussians per view -- the one that runs;
//   visible_sort_kernel      (first in this file, the round's first version) keeps them in an L2-resident ping-pong:
//                            any segment size; takes the views above that bound, or all of them with OCRF_VIS_SORT_REG=0.
// Both end with the same epilogue (see below) and write the same bits.
//
// visible_sort_kernel: per 8-bit digit place, CTA r of the cluster
//   (1) reads the eight digit histograms of the pass through distributed shared memory and derives, per digit, the
//       global base + the keys of lower-ranked CTAs: the stable destination of its own keys,
//   (2) ranks its eighth of the segment (warp match_any multi-split, the ranking of radix_sort.cu; 1024 threads x 4
//       keys, so the serial part of the ranking is four steps) and writes keys and values to the other half of the
//       ping-pong (L2-resident),
//   (3) and, with the same scatter, counts the NEXT pass's digit of every key it writes into the histogram of the CTA
//       that will own the destination (a distributed-shared-memory reduction),
// so ONE cluster barrier per pass orders both the data and the next histograms.  Four passes cover the 32 depth bits;
// no look-back chain, no global histogram, one launch.  The epilogue replaces scan_sorted_tiles_kernel: the inclusive
// scan of tiles_touched in sorted order (where every Gaussian's pairs sit in the pair stream), the cluster exchanging
// eight partial sums.  Any segment size works (a CTA loops over sub-tiles of 4096 keys).
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ocrf {

constexpr int VS_CLUSTER = 8;
constexpr int VS_THREADS = 1024;
constexpr int VS_WARPS = VS_THREADS / 32;
constexpr int VS_ITEMS = 4;
constexpr int VS_SUBTILE = VS_THREADS * VS_ITEMS;

// exclusive scan of one value per thread over the first 256 threads (8 warps); every thread of the CTA calls
__device__ __forceinline__ uint32_t vs_exclusive_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  __syncthreads();  // s_warp may still be read from the previous use
  if (lane == 31 && warp < 8) s_warp[warp] = incl;
  __syncthreads();
  uint32_t off = 0;
#pragma unroll
  for (int w = 0; w < 8; w++)
    if (w < warp) off += s_warp[w];
  return off + incl - v;
}

__device__ __forceinline__ void dsmem_add_u32(uint32_t* remote, uint32_t v) {
  // remote: a generic address returned by cluster.map_shared_rank (shared::cluster window)
  atomicAdd(remote, v);
}

__global__ void __cluster_dims__(VS_CLUSTER, 1, 1) __launch_bounds__(VS_THREADS) visible_sort_kernel(
    int P, const uint32_t* __restrict__ view_start, uint64_t* keys0, uint32_t* vals0, uint64_t* keys1, uint32_t* vals1,
    const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ offsets,
    uint32_t* __restrict__ sorted_offsets, uint32_t skip_upto) {
  pdl_enter();
  // (views of at most skip_upto keys were sorted by visible_sort_cta_kernel; the whole cluster sees the same count)
  if (view_start[blockIdx.y + 1] - view_start[blockIdx.y] <= skip_upto) return;
  cg::cluster_group cluster = cg::this_cluster();
  // digit counts of MY part of the segment, three buffers in rotation: pass p reads [p % 3] (all CTAs, remotely), fills
  // [(p + 1) % 3] for the next pass (remote reductions of whoever writes into my part) and clears [(p + 2) % 3], which
  // was last read before the previous cluster barrier and is next written after the coming one
  __shared__ uint32_t s_hist[3][256];
  __shared__ uint32_t s_base[256];                 // running destination of my next key of every digit
  __shared__ uint16_t s_warp_hist[VS_WARPS][256];  // per-warp digit counts of the current sub-tile (<= 128 each)
  __shared__ uint32_t s_scan[VS_WARPS];
  __shared__ uint32_t s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster.block_rank();
  const int v = blockIdx.y;
  const uint32_t seg_b = view_start[v], n = view_start[v + 1] - seg_b;
  const uint32_t per = max(1u, (n + VS_CLUSTER - 1) / VS_CLUSTER);
  const uint32_t cb = seg_b + min(n, rank * per), ce = seg_b + min(n, (rank + 1) * per);
  const uint32_t lt_mask = (1u << lane) - 1;

  // histogram of the first digit place over my part
  if (tid < 256) { s_hist[0][tid] = 0; s_hist[1][tid] = 0; s_hist[2][tid] = 0; }
  __syncthreads();
  for (uint32_t base = cb; base < ce; base += VS_THREADS) {  // (uniform trip count: the match below is warp-wide)
    const uint32_t i = base + tid;
    const bool valid = i < ce;
    const uint32_t d = valid ? ((uint32_t)keys0[i] & 255u) : 0x100u + lane;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (valid && (peers & lt_mask) == 0) atomicAdd(&s_hist[0][d], __popc(peers));
  }
  __syncthreads();
  cluster.sync();

  for (int pass = 0; pass < 4; pass++) {
    const uint64_t* src_k = (pass & 1) ? keys1 : keys0;
    const uint32_t* src_v = (pass & 1) ? vals1 : vals0;
    uint64_t* dst_k = (pass & 1) ? keys0 : keys1;
    uint32_t* dst_v = (pass & 1) ? vals0 : vals1;
    const int shift = 8 * pass;
    uint32_t* hist_now = s_hist[pass % 3];
    uint32_t* hist_next = s_hist[(pass + 1) % 3];
    if (tid < 256) s_hist[(pass + 2) % 3][tid] = 0;
    // (1) thread d < 256: keys with digit d in the whole segment, and in the parts before mine
    uint32_t total = 0, before = 0;
    if (tid < 256) {
#pragma unroll
      for (uint32_t r = 0; r < VS_CLUSTER; r++) {
        const uint32_t c = *cluster.map_shared_rank(&hist_now[tid], r);
        total += c;
        before += r < rank ? c : 0u;
      }
    }
    const uint32_t excl = vs_exclusive_scan_256(total, s_scan);
    if (tid < 256) s_base[tid] = seg_b + excl + before;
    // (2) stable ranks + scatter, one sub-tile of 4096 keys at a time
    for (uint32_t sub = cb; sub < ce; sub += VS_SUBTILE) {
      const uint32_t n_valid = min((uint32_t)VS_SUBTILE, ce - sub);
      for (int e = tid; e < VS_WARPS * 128; e += VS_THREADS) reinterpret_cast<uint32_t*>(s_warp_hist)[e] = 0u;
      uint64_t key[VS_ITEMS];
      uint32_t val[VS_ITEMS], peers[VS_ITEMS], rk[VS_ITEMS];
      uint16_t* my_hist = s_warp_hist[warp];
      const uint32_t warp_base = warp * (32 * VS_ITEMS);
#pragma unroll
      for (int i = 0; i < VS_ITEMS; i++) {
        const uint32_t loc = warp_base + i * 32 + lane;  // index order inside the warp: the ranking below is stable
        const bool valid = loc < n_valid;
        key[i] = valid ? src_k[sub + loc] : 0ull;
        val[i] = valid ? src_v[sub + loc] : 0u;
      }
#pragma unroll
      for (int i = 0; i < VS_ITEMS; i++) {  // the peer masks are independent of each other: issue them together
        const bool valid = warp_base + i * 32 + lane < n_valid;
        peers[i] = __match_any_sync(0xffffffffu, valid ? ((uint32_t)(key[i] >> shift) & 255u) : 0x100u + lane);
      }
      __syncthreads();  // s_warp_hist is zero; s_base of the previous sub-tile / of step (1) is in place
#pragma unroll
      for (int i = 0; i < VS_ITEMS; i++) {
        const bool valid = warp_base + i * 32 + lane < n_valid;
        const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        uint32_t b4 = 0;
        if (valid) b4 = my_hist[d];
        __syncwarp();
        if (valid && (peers[i] & lt_mask) == 0) my_hist[d] = (uint16_t)(b4 + __popc(peers[i]));
        __syncwarp();
        rk[i] = b4 + __popc(peers[i] & lt_mask);
      }
      __syncthreads();
      uint32_t tile_count = 0;
      if (tid < 256) {
#pragma unroll 8
        for (int w = 0; w < VS_WARPS; w++) {  // thread d: exclusive prefix over the warps
          const uint32_t c = s_warp_hist[w][tid];
          s_warp_hist[w][tid] = (uint16_t)tile_count;
          tile_count += c;
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < VS_ITEMS; i++) {
        const uint32_t loc = warp_base + i * 32 + lane;
        const bool valid = loc < n_valid;
        uint32_t slot = 0xffff0000u + lane;  // (owner, next digit) of my key; unique for an invalid lane
        if (valid) {
          const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
          const uint32_t pos = s_base[d] + my_hist[d] + rk[i];
          dst_k[pos] = key[i];
          dst_v[pos] = val[i];
          const uint32_t owner = min((pos - seg_b) / per, (uint32_t)VS_CLUSTER - 1);
          slot = (owner << 8) | ((uint32_t)(key[i] >> (shift + 8)) & 255u);
        }
        if (pass < 3) {
          // (3) the next pass's digit of this key, counted where the key now lives.  The high bytes of a depth take a
          // handful of values, so unaggregated every key of the view would hit the same few remote counters: the
          // lanes of a warp that share (owner, digit) send ONE remote reduction
          const uint32_t same = __match_any_sync(0xffffffffu, slot);
          if (valid && (same & lt_mask) == 0)
            dsmem_add_u32(cluster.map_shared_rank(&hist_next[slot & 255u], slot >> 8), (uint32_t)__popc(same));
        }
      }
      __syncthreads();
      if (tid < 256) s_base[tid] += tile_count;
    }
    __threadfence();
    cluster.sync();  // the pass and the next histograms are complete and visible to the whole cluster
  }

  // ---- epilogue: inclusive scan of tiles_touched over the sorted order (four passes end in keys0 / vals0) ----
  uint32_t mine = 0;
  for (uint32_t i = cb + tid; i < ce; i += VS_THREADS) {
    const uint32_t x = tiles_touched[vals0[i]];
    sorted_offsets[i] = x;  // parked; rewritten below
    mine += x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0) s_scan[warp] = mine;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < VS_WARPS; w++) t += s_scan[w];
    s_total = t;
  }
  __syncthreads();
  cluster.sync();
  uint32_t run = v ? offsets[(size_t)v * P - 1] : 0u;  // pairs of the views before mine (preprocess's scan)
  for (uint32_t r = 0; r < rank; r++) run += *cluster.map_shared_rank(&s_total, r);
  for (uint32_t sub = cb; sub < ce; sub += VS_THREADS) {  // one key per thread, a CTA-wide scan per 1024 keys
    const uint32_t j = sub + tid;
    const uint32_t x = j < ce ? sorted_offsets[j] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll 8
    for (int w = 0; w < VS_WARPS; w++) {
      const uint32_t c = s_scan[w];
      off += w < warp ? c : 0u;
      tot += c;
    }
    if (j < ce) sorted_offsets[j] = run + off + incl;
    run += tot;
  }
  cluster.sync();  // no CTA leaves while a peer may still read its shared memory
}

// ---- the same sort with the keys in REGISTERS: the kernel that runs for up to 131 072 visible Gaussians per view ----
// What the kernel above pays per pass (measured: ~12 us at 17 % issue-active for 2.4 k keys per CTA) is not arithmetic:
// every key sends a remote shared-memory reduction for the next histogram, the scatter goes through L2, a
// __threadfence and a cluster barrier close the pass, and the next pass starts with an L2 round trip for its keys.
// Here a cluster of 8 CTAs holds the view's keys as (depth, position in the segment) in registers -- R = ceil(n / 8192)
// rows of 32 per warp -- and a pass is: rank inside the warp, per-CTA digit totals, ONE cluster barrier, read the eight
// totals through distributed shared memory, scatter every key with one 8-byte remote shared-memory store to the CTA
// that owns its destination slot, a second cluster barrier, read the own slots back.  Nothing touches global memory
// between the first load and the result; the Gaussian indices are gathered once, after the last pass.
//
// Ranking: tools/probe/match_probe.cu measured MATCH.ANY on this part at ~2 cycles per DISTINCT value per SM (64 cycles
// for 32 distinct digits, 4 for one), against 31 cycles flat for eight ballots.  The three low digit places of a depth
// are noise (32 distinct digits per warp): eight ballots; the top byte (sign + 7 exponent bits) takes a handful of
// values: match.  (A first version with ONE CTA per view and 20 rows per warp was correct and 2x SLOWER than the
// kernel above -- 115 vs 60 us: 768 matches per pass on one SM's match unit.)
constexpr int VR_CLUSTER = 8;
constexpr int VR_THREADS = 1024;
constexpr int VR_WARPS = VR_THREADS / 32;
constexpr int VR_ITEMS = 16;
constexpr uint32_t VR_CAP = (uint32_t)VR_CLUSTER * VR_THREADS * VR_ITEMS;  // 131 072 keys per view
constexpr uint32_t VR_IDX_MASK = 0x3ffffu;                                  // position in the segment (< 2^17), rank above

struct alignas(16) VrFixed {  // fixed part of the dynamic shared memory; the (depth, position) slots of the CTA follow
  uint16_t warp_hist[VR_WARPS][256];  // per-warp digit counts (<= 512), then the exclusive prefix over the warps
  uint32_t cta_cnt[256];              // digit counts of this CTA's slots (read by the whole cluster)
  uint32_t base[256];                 // first destination of this CTA's keys of every digit
  uint32_t scan[VR_WARPS];
  uint32_t total;
};

#ifdef OCRF_VS_STAMPS  // tools/probe/vsort_probe.cu: clock64 of thread 0 of every CTA at the phases of the kernel
__device__ long long* g_vs_stamps = nullptr;
#define VR_STAMP(id)                                                                                          \
  do {                                                                                                        \
    if (threadIdx.x == 0 && g_vs_stamps) g_vs_stamps[(blockIdx.y * VR_CLUSTER + blockIdx.x) * 64 + (id)] = clock64(); \
  } while (0)
#else
#define VR_STAMP(id) do { } while (0)
#endif

__device__ __forceinline__ uint32_t vr_peers(uint32_t d, bool few_values) {  // lanes of the warp holding my digit
  if (few_values) return __match_any_sync(0xffffffffu, d);
  uint32_t peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; b++) {
    const bool bit = (d >> b) & 1u;
    const uint32_t vote = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? vote : ~vote;
  }
  return peers;
}

__global__ void __cluster_dims__(VR_CLUSTER, 1, 1) __launch_bounds__(VR_THREADS, 1) visible_sort_reg_kernel(
    int P, const uint32_t* __restrict__ view_start, uint64_t* keys0, uint32_t* vals0,
    const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ offsets,
    uint32_t* __restrict__ sorted_offsets) {
  pdl_enter();
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char vr_raw[];
  VrFixed& sm = *reinterpret_cast<VrFixed*>(vr_raw);
  uint2* buf = reinterpret_cast<uint2*>(vr_raw + sizeof(VrFixed));  // [1024 R] slots of this CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster.block_rank();
  const int v = blockIdx.y;
  const uint32_t seg_b = view_start[v], n = view_start[v + 1] - seg_b;
  if (n == 0 || n > VR_CAP) return;  // (the whole cluster; larger views are sorted by visible_sort_kernel)
  // R rows of 32 consecutive slots per warp, 1024 R slots per CTA, CTA-major: slots past n hold the largest key
  const uint32_t R = (n + VR_CLUSTER * VR_THREADS - 1) / (VR_CLUSTER * VR_THREADS);
  const uint32_t per = (uint32_t)VR_THREADS * R;
  const uint32_t inv_R = (65536u + R - 1) / R;  // (x * inv_R) >> 16 == x / R for x < 8 R <= 128
  const uint32_t local_base = (uint32_t)warp * 32u * R;
  const uint32_t glob_base = rank * per + local_base;
  const uint32_t lt_mask = (1u << lane) - 1;
  uint16_t* my_hist = sm.warp_hist[warp];
  VR_STAMP(0);
#pragma unroll
  for (int e = 0; e < 4; e++) reinterpret_cast<uint32_t*>(my_hist)[e * 32 + lane] = 0u;
  __syncwarp();

  uint32_t k[VR_ITEMS], m[VR_ITEMS];  // depth bits; position in the segment | rank of the pass << 18
  {
    const uint32_t* lo = reinterpret_cast<const uint32_t*>(keys0 + seg_b);  // low word of (view << 32 | depth)
#pragma unroll
    for (int i = 0; i < VR_ITEMS; i++) {
      k[i] = 0xffffffffu;
      m[i] = 0u;
      if (i < R) {
        const uint32_t loc = glob_base + i * 32 + lane;
        if (loc < n) k[i] = lo[2 * (size_t)loc];
        m[i] = loc;
      }
    }
  }

#pragma unroll 1
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 8 * pass;
    const bool few_values = pass == 3;
    VR_STAMP(1 + 8 * pass);  // keys in registers
    // stable rank inside the warp, row by row (the padding keys are ordinary keys: they sort behind everything);
    // a warp's counters are zeroed by the warp itself (below): no CTA barrier in front of the ranking
#pragma unroll
    for (int i = 0; i < VR_ITEMS; i++) {
      if (i < R) {
        const uint32_t d = (k[i] >> shift) & 255u;
        const uint32_t peers = vr_peers(d, few_values);
        const uint32_t b4 = my_hist[d];
        __syncwarp();
        if ((peers & lt_mask) == 0) my_hist[d] = (uint16_t)(b4 + __popc(peers));
        __syncwarp();
        m[i] = (m[i] & VR_IDX_MASK) | ((b4 + __popc(peers & lt_mask)) << 18);
      }
    }
    __syncthreads();
    VR_STAMP(2 + 8 * pass);  // ranked
    if (tid < 256) {
      uint32_t total = 0;
#pragma unroll 8
      for (int w = 0; w < VR_WARPS; w++) {  // thread d: exclusive prefix over the warps
        const uint32_t c = sm.warp_hist[w][tid];
        sm.warp_hist[w][tid] = (uint16_t)total;
        total += c;
      }
      sm.cta_cnt[tid] = total;
    }
    VR_STAMP(3 + 8 * pass);  // own digit counts written
    cluster.sync();  // every CTA's digit counts are in place (and every CTA has read back the previous pass)
    VR_STAMP(4 + 8 * pass);  // first cluster barrier passed
    uint32_t total = 0, before = 0;
    if (tid < 256) {
#pragma unroll
      for (uint32_t r = 0; r < VR_CLUSTER; r++) {
        const uint32_t c = *cluster.map_shared_rank(&sm.cta_cnt[tid], r);
        total += c;
        before += r < rank ? c : 0u;
      }
    }
    // exclusive scan over the 256 digits (warps 0-7; sm.scan was last read before the cluster barrier)
    uint32_t incl = total;
    if (warp < 8) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) sm.scan[warp] = incl;
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t off = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) off += w < warp ? sm.scan[w] : 0u;
      sm.base[tid] = off + incl - total + before;
    }
    __syncthreads();
    VR_STAMP(5 + 8 * pass);  // destinations known
#pragma unroll
    for (int i = 0; i < VR_ITEMS; i++) {
      if (i < R) {
        const uint32_t d = (k[i] >> shift) & 255u;
        const uint32_t pos = sm.base[d] + my_hist[d] + (m[i] >> 18);
        const uint32_t owner = ((pos >> 10) * inv_R) >> 16;
        *cluster.map_shared_rank(&buf[pos - owner * per], owner) = make_uint2(k[i], m[i] & VR_IDX_MASK);
      }
    }
    __syncwarp();  // the warp has read its prefixes: clear its counters for the next pass
#pragma unroll
    for (int e = 0; e < 4; e++) reinterpret_cast<uint32_t*>(my_hist)[e * 32 + lane] = 0u;
    VR_STAMP(6 + 8 * pass);  // scattered
    cluster.sync();  // every key of the pass sits in its owner's shared memory
    VR_STAMP(7 + 8 * pass);  // second cluster barrier passed
#pragma unroll
    for (int i = 0; i < VR_ITEMS; i++) {
      if (i < R) {
        const uint2 kv = buf[local_base + i * 32 + lane];
        k[i] = kv.x;
        m[i] = kv.y;
      }
    }
  }

  // ---- the Gaussian of every sorted slot (gathered once), the result, and the scan of tiles_touched in that order ----
  // (keys0 was last read in the load phase, many cluster barriers ago: it can be overwritten now; the tile counts are
  // read-only: both global round trips of the gather chain run before the barrier, not after it)
#pragma unroll
  for (int i = 0; i < VR_ITEMS; i++) {
    if (i < R) {
      const uint32_t loc = glob_base + i * 32 + lane;
      m[i] = loc < n ? vals0[seg_b + m[i]] : 0u;
    }
  }
#pragma unroll
  for (int i = 0; i < VR_ITEMS; i++) {
    if (i < R) {
      const uint32_t loc = glob_base + i * 32 + lane;
      if (loc < n) {
        keys0[seg_b + loc] = ((uint64_t)(uint32_t)v << 32) | k[i];
        k[i] = tiles_touched[m[i]];
      } else {
        k[i] = 0u;
      }
    }
  }
  VR_STAMP(33);  // passes done, Gaussian indices and tile counts gathered
  uint32_t carry = 0;
#pragma unroll
  for (int i = 0; i < VR_ITEMS; i++) {
    if (i < R) {
      uint32_t incl = k[i];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
      }
      k[i] = carry + incl;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if (lane == 0) sm.scan[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = sm.scan[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sm.total = t;
  }
  VR_STAMP(34);  // tile counts scanned inside the CTA
  // ONE barrier: every read of the unsorted vals0 (by any CTA) precedes the first write of the sorted one, and every
  // CTA's total is in place
  cluster.sync();
  VR_STAMP(35);
  uint32_t run = v ? offsets[(size_t)v * P - 1] : 0u;  // pairs of the views before mine (preprocess's scan)
  for (uint32_t r = 0; r < rank; r++) run += *cluster.map_shared_rank(&sm.total, r);
#pragma unroll 8
  for (int w = 0; w < VR_WARPS; w++) run += w < warp ? sm.scan[w] : 0u;
#pragma unroll
  for (int i = 0; i < VR_ITEMS; i++) {
    if (i < R) {
      const uint32_t loc = glob_base + i * 32 + lane;
      if (loc < n) {
        vals0[seg_b + loc] = m[i];
        sorted_offsets[seg_b + loc] = run + k[i];
      }
    }
  }
  VR_STAMP(36);
  cluster.sync();  // no CTA leaves while a peer may still read its shared memory
  VR_STAMP(37);
}

inline bool vis_sort_registers_enabled() {  // OCRF_VIS_SORT_REG=0: every view through visible_sort_kernel (A/B measurements)
  static const bool on = !(getenv("OCRF_VIS_SORT_REG") != nullptr && atoi(getenv("OCRF_VIS_SORT_REG")) == 0);
  return on;
}

// keys0 / vals0: where the preprocess compacted the visible Gaussians AND where the sorted result lands (4 passes);
// keys1 / vals1: the other half of the ping-pong.
int visible_sort(cudaStream_t st, const OcrfShape* sh, const uint32_t* view_start, uint64_t* keys0, uint32_t* vals0,
                 uint64_t* keys1, uint32_t* vals1, const uint32_t* tiles_touched, const uint32_t* offsets,
                 uint32_t* sorted_offsets) {
  uint32_t skip_upto = 0;
  if (vis_sort_registers_enabled()) {
    static int configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(visible_sort_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(VrFixed) + (size_t)VR_ITEMS * VR_THREADS * sizeof(uint2)));
      if (e != cudaSuccess) return (int)e;
      if (dev >= 0 && dev < 64) configured[dev] = 1;
    }
    // no view has more visible Gaussians than the sample has Gaussians: size the slots for that
    const size_t per_row = (size_t)VR_CLUSTER * VR_THREADS;
    const size_t rows = std::min<size_t>(VR_ITEMS, ((size_t)std::max(sh->P, 1) + per_row - 1) / per_row);
    OCRF_LAUNCH(visible_sort_reg_kernel, dim3(VR_CLUSTER, sh->V), dim3(VR_THREADS),
                sizeof(VrFixed) + rows * VR_THREADS * sizeof(uint2), st, sh->P, view_start, keys0, vals0, tiles_touched,
                offsets, sorted_offsets);
    if ((uint32_t)sh->P <= VR_CAP) return 0;
    skip_upto = VR_CAP;
  }
  OCRF_LAUNCH(visible_sort_kernel, dim3(VS_CLUSTER, sh->V), dim3(VS_THREADS), 0, st, sh->P, view_start, keys0, vals0, keys1,
              vals1, tiles_touched, offsets, sorted_offsets, skip_upto);
  return 0;
}

}  // namespace ocrf
