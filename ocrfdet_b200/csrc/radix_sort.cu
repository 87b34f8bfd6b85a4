// Hand-written onesweep radix sort of (uint64 key, uint32 value) pairs -- stage 2 of the render path.
//
// Replaces cub::DeviceRadixSort::SortPairs as the reference calls it
// (cuda_rasterizer/rasterizer_impl.cu:303-308): stable, ascending, over key bits [0, end_bit).
// Any stable sort on those bits produces the same permutation, so the output is bit-identical to
// the reference's point_list.
//
// Structure (Adinets & Merrill, "Onesweep"): one histogram kernel counts all digit places in a
// single read of the keys; then one kernel per 8-bit digit place in which every CTA
//   (1) takes a ticket (so chained waits only ever look at CTAs that already run),
//   (2) ranks its 4096 keys with warp-level match_any multi-split (stable inside the CTA),
//   (3) publishes its per-digit counts and resolves its global offsets by decoupled look-back,
//   (4) reorders keys/values through shared memory so the global scatter is written in runs.
// HBM traffic per pass is one read and one write of the 12-byte pairs -- the minimum for an LSD
// pass -- and the look-back state is 1 KB per CTA.
//
// The element count lives on the device (geom header) so the whole pipeline can run without a host
// round trip; `n_cap` only bounds the grid.
#include "common.cuh"

namespace ocrf {

constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr uint32_t LB_FLAG_LOCAL = 1u << 30;
constexpr uint32_t LB_FLAG_INCL = 2u << 30;
constexpr uint32_t LB_VALUE_MASK = (1u << 30) - 1;

__device__ __forceinline__ uint32_t effective_n(const uint32_t* n_dev, uint64_t n_cap) {
  const uint32_t n = *n_dev;
  return (uint64_t)n <= n_cap ? n : 0u;  // over capacity: sort nothing (the error flag is raised elsewhere)
}

__global__ void __launch_bounds__(SORT_THREADS) sort_histogram_kernel(const uint64_t* __restrict__ keys,
                                                                      const uint32_t* __restrict__ n_dev,
                                                                      uint64_t n_cap, int passes, int begin_bit,
                                                                      int end_bit, uint32_t* __restrict__ hist,
                                                                      uint32_t* __restrict__ zero_extra,
                                                                      uint32_t zero_words) {
  pdl_enter();
  __shared__ uint32_t s_hist[SORT_MAX_PASSES * 256];
  const uint32_t n = effective_n(n_dev, n_cap);
  for (uint32_t i = blockIdx.x * SORT_THREADS + threadIdx.x; i < zero_words; i += gridDim.x * SORT_THREADS) zero_extra[i] = 0u;
  for (int i = threadIdx.x; i < passes * 256; i += SORT_THREADS) s_hist[i] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * SORT_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * SORT_THREADS) {
    const uint64_t k = keys[i];
    for (int p = 0; p < passes; p++) {
      const int shift = begin_bit + p * 8;
      const int bits = min(8, end_bit - shift);
      atomicAdd(&s_hist[p * 256 + (uint32_t)((k >> shift) & ((1u << bits) - 1))], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += SORT_THREADS) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// exclusive scan of one value per thread over the 256-thread CTA
__device__ __forceinline__ uint32_t cta_exclusive_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t off = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++)
    if (w < warp) off += s_warp[w];
  __syncthreads();
  return off + incl - v;
}

template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS, ITEMS >= 16 ? 3 : 4) sort_onesweep_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ n_dev, uint64_t n_cap, int shift, int bits,
    const uint32_t* __restrict__ hist /*[256] this pass*/, uint32_t* __restrict__ status /*[tiles][256] this pass*/,
    uint32_t* __restrict__ ticket) {
  pdl_enter();
  constexpr int TILE_N = SORT_THREADS * ITEMS;  // keys per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);                                // [TILE_N]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw + TILE_N * 8);                // [TILE_N]
  uint32_t* s_warp_hist = reinterpret_cast<uint32_t*>(smem_raw + TILE_N * 12);          // [SORT_WARPS][256]
  uint32_t* s_tile_excl = s_warp_hist + SORT_WARPS * 256;                                   // [256]
  uint32_t* s_digit_base = s_tile_excl + 256;                                               // [256]
  uint32_t* s_scan = s_digit_base + 256;                                                    // [8]
  __shared__ uint32_t s_tile_id;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = effective_n(n_dev, n_cap);
  if (tid == 0) s_tile_id = atomicAdd(ticket, 1u);
  for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) s_warp_hist[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile_id;
  const uint64_t base = (uint64_t)tile * TILE_N;
  if (base >= n) return;
  const uint32_t n_valid = (uint32_t)min((uint64_t)TILE_N, (uint64_t)n - base);
  const uint32_t digit_mask = (1u << bits) - 1;

  // ---- load (warp-striped: item i of lane l sits at warp_base + i*32 + l, i.e. index order) ----
  uint64_t key[ITEMS];
  const uint32_t warp_base = warp * (32 * ITEMS);
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const uint32_t loc = warp_base + i * 32 + lane;
    key[i] = loc < n_valid ? keys_in[base + loc] : ~0ull;  // padding ranks behind every valid key of the (last) tile
  }

  // ---- publish this tile's digit counts EARLY (cheap shared-memory histogram) so that successors can
  //      resolve their look-back while we are still ranking; the stable ranks come afterwards ----
  s_tile_excl[tid] = 0;  // reused as the counting histogram until the scan below
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) atomicAdd(&s_tile_excl[(uint32_t)(key[i] >> shift) & digit_mask], 1u);
  __syncthreads();
  const uint32_t tile_count = s_tile_excl[tid];
  uint32_t* my_status = status + (size_t)tile * 256 + tid;
  atomicExch(my_status, (tile == 0 ? LB_FLAG_INCL : LB_FLAG_LOCAL) | tile_count);

  // ---- rank inside the warp: match_any multi-split, one item row at a time (stable) ----
  uint32_t rank[ITEMS];
  uint32_t* my_hist = s_warp_hist + warp * 256;
  const uint32_t lt_mask = (1u << lane) - 1;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t before = my_hist[d];
    __syncwarp();
    if ((peers & lt_mask) == 0) my_hist[d] = before + __popc(peers);  // lowest lane of the group updates
    __syncwarp();
    rank[i] = before + __popc(peers & lt_mask);
  }
  __syncthreads();

  // ---- per digit (thread t owns digit t): exclusive warp offsets ----
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
      const uint32_t c = s_warp_hist[w * 256 + tid];
      s_warp_hist[w * 256 + tid] = run;  // exclusive offset of warp w for this digit
      run += c;
    }
  }
  const uint32_t tile_excl = cta_exclusive_scan_256(tile_count, s_scan);
  const uint32_t global_digit_excl = cta_exclusive_scan_256(hist[tid], s_scan);
  uint32_t lookback = 0;
  if (tile > 0) {
    int look = (int)tile - 1;
    while (true) {
      const uint32_t st = *reinterpret_cast<volatile uint32_t*>(status + (size_t)look * 256 + tid);
      if ((st >> 30) == 0) continue;
      lookback += st & LB_VALUE_MASK;
      if ((st >> 30) == 2) break;
      look--;
    }
    atomicExch(my_status, LB_FLAG_INCL | (lookback + tile_count));
  }
  s_tile_excl[tid] = tile_excl;
  s_digit_base[tid] = global_digit_excl + lookback - tile_excl;
  __syncthreads();

  // ---- reorder through shared memory, then write runs ----
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
    const uint32_t pos = s_tile_excl[d] + s_warp_hist[warp * 256 + d] + rank[i];
    s_keys[pos] = key[i];
    const uint32_t loc = warp_base + i * 32 + lane;
    s_vals[pos] = loc < n_valid ? vals_in[base + loc] : 0u;  // values are only touched here (coalesced, L2-hot)
  }
  __syncthreads();
  for (uint32_t k = tid; k < n_valid; k += SORT_THREADS) {
    const uint64_t kk = s_keys[k];
    const uint32_t d = (uint32_t)(kk >> shift) & digit_mask;
    const uint32_t dst = s_digit_base[d] + k;
    keys_out[dst] = kk;
    vals_out[dst] = s_vals[k];
  }
}

template <int ITEMS>
constexpr size_t sort_smem_bytes() { return (size_t)SORT_THREADS * ITEMS * 12 + (SORT_WARPS * 256 + 256 + 256 + 8) * 4; }

template <int ITEMS>
static int launch_onesweep(cudaStream_t st, unsigned tiles, const uint64_t* kin, const uint32_t* vin, uint64_t* kout,
                           uint32_t* vout, const uint32_t* n_dev, uint64_t n_cap, int shift, int bits,
                           const uint32_t* hist, uint32_t* status, uint32_t* ticket) {
  static unsigned long long attr_done = 0;  // per-device bit mask
  const cudaError_t ae = ensure_dynamic_smem(sort_onesweep_kernel<ITEMS>, sort_smem_bytes<ITEMS>(), attr_done);
  if (ae != cudaSuccess) return (int)ae;
  OCRF_LAUNCH(sort_onesweep_kernel<ITEMS>, dim3(tiles), dim3(SORT_THREADS), sort_smem_bytes<ITEMS>(), st, kin, vin, kout, vout, n_dev, n_cap, shift,
                                                                                    bits, hist, status, ticket);
  return 0;
}

// Sorts n_cap-bounded pairs; data starts in (keys_a, vals_a) and the result lands in (keys_b, vals_b)
// when `passes` is odd, in (keys_a, vals_a) when it is even -- callers pick a/b accordingly.
int sort_pairs_device(cudaStream_t st, const uint32_t* n_dev, uint64_t n_cap, int begin_bit, int end_bit,
                      uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, void* ws, bool ws_zeroed,
                      uint32_t* zero_extra, uint32_t zero_words) {
  if (begin_bit < 0 || end_bit <= begin_bit || end_bit > 64) return OCRF_EINVAL;
  const int passes = (end_bit - begin_bit + 7) / 8;
  if (n_cap == 0) return 0;
  const bool small = n_cap <= SORT_SMALL_MAX;
  const SortWs L = sort_ws_layout(n_cap);
  const uint64_t tiles = sort_tiles(n_cap);
  if (!ws_zeroed) cudaMemsetAsync(ws, 0, L.status + (size_t)passes * (tiles + 1) * 256 * 4, st);
  uint32_t* hist = at<uint32_t>(ws, L.hist);
  const int hgrid = (int)min((uint64_t)num_sms() * 8, (n_cap + SORT_THREADS * 4 - 1) / (SORT_THREADS * 4));
  OCRF_LAUNCH(sort_histogram_kernel, dim3(hgrid), dim3(SORT_THREADS), 0, st, keys_a, n_dev, n_cap, passes, begin_bit, end_bit, hist,
               zero_extra, zero_words);
  uint64_t* kin = keys_a;
  uint32_t* vin = vals_a;
  uint64_t* kout = keys_b;
  uint32_t* vout = vals_b;
  for (int p = 0; p < passes; p++) {
    const int shift = begin_bit + p * 8;
    const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
    uint32_t* status = at<uint32_t>(ws, L.status) + (size_t)p * (tiles + 1) * 256;
    const int rc = small ? launch_onesweep<SORT_ITEMS_SMALL>(st, (unsigned)tiles, kin, vin, kout, vout, n_dev, n_cap, shift,
                                                             bits, hist + p * 256, status, at<uint32_t>(ws, L.ticket) + p)
                         : launch_onesweep<SORT_ITEMS>(st, (unsigned)tiles, kin, vin, kout, vout, n_dev, n_cap, shift, bits,
                                                       hist + p * 256, status, at<uint32_t>(ws, L.ticket) + p);
    if (rc != 0) return rc;
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  OCRF_CHECK_LAST();
  return 0;
}

__global__ void store_u32_kernel(uint32_t* p, uint32_t v) { *p = v; }

}  // namespace ocrf

using namespace ocrf;

extern "C" size_t ocrf_sort_workspace_bytes(uint64_t n) { return sort_ws_layout(n).total + 128; }

extern "C" int ocrf_sort_pairs(void* stream, uint64_t n, int end_bit, const uint64_t* keys_in, const uint32_t* vals_in,
                               uint64_t* keys_out, uint32_t* vals_out, uint64_t* keys_tmp, uint32_t* vals_tmp,
                               void* ws) {
  if (n == 0) return 0;
  if (!keys_in || !vals_in || !keys_out || !vals_out || !keys_tmp || !vals_tmp || !ws) return OCRF_EINVAL;
  if (n >= (1ull << 30)) return OCRF_ECAPACITY;
  if (end_bit <= 0 || end_bit > 64) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int passes = (end_bit + 7) / 8;
  // stage the input so that the ping-pong ends in keys_out
  uint64_t* ka = (passes & 1) ? keys_tmp : keys_out;
  uint32_t* va = (passes & 1) ? vals_tmp : vals_out;
  uint64_t* kb = (passes & 1) ? keys_out : keys_tmp;
  uint32_t* vb = (passes & 1) ? vals_out : vals_tmp;
  cudaMemcpyAsync(ka, keys_in, n * 8, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(va, vals_in, n * 4, cudaMemcpyDeviceToDevice, st);
  // the element count is read on the device: park it in the last 4 bytes of the workspace
  const SortWs L = sort_ws_layout(n);
  uint32_t* n_dev = at<uint32_t>(ws, L.total);
  int rc = 0;
  store_u32_kernel<<<1, 1, 0, st>>>(n_dev, (uint32_t)n);
  rc = sort_pairs_device(st, n_dev, n, 0, end_bit, ka, va, kb, vb, ws);
  return rc;
}
