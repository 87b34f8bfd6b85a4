// Shared device/host helpers of libocrf_raster.so.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

#include "ocrf_raster.h"

namespace ocrf {

constexpr int TILE = 16;               // blend tile edge (config.h:16-17 of the reference: BLOCK_X = BLOCK_Y = 16)
constexpr int TILE_PIX = TILE * TILE;  // pixels per tile
constexpr int NUM_SMS_B200 = 148;      // B200; only the fallback of num_sms() below

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs).
inline int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return NUM_SMS_B200;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = NUM_SMS_B200;
    cache[dev] = n;
  }
  return cache[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: remember per device ordinal
// (`done_mask`, one static word per call site) where it has been raised already.
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kernel, size_t bytes, unsigned long long& done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && ((done_mask >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done_mask |= 1ull << dev;
  return e;
}

// geom header slots (uint32)
constexpr int HDR_NUM_PAIRS = 0;
constexpr int HDR_ERROR = 1;
constexpr int HDR_TICKET = 2;
constexpr int HDR_NUM_VIS = 3;  // visible (view, Gaussian) pairs of the batch
constexpr uint32_t ERR_PAIR_OVERFLOW = 1u;
constexpr uint32_t ERR_PREFILTERED = 2u;

// Sticky status words of a caller (ocrf_bin_forward's `sticky_status`): the geom header is re-zeroed by every
// forward, so a caller that does not read it back after each call (capacity mode, graph replays) gives the library
// two words that are only ever OR-ed / max-ed into: [0] error flags of any call so far, [1] largest pair count.
__device__ __forceinline__ void publish_status(const uint32_t* header, uint64_t n_cap, uint32_t* sticky) {
  const uint32_t n = header[HDR_NUM_PAIRS];
  const uint32_t e = header[HDR_ERROR] | ((uint64_t)n > n_cap ? ERR_PAIR_OVERFLOW : 0u);
  if (e) atomicOr(&sticky[0], e);
  atomicMax(&sticky[1], n);
}

struct Camera {  // OCRF_CAM_STRIDE floats
  float view[16];
  float proj[16];
  float campos[3];
  float tanfovx, tanfovy;
  float pad[3];
};
static_assert(sizeof(Camera) == OCRF_CAM_STRIDE * 4, "camera record size");

// One (tile, Gaussian) pair as the blend kernels consume it (C == 3).
// The conic (A, B, C) is stored PRE-SCALED into the exponent the kernels evaluate: with d = mean - pixel,
//   log2(G) = power * log2(e) = qa dx^2 + qb dx dy + qc dy^2,   qa = -A log2(e)/2, qb = -B log2(e), qc = -C log2(e)/2
// (the reference's power = -0.5 (A dx^2 + C dy^2) - B dx dy, forward.cu:336), so a test costs no -0.5 and no log2(e)
// multiply per (pixel, Gaussian); the backward recovers A, B, C once per record (record_conic below).
struct __align__(16) Record {
  float x, y, qa, qb;  // pixel-space mean, scaled conic A, B     (read for every test)
  float qc, op;        // scaled conic C, opacity                  (read for every test)
  uint32_t orig;       // 1-based position in the tile's full (unculled) sorted list == the reference's "contributor"
  float r;             // red
  float g, b;          // green, blue                              (read only when blending)
  uint32_t id;         // Gaussian index inside its sample
  float depth;         // view-space depth (median-depth output)
};
static_assert(sizeof(Record) == OCRF_RECORD_BYTES, "record size");

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
__device__ __forceinline__ float4 record_head(float2 xy, float4 conic_opacity) {  // first float4 of a Record
  return make_float4(xy.x, xy.y, conic_opacity.x * (-0.5f * LOG2E), conic_opacity.y * (-LOG2E));
}
__device__ __forceinline__ float record_qc(float4 conic_opacity) { return conic_opacity.z * (-0.5f * LOG2E); }
__device__ __forceinline__ void record_conic(float qa, float qb, float qc, float& A, float& B, float& C) {
  A = qa * (-2.f * LN2);
  B = qb * (-LN2);
  C = qc * (-2.f * LN2);
}

__host__ __device__ inline size_t align128(size_t x) { return (x + 127) & ~size_t(127); }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

inline int tiles_x(const OcrfShape& s) { return (s.W + TILE - 1) / TILE; }
inline int tiles_y(const OcrfShape& s) { return (s.H + TILE - 1) / TILE; }

template <typename T>
__host__ __device__ inline T* at(void* base, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(base) + off); }
template <typename T>
__host__ __device__ inline const T* at(const void* base, size_t off) {
  return reinterpret_cast<const T*>(static_cast<const char*>(base) + off);
}

// ---- radix sort geometry (shared between layout code and kernels) ----
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // keys per CTA per pass
// Optional quarter-size tile for small inputs.  Measured on the (view | depth) sort of ~114k visible Gaussians:
// no gain, a pass is bound by its fixed chain of L2 round trips (ticket, keys, look-back), not by per-CTA work.
constexpr int SORT_ITEMS_SMALL = 4;
constexpr uint64_t SORT_SMALL_MAX = 0;  // capacity at or below which the small tile is used (measured: no gain -> off)
inline uint64_t sort_tile_size(uint64_t n_cap) {
  return (uint64_t)SORT_THREADS * (n_cap <= SORT_SMALL_MAX ? SORT_ITEMS_SMALL : SORT_ITEMS);
}
constexpr int SORT_MAX_PASSES = 8;

inline uint64_t sort_tiles(uint64_t n) { return (n + sort_tile_size(n) - 1) / sort_tile_size(n); }

struct SortWs {
  size_t hist;    // uint32 [SORT_MAX_PASSES][256]
  size_t ticket;  // uint32 [SORT_MAX_PASSES] (padded to 128 B)
  size_t status;  // uint32 [SORT_MAX_PASSES][tiles + 1][256]
  size_t total;
};
inline SortWs sort_ws_layout(uint64_t n) {
  SortWs w;
  size_t off = 0;
  w.hist = off;   off = ((off + SORT_MAX_PASSES * 256 * 4) + 127) & ~size_t(127);
  w.ticket = off; off = ((off + SORT_MAX_PASSES * 4) + 127) & ~size_t(127);
  w.status = off; off = ((off + (size_t)SORT_MAX_PASSES * (sort_tiles(n) + 1) * 256 * 4) + 127) & ~size_t(127);
  w.total = off;
  return w;
}
// radix_sort.cu: stable sort on key bits [begin_bit, end_bit).  Data starts in (keys_a, vals_a); result in
// (keys_b, vals_b) when the pass count ceil((end_bit-begin_bit)/8) is odd, back in (keys_a, vals_a) when even.
// n is read from *n_dev (<= n_cap).
// ws_zeroed: the caller has already zeroed the workspace (it sits inside a larger memset);
// zero_extra/zero_words: a small region the histogram kernel clears on the side (saves a memset node between
// two chain kernels).
int sort_pairs_device(cudaStream_t st, const uint32_t* n_dev, uint64_t n_cap, int begin_bit, int end_bit,
                      uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, void* ws,
                      bool ws_zeroed = false, uint32_t* zero_extra = nullptr, uint32_t zero_words = 0);

// ---- preprocess geometry ----
constexpr int PRE_THREADS = 256;

// ---- depth-first binning (binning.cu) ----
inline int view_bits(int V) {            // bits needed for view indices 0..V-1
  int b = 0;
  while (b < 31 && ((V - 1) >> b)) b++;
  return b;
}
inline int vis_sort_end_bit(int V) { return 32 + view_bits(V); }
// The visible Gaussians are sorted by the on-chip cluster sort (visible_sort.cu: 4 passes over the depth bits, input
// and result in the same half of the ping-pong) unless OCRF_VIS_SORT=global asks for round 1's global onesweep sort
// over (view | depth) (A/B measurements; its pass count decides in which half the compacted input must start).
inline bool vis_sort_onchip() {
  static const bool on = !(getenv("OCRF_VIS_SORT") != nullptr && getenv("OCRF_VIS_SORT")[0] == 'g');
  return on;
}
inline bool vis_sort_starts_in_tmp(int V) { return vis_sort_onchip() ? false : ((((vis_sort_end_bit(V) + 7) / 8) & 1) != 0); }
int visible_sort(cudaStream_t st, const OcrfShape* sh, const uint32_t* view_start, uint64_t* keys0, uint32_t* vals0,
                 uint64_t* keys1, uint32_t* vals1, const uint32_t* tiles_touched, const uint32_t* offsets,
                 uint32_t* sorted_offsets);
// multisplit.cu
struct Record;
int multisplit_bin(cudaStream_t st, const OcrfShape* sh, uint64_t pair_capacity, int use_sh, const int32_t* radii,
                   const float* colors, uint32_t* header, const uint32_t* view_start, const uint32_t* sorted_offsets,
                   const uint32_t* vis_vals, const float2* xy, const float4* conic_opacity, const float* depths,
                   const float* rgb, uint32_t* tables, size_t table_words, uint32_t* tile_arrays, uint2* items,
                   uint2* ranges, uint2* ranges_render, Record* records, uint32_t* sticky);
inline uint32_t multisplit_chunk_pairs(uint64_t pair_capacity) {  // pairs per chunk: multiples of 4096, <= ~2048 chunks
  uint64_t rounds = (pair_capacity + 4096ull * 2048 - 1) / (4096ull * 2048);
  return (uint32_t)((rounds < 1 ? 1 : rounds) * 4096);
}

// render_tc_fwd.cu / render_tc_bwd.cu: the many-channel blend on tcgen05 (32 < C <= 80, C % 4 == 0; OCRF_TC=0 disables)
namespace tc {
bool forward_tc_supported(int C);
int launch_forward_tc(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                      const float* colors, const float* bg, float* fT, uint32_t* nc, uint32_t* mc, float* out_color,
                      float* out_depth, float* out_opacity);
int launch_backward_tc(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                       const float* colors, const float* bg, const float* fT, const uint32_t* nc, const uint32_t* mc,
                       const float* dL_dcolor, const float* dL_dopa, double* ggrad, float* dL_dcolors);
}  // namespace tc

// ---- PTX helpers: mbarrier + 1D bulk async copy (TMA unit, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared -> global bulk copy (bulk async-group completion); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA unit)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- programmatic dependent launch (PDL) for the kernels of the render chain ----
// The path is a chain of ~15 small dependent kernels; with plain stream order every link pays the full launch
// latency (~3-4 us) after its predecessor has drained.  Chain kernels are launched with the programmatic stream
// serialization attribute and begin with pdl_enter(): griddepcontrol.wait (the predecessor grid has completed and
// its writes are visible) followed by griddepcontrol.launch_dependents (the successor may be scheduled as soon as
// every CTA of this grid has started).  Nothing touches global memory before pdl_enter(), so the semantics are
// exactly those of stream order; only the launch latency and the CTA ramp overlap with the predecessor's tail.
// OCRF_PDL=0 in the environment falls back to plain launches.  A kernel launched without the attribute treats
// both instructions as no-ops.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool on = !(getenv("OCRF_PDL") != nullptr && atoi(getenv("OCRF_PDL")) == 0);
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

// ---- decoupled look-back, resolved by one warp (32 predecessors per memory round trip) ----
// status words: bits 63:62 = 0 not published, 1 = the CTA's own total, 2 = inclusive prefix; bits 61:0 = value.
constexpr unsigned long long LBK_FLAG_LOCAL = 1ull << 62;
constexpr unsigned long long LBK_FLAG_INCL = 2ull << 62;
constexpr unsigned long long LBK_VALUE_MASK = (1ull << 62) - 1;

// Called by all 32 lanes of ONE warp of CTA `bid` (chain order = bid).  Publishes `total`, returns the exclusive
// prefix of the chain (same value in every lane) and publishes the inclusive one.
__device__ __forceinline__ unsigned long long lookback_warp(unsigned long long* status, int bid,
                                                            unsigned long long total) {
  const int lane = threadIdx.x & 31;
  if (bid == 0) {
    if (lane == 0) atomicExch(&status[0], LBK_FLAG_INCL | total);
    return 0ull;
  }
  if (lane == 0) atomicExch(&status[bid], LBK_FLAG_LOCAL | total);
  unsigned long long excl = 0;
  int look = bid - 1;
  while (true) {
    const int idx = look - lane;
    unsigned long long st = 2ull << 62;  // virtual inclusive 0 before the start of the chain
    if (idx >= 0) st = *reinterpret_cast<volatile unsigned long long*>(&status[idx]);
    const unsigned flag = (unsigned)(st >> 62);
    const unsigned incl_mask = __ballot_sync(0xffffffffu, flag == 2u);
    const unsigned pending = __ballot_sync(0xffffffffu, flag == 0u);
    const int first = incl_mask ? (__ffs(incl_mask) - 1) : 31;
    const unsigned need = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
    if (pending & need) continue;  // a predecessor we need has not published yet
    unsigned long long v = lane <= first ? (st & LBK_VALUE_MASK) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (incl_mask) break;
    look -= 32;
  }
  if (lane == 0) atomicExch(&status[bid], LBK_FLAG_INCL | (excl + total));
  return excl;
}

// ---- packed FP32 (sm_100: FFMA2 / FMUL2 / FADD2 issue two fp32 operations per lane per instruction) ----
// A float2 whose halves are the same value is encoded by ptxas as a broadcast operand (Rx.F32): no move is spent.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmov.b64 rc, {%6,%7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmul.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nadd.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exp() for the blend: one MUFU.EX2.  |x| <= ~6 on every path whose result is used, so the
// absolute error of x*log2(e) is < 5e-7 and the relative error of the result < 4e-7.
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.4426950408889634f); }

}  // namespace ocrf

// launch_chain() whose failure leaves the enclosing int-returning function with the cudaError_t
#define OCRF_LAUNCH(...)                                  \
  do {                                                    \
    cudaError_t le__ = ocrf::launch_chain(__VA_ARGS__);   \
    if (le__ != cudaSuccess) return (int)le__;            \
  } while (0)

#define OCRF_CHECK_LAST()                   \
  do {                                      \
    cudaError_t e__ = cudaGetLastError();   \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)
