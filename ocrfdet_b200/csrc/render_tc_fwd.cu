// Stage 3 for many feature channels (32 < C <= 80, BASELINE config 4): the colour accumulation of the forward blend
// on the 5th-generation tensor cores.
//
// With C channels the blend of a 16x16 tile is  out[256 px][C] = W[256 px][R records] . F[R][C],  W = alpha * T being
// the blend weight the scalar recurrence of renderCUDA (cuda_rasterizer/forward.cu:330-360) produces per (pixel,
// record).  render_forward_wide_kernel spends C/2 FFMA2 + C/4 LDS.128 per (pixel, record) on that product; here the
// threads only run the scalar part and the product is issued as tcgen05.mma (kind::tf32, M = 128, N = C, K = 8):
//
//   roles      8 worker warps (one thread per pixel; TMEM lane = pixel % 128, M-block = pixel / 128)
//              1 MMA warp (one elected lane issues tcgen05.mma / tcgen05.commit)
//              1 loader warp (gathers the feature rows of the next records)
//   A operand  the weights, written by their own threads into tensor memory with tcgen05.st (16 columns per 8 records:
//              w and w_lo = w - trunc_tf32(w)); no shared-memory traffic, no layout arithmetic
//   B operand  F^T [C x 8 records] K-major in shared memory, staged by the loader warp: cp.async gathers the rows
//              of the records three blocks ahead into a landing ring, the transposed core-matrix image (f and f_lo)
//              is written from there
//   D          2 x C fp32 columns of tensor memory (the two M-blocks of the tile), read once in the epilogue
//   precision  3xTF32: w.f + w_lo.f + w.f_lo (the tensor core truncates fp32 to tf32, so "hi" is the value itself);
//              error ~ 2^-21 relative per term, inside the 1e-5 forward bar (tests/test_gpu_parity.py C = 40, 80)
//   pipeline   three stages of 8 records; mbarriers ready_f (loader -> workers, MMA), ready_w (workers -> MMA),
//              free (tcgen05.commit -> loader, workers).  Tensor memory: 160 D + 3 x 32 A = 256 columns -> 2 CTAs/SM.
//   scalar     two branch-free phases per block: all eight alphas (independent: ILP), then the T recurrence.
// Early ray termination: a warp whose pixels are all finished reports it once; when all eight have, the loader
// publishes the stop marker instead of the next block.
#include "tc_common.cuh"

namespace ocrf {
namespace tc {

constexpr int KB = 8;          // records per stage == K of one tf32 tcgen05.mma
constexpr int STAGES = 3;      // weight stages in tensor memory
constexpr int FST = 6;         // feature / record stages in shared memory (the loader runs further ahead than the weights)
constexpr int LOOKAHEAD = 3;   // blocks between the cp.async of a block and its use
constexpr int RING = LOOKAHEAD + 1;
// A CTA blends MB M-blocks of 128 pixels: MB = 2 is a whole 16x16 tile (256 tensor-memory columns, 2 CTAs per SM),
// MB = 1 the upper or lower 16x8 half of one (128 columns, 4 CTAs per SM).  The lists of this path are short where
// the scene is opaque (BASELINE config 4: a tile is finished after ~76 of its ~400 records), so a tile is mostly
// start-up and drain latency; with half tiles twice as many of them are in flight per SM.  The loader work per tile
// doubles (it is idle most of the time), the tensor-core and worker work does not.

__device__ __forceinline__ float ex2_approx_t(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int CP>
struct FwdSmem {
  static constexpr int NG = CP / 8;                 // groups of 8 channels = core-matrix rows of B
  static constexpr uint32_t LBO = 128;              // the two 4-record chunks of a stage: adjacent core matrices
  static constexpr uint32_t SBO = 256;              // next group of 8 channels
  static constexpr uint32_t PART = NG * 256;        // bytes of one operand image (f or f_lo) of one stage
  alignas(128) unsigned char f[FST][2][PART];       // [stage][raw | lo]
  alignas(128) Record rec[FST][KB];
  alignas(128) float land_f[RING][KB][CP];          // cp.async landing zone (row = record)
  alignas(128) Record land_rec[RING][KB];
  alignas(8) uint64_t ready_f[FST], free_f[FST], ready_w[STAGES], free_[STAGES], final_, meta_;
  int cnt[FST];
  float bg[CP];
  uint32_t nonzero[STAGES];  // worker warps whose weights of the stage's block are not all zero
  uint32_t tmem, done_warps, max_contrib, started;
};

template <int CP, int MB>
__global__ void __launch_bounds__(MB * 128 + 64, 4 / MB) render_forward_tc_kernel(
    int W, int H, int C, int P, int views_per_sample, const uint2* __restrict__ ranges /* culled lists */,
    const Record* __restrict__ records, const float* __restrict__ feats, const float* __restrict__ bg,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ max_contrib,
    float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_opacity) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FwdSmem<CP>& sm = *reinterpret_cast<FwdSmem<CP>*>(smem_raw);
  constexpr int WORKER_WARPS = MB * 4;
  constexpr uint32_t FWD_TMEM_COLS = MB * 128;
  constexpr uint32_t A_COL0 = MB * 80;  // first column of the weight stages (D occupies [0, MB CP) <= MB 80)
  constexpr int HALVES = 2 / MB;        // CTAs per tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x / HALVES, tiles_per_view = tiles_x * gridDim.y;
  const int view = blockIdx.z;
  const int tile_x = blockIdx.x / HALVES, half = blockIdx.x % HALVES;
  const int tile = blockIdx.y * tiles_x + tile_x;
  const uint2 range = ranges[(size_t)view * tiles_per_view + tile];
  const int n = (int)(range.y - range.x);
  const int nblocks = (n + KB - 1) / KB;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&sm.ready_w[s], WORKER_WARPS);
      mbar_init(&sm.free_[s], 1);
    }
#pragma unroll
    for (int s = 0; s < FST; s++) {
      mbar_init(&sm.ready_f[s], 1);
      mbar_init(&sm.free_f[s], 1);
    }
    mbar_init(&sm.final_, 1);
    mbar_init(&sm.meta_, 1);
    mbar_fence_init();
    sm.done_warps = 0;
    sm.max_contrib = 0;
#pragma unroll
    for (int s = 0; s < STAGES; s++) sm.nonzero[s] = 0;
  }
  if (warp == WORKER_WARPS) tmem_alloc<FWD_TMEM_COLS>(&sm.tmem);
  if (tid < CP) sm.bg[tid] = tid < C ? bg[tid] : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = sm.tmem;

  if (warp < WORKER_WARPS) {
    // ------------------------------------------------ workers ------------------------------------------------
    const int wt = half * 4 + warp;  // warp of the tile: an 8 x 4 pixel block
    const int px = tile_x * TILE + (wt & 1) * 8 + (lane & 7);
    const int py = blockIdx.y * TILE + (wt >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const int mb = warp >> 2;
    const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16);
    // a finished (or out-of-image) pixel is a NaN row coordinate: its power is NaN, every comparison on it is false,
    // so it needs no predicate of its own in the loop
    const float qnan = __int_as_float(0x7fc00000);
    float fyn = inside ? fy : qnan;
    bool reported = false;
    float T = 1.f;
    uint32_t last_idx = 0, cross_idx = 0;  // 1-based positions in the culled list: last blended record / the record
                                           // that took T across 0.5 (their `orig` and depth are fetched once, at the end)
    int b = 0;
    for (;; b++) {
      const int s = b % STAGES, sf = b % FST;
      const uint32_t use = (uint32_t)(b / STAGES), usef = (uint32_t)(b / FST);
      mbar_wait_wd(&sm.ready_f[sf], usef & 1);
      const int cnt = sm.cnt[sf];
      if (cnt < 0) break;
      uint32_t wv[16];
#pragma unroll
      for (int j = 0; j < 16; j++) wv[j] = 0u;
      const bool warp_done = __all_sync(0xffffffffu, fyn != fyn);
      bool nonzero = false;
      if (!warp_done) {
        float al[KB];
        // phase 1: the eight alphas, independent of each other
#pragma unroll
        for (int j = 0; j < KB; j++) {
          const float4 a = reinterpret_cast<const float4*>(&sm.rec[sf][j])[0];
          const float2 c = reinterpret_cast<const float2*>(&sm.rec[sf][j])[2];  // qc, opacity
          const float dx = a.x - fx, dy = a.y - fyn;
          const float power = a.z * dx * dx + c.x * dy * dy + a.w * dx * dy;  // log2 domain (scaled conic)
          const float alpha = fminf(0.99f, c.y * ex2_approx_t(power));
          al[j] = (j < cnt && power <= 0.0f && alpha >= 1.0f / 255.0f) ? alpha : 0.f;
        }
        // Behind the last record that any pixel of the tile blends (position ~76 of ~400 in an opaque scene) the
        // lists go on -- the exact tile cull knows the geometry, not the occlusion -- and nobody in the warp has a
        // non-zero alpha any more: the recurrence is skipped for the whole block (the weights stay zero)
        float amax = al[0];
#pragma unroll
        for (int j = 1; j < KB; j++) amax = fmaxf(amax, al[j]);
        if (__any_sync(0xffffffffu, amax > 0.f)) {
        nonzero = true;
        // phase 2: the transmittance recurrence (forward.cu:343-358: stop BEFORE blending once T would drop below 1e-4)
#pragma unroll
        for (int j = 0; j < KB; j++) {
          const bool ok = al[j] > 0.f && fyn == fyn;
          const float test_T = T * (1.f - al[j]);
          const bool blend = ok && test_T >= 0.0001f;
          const float w = blend ? al[j] * T : 0.f;
          wv[j] = __float_as_uint(w);
          wv[KB + j] = __float_as_uint(tf32_lo(w));
          const uint32_t idx = (uint32_t)(b * KB + j + 1);
          cross_idx = (blend && T > 0.5f && test_T < 0.5f) ? idx : cross_idx;
          last_idx = blend ? idx : last_idx;
          T = blend ? test_T : T;
          fyn = (ok && !blend) ? qnan : fyn;
        }
        }
      } else if (!reported) {
        if (lane == 0) atomicAdd(&sm.done_warps, 1u);
        reported = true;
      }
      mbar_wait_wd(&sm.free_[s], (use & 1) ^ 1);  // the MMAs that read this stage's columns three blocks ago are done
      fence_after_sync();
      tmem_st16(lane_addr + A_COL0 + s * (MB * 16) + mb * 16, wv);
      tmem_wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (nonzero) atomicAdd(&sm.nonzero[s], 1u);  // (ordered before the arrival: the MMA warp reads it after its wait)
        mbar_arrive(&sm.ready_w[s]);
      }
    }
    // ---- epilogue: D -> registers -> out_color
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    mbar_wait_wd(&sm.meta_, 0);
    mbar_wait_wd(&sm.final_, 0);  // every product has completed (the MMA warp commits even when it issued none)
    fence_after_sync();
    const bool have_d = *reinterpret_cast<volatile uint32_t*>(&sm.started) != 0u;
    float* oc = out_color + (size_t)view * C * HW + pix;
#pragma unroll
    for (int c0 = 0; c0 < CP; c0 += 16) {
      uint32_t v[16];
      if (have_d) {
        tmem_ld16(lane_addr + mb * CP + c0, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = 0u;
      }
      if (inside) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (C == CP || c0 + i < C) *oc = fmaf(T, sm.bg[c0 + i], __uint_as_float(v[i]));
          oc += HW;
        }
      }
    }
    if (inside) {
      final_T[view * HW + pix] = T;
      n_contrib[view * HW + pix] = last_idx ? records[range.x + last_idx - 1].orig : 0u;
      if (out_depth) out_depth[view * HW + pix] = cross_idx ? records[range.x + cross_idx - 1].depth : 15.f;
      if (out_opacity) out_opacity[view * HW + pix] = 1.f - T;
    }
    const uint32_t my_max = __reduce_max_sync(0xffffffffu, last_idx);
    if (lane == 0 && my_max) atomicMax(&sm.max_contrib, my_max);
  } else if (warp == WORKER_WARPS) {
    // ------------------------------------------------ MMA issuer ------------------------------------------------
    constexpr uint32_t IDESC = idesc_tf32(128, CP);
    int b = 0;
    uint32_t started = 0;  // D holds a product (the first one overwrites, the later ones accumulate)
    uint32_t seen0 = 0, seen1 = 0, seen2 = 0;
    static_assert(STAGES == 3, "three stage counters");
    for (;; b++) {
      const int s = b % STAGES, sf = b % FST;
      const uint32_t use = (uint32_t)(b / STAGES), usef = (uint32_t)(b / FST);
      mbar_wait_wd(&sm.ready_f[sf], usef & 1);
      if (sm.cnt[sf] < 0) break;
      mbar_wait_wd(&sm.ready_w[s], use & 1);
      fence_after_sync();
      // (the counters only grow: what this use of the stage added is the difference to what the last use left)
      const uint32_t nz = *reinterpret_cast<volatile uint32_t*>(&sm.nonzero[s]);
      const uint32_t before = s == 0 ? seen0 : s == 1 ? seen1 : seen2;
      const bool any_w = nz != before;  // (warp-uniform)
      if (s == 0) seen0 = nz; else if (s == 1) seen1 = nz; else seen2 = nz;
      if (lane == 0) {
        if (any_w) {  // a block whose weights are all zero adds nothing: no product is issued for it
          const uint64_t d_raw = smem_desc(smem_u32(&sm.f[sf][0][0]), FwdSmem<CP>::LBO, FwdSmem<CP>::SBO);
          const uint64_t d_lo = smem_desc(smem_u32(&sm.f[sf][1][0]), FwdSmem<CP>::LBO, FwdSmem<CP>::SBO);
#pragma unroll
          for (int mb = 0; mb < MB; mb++) {
            const uint32_t d = tm + mb * CP;
            const uint32_t a = tm + A_COL0 + s * (MB * 16) + mb * 16;
            mma_ts_tf32(d, a, d_raw, IDESC, started);  // w . f
            mma_ts_tf32(d, a + KB, d_raw, IDESC, 1);    // w_lo . f
            mma_ts_tf32(d, a, d_lo, IDESC, 1);          // w . f_lo
          }
        }
        mma_commit(&sm.free_[s]);    // the weight columns of the stage
        mma_commit(&sm.free_f[sf]);  // the feature / record stage
      }
      started = started || any_w;
      __syncwarp();
    }
    if (lane == 0) {
      sm.started = started ? 1u : 0u;
      mbar_arrive(&sm.meta_);     // (a generic arrival: releases the flag to the workers)
      mma_commit(&sm.final_);     // every product issued so far has completed
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ loader ------------------------------------------------
    const int r = lane & 7, q = lane >> 3;
    const Record* src = records + range.x;
    const float* fbase = feats + (size_t)(view / views_per_sample) * P * C;
    auto load_ids = [&](int group) -> uint32_t {  // ids of records 32 group + lane
      const int j = group * 32 + lane;
      return j < n ? src[j].id : 0xffffffffu;
    };
    uint32_t idg_cur = load_ids(0), idg_next = load_ids(1);
    auto issue = [&](int bb) {  // gather block bb into the landing ring (always one commit group)
      if (bb > 0 && (bb & 3) == 0) {
        idg_cur = idg_next;
        idg_next = load_ids((bb >> 2) + 1);
      }
      const uint32_t id = __shfl_sync(0xffffffffu, idg_cur, (bb & 3) * 8 + r);
      if (bb < nblocks) {
        const int slot = bb % RING;
        if (id != 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < CP / 16; i++) {
            const int cc = q + 4 * i;
            if (4 * cc < C) cp_async16(&sm.land_f[slot][r][4 * cc], fbase + (size_t)id * C + 4 * cc);
          }
        }
        if (lane < 24 && bb * KB * 3 + lane < n * 3)
          cp_async16(reinterpret_cast<char*>(&sm.land_rec[slot][0]) + lane * 16,
                     reinterpret_cast<const char*>(src + bb * KB) + lane * 16);
      }
      cp_async_commit();
    };
    for (int bb = 0; bb < LOOKAHEAD; bb++) issue(bb);
    for (int b = 0;; b++) {
      const int s = b % FST;
      const uint32_t use = (uint32_t)(b / FST);
      issue(b + LOOKAHEAD);
      cp_async_wait<LOOKAHEAD>();  // block b has landed (this lane's own chunks)
      const bool stop = b >= nblocks || *reinterpret_cast<volatile uint32_t*>(&sm.done_warps) == WORKER_WARPS;
      mbar_wait_wd(&sm.free_f[s], (use & 1) ^ 1);
      if (stop) {
        if (lane == 0) {
          sm.cnt[s] = -1;
          mbar_arrive(&sm.ready_f[s]);
        }
        break;
      }
      const int slot = b % RING;
      const bool valid = b * KB + r < n;
#pragma unroll
      for (int i = 0; i < CP / 16; i++) {
        const int cc = q + 4 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && 4 * cc < C) v = *reinterpret_cast<const float4*>(&sm.land_f[slot][r][4 * cc]);
        // transposing store: element k of the chunk goes to channel row 4 cc + k.  Lanes that differ only in r >> 2 or
        // q >> 1 would hit the same bank (their targets are 128 / 256 bytes apart), so every lane starts its four
        // elements at a different k: rot = (r >> 2) + 2 (q >> 1) -> 32 distinct banks per store
        const int rot = (r >> 2) + 2 * (q >> 1);
        float e0 = v.x, e1 = v.y, e2 = v.z, e3 = v.w;
        if (rot & 1) { const float t0 = e0; e0 = e1; e1 = e2; e2 = e3; e3 = t0; }
        if (rot & 2) { const float t0 = e0, t1 = e1; e0 = e2; e1 = e3; e2 = t0; e3 = t1; }
        const float e[4] = {e0, e1, e2, e3};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int c = 4 * cc + ((k + rot) & 3);
          const uint32_t off = (uint32_t)((c & 7) * 16 + (c >> 3) * FwdSmem<CP>::SBO + (r & 3) * 4 + (r >> 2) * FwdSmem<CP>::LBO);
          *reinterpret_cast<float*>(&sm.f[s][0][off]) = e[k];
          *reinterpret_cast<float*>(&sm.f[s][1][off]) = tf32_lo(e[k]);
        }
      }
      if (lane < 24)
        *reinterpret_cast<float4*>(reinterpret_cast<char*>(&sm.rec[s][0]) + lane * 16) =
            *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(&sm.land_rec[slot][0]) + lane * 16);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        sm.cnt[s] = min(KB, n - b * KB);
        mbar_arrive(&sm.ready_f[s]);
      }
    }
    cp_async_wait<0>();
  }
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    if (MB == 2) max_contrib[(size_t)view * tiles_per_view + tile] = sm.max_contrib;
    else atomicMax(&max_contrib[(size_t)view * tiles_per_view + tile], sm.max_contrib);  // (zeroed before the launch)
  }
  if (warp == WORKER_WARPS) tmem_dealloc<FWD_TMEM_COLS>(tm);
}

template <int CP, int MB>
static int launch_forward_tc_cp(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                                const float* colors, const float* bg, float* fT, uint32_t* nc, uint32_t* mc,
                                float* out_color, float* out_depth, float* out_opacity) {
  // tensor memory allows 4 / MB CTAs per SM; asking for just over 1 / (4 / MB + 1) of the shared memory keeps one
  // more from becoming resident and spinning in tcgen05.alloc
  const size_t want = MB == 2 ? 80 * 1024 : 48 * 1024;
  const size_t dyn = sizeof(FwdSmem<CP>) > want ? sizeof(FwdSmem<CP>) : want;
  static int configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(render_forward_tc_kernel<CP, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return (int)e;
    configured[dev] = 1;
  }
  if (MB == 1) {  // the two halves of a tile combine their last contributor with atomicMax
    cudaError_t e = cudaMemsetAsync(mc, 0, (size_t)grid.x * grid.y * grid.z * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    grid.x *= 2;
  }
  OCRF_LAUNCH(render_forward_tc_kernel<CP, MB>, dim3(grid), dim3(MB * 128 + 64), dyn, st, sh->W, sh->H, sh->C, sh->P,
              sh->views_per_sample, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
  return 0;
}

bool forward_tc_supported(int C) {
  static const bool off = getenv("OCRF_TC") != nullptr && atoi(getenv("OCRF_TC")) == 0;
  return !off && C > 32 && C <= 80 && (C % 4) == 0;
}

int launch_forward_tc(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                      const float* colors, const float* bg, float* fT, uint32_t* nc, uint32_t* mc, float* out_color,
                      float* out_depth, float* out_opacity) {
  // whole tiles by default (half tiles, four CTAs per SM, measured 631 against 606 us at BASELINE config 4)
  static const int mb = (getenv("OCRF_TC_FWD_MB") != nullptr && atoi(getenv("OCRF_TC_FWD_MB")) == 1) ? 1 : 2;
#define OCRF_TCF(CPV)                                                                                                       \
  return mb == 2 ? launch_forward_tc_cp<CPV, 2>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity) \
                 : launch_forward_tc_cp<CPV, 1>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity)
  if (sh->C <= 48) OCRF_TCF(48);
  if (sh->C <= 64) OCRF_TCF(64);
  OCRF_TCF(80);
#undef OCRF_TCF
}

}  // namespace tc
}  // namespace ocrf
