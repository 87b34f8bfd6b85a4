// Stage 3: 16x16-tile forward alpha compositing of colour/features, median depth and opacity.
//
// Replaces renderCUDA<3> (cuda_rasterizer/forward.cu:261-374), adds the w-depth fork's median depth
// (/root/reference/diff-gaussian-rasterization-w-depth/README.md:8-13) and the opacity map 1 - T.
//
// Design
//   * grid = (tiles_x, tiles_y, V): every view of the batch in one launch.
//   * The tile's (already packed, contiguous) records are streamed into shared memory by the TMA
//     unit: one elected thread issues a 1-D cp.async.bulk per batch of 256 records into a two-stage
//     ring guarded by mbarriers, so the copy of batch r+1 runs under the blending of batch r and no
//     thread spends registers or LSU issue slots on staging.
//   * A warp owns a compact 8x4 (PPT=1) or 8x8 (PPT=2) pixel block rather than a 16x2 strip: fewer
//     Gaussians of the tile list reach any pixel of the warp, so more iterations are skipped by all
//     32 lanes at once.
//   * Every thread reads a record with three broadcast LDS.128; exp() is one MUFU.EX2.
//   * Early ray termination as the reference: the CTA leaves as soon as every pixel is done.
// The blend arithmetic keeps the reference's expressions and thresholds (power > 0, alpha < 1/255,
// min(0.99, .), stop-before-blend at T < 1e-4), see SURVEY.md appendix A4.
#include "common.cuh"

namespace ocrf {

constexpr int FWD_BATCH = 256;  // records per shared-memory stage

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int PPT>
__global__ void __launch_bounds__(TILE_PIX / PPT) render_forward_c3_scalar_kernel(
    int W, int H, const uint2* __restrict__ ranges /* culled lists */, const Record* __restrict__ records,
    const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    uint32_t* __restrict__ max_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_opacity) {
  pdl_enter();
  constexpr int NT = TILE_PIX / PPT;
  __shared__ __align__(128) Record s_rec[2][FWD_BATCH];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_max;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x, tiles_per_view = gridDim.x * gridDim.y;
  const int view = blockIdx.z;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const uint2 range = ranges[(size_t)view * tiles_per_view + tile];
  const int n = (int)(range.y - range.x);
  const int rounds = (n + FWD_BATCH - 1) / FWD_BATCH;
  const Record* src = records + range.x;

  // pixel block of this warp
  const int bx = (warp & 1) * 8 + (lane & 7);
  const int by = (warp >> 1) * (4 * PPT) + (lane >> 3);
  int px[PPT], py[PPT];
  bool inside[PPT], done[PPT];
  float fx[PPT], fy[PPT];
#pragma unroll
  for (int p = 0; p < PPT; p++) {
    px[p] = blockIdx.x * TILE + bx;
    py[p] = blockIdx.y * TILE + by + 4 * p;
    inside[p] = px[p] < W && py[p] < H;
    done[p] = !inside[p];
    fx[p] = (float)px[p];
    fy[p] = (float)py[p];
  }

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
    s_max = 0;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (r < rounds) {
        const uint32_t bytes = (uint32_t)min(FWD_BATCH, n - r * FWD_BATCH) * sizeof(Record);
        mbar_expect_tx(&s_bar[r], bytes);
        bulk_g2s(&s_rec[r][0], src + r * FWD_BATCH, bytes, &s_bar[r]);
      }
  }

  float T[PPT], D[PPT], C0[PPT], C1[PPT], C2[PPT];
  uint32_t last[PPT];   // reference numbering (position in the unculled list), stored as n_contrib
  uint32_t lastc = 0;   // position in the culled list of the last record any of my pixels blended
#pragma unroll
  for (int p = 0; p < PPT; p++) {
    T[p] = 1.f; D[p] = 15.f; C0[p] = C1[p] = C2[p] = 0.f; last[p] = 0;
  }

  for (int r = 0; r < rounds; r++) {
    const int st = r & 1;
    mbar_wait(&s_bar[st], (r >> 1) & 1);
    const int cnt = min(FWD_BATCH, n - r * FWD_BATCH);
    const float4* rec4 = reinterpret_cast<const float4*>(&s_rec[st][0]);
    // Control flow inside the batch is kept WARP-UNIFORM (votes over the full warp): with per-lane
    // `continue`s the compiler emits no reconvergence point at the loop head and the warp splits into
    // sub-groups that walk the list independently (measured: 9 of 32 lanes active, 3x the issue slots).
    bool all_done = true;
#pragma unroll
    for (int p = 0; p < PPT; p++) all_done = all_done && done[p];
    bool warp_done = __all_sync(0xffffffffu, all_done);
    if (!warp_done) {
      for (int j = 0; j < cnt; j++) {
        const float4 a = rec4[3 * j], b = rec4[3 * j + 1];
        bool any_blend = false;
        float alpha_p[PPT];
#pragma unroll
        for (int p = 0; p < PPT; p++) {
          const float dx = a.x - fx[p], dy = a.y - fy[p];
          const float power = a.z * dx * dx + b.x * dy * dy + a.w * dx * dy;  // log2 domain (scaled conic)
          const float alpha = fminf(0.99f, b.y * ex2_approx(power));
          const bool ok = !done[p] && power <= 0.0f && alpha >= 1.0f / 255.0f;
          alpha_p[p] = ok ? alpha : 0.f;
          any_blend = any_blend || ok;
        }
        if (!__any_sync(0xffffffffu, any_blend)) continue;  // nobody in the warp blends this Gaussian
        const float4 c = rec4[3 * j + 2];  // g, b, id, depth
        bool newly_done = false;
#pragma unroll
        for (int p = 0; p < PPT; p++) {
          const float alpha = alpha_p[p];
          const float test_T = T[p] * (1.f - alpha);
          const bool blend = alpha != 0.f && test_T >= 0.0001f;
          const bool stop = alpha != 0.f && test_T < 0.0001f;
          const float w = blend ? alpha * T[p] : 0.f;
          C0[p] = fmaf(b.w, w, C0[p]);
          C1[p] = fmaf(c.x, w, C1[p]);
          C2[p] = fmaf(c.y, w, C2[p]);
          if (blend && T[p] > 0.5f && test_T < 0.5f) D[p] = c.w;
          if (blend) {
            T[p] = test_T;
            last[p] = __float_as_uint(b.z);
            lastc = (uint32_t)(r * FWD_BATCH + j + 1);
          }
          done[p] = done[p] || stop;
          newly_done = newly_done || stop;
        }
        if (__any_sync(0xffffffffu, newly_done)) {
          all_done = true;
#pragma unroll
          for (int p = 0; p < PPT; p++) all_done = all_done && done[p];
          if (__all_sync(0xffffffffu, all_done)) break;
        }
      }
    }
    // everyone is finished with stage `st`; leave early once the whole tile is saturated
    const int num_done = __syncthreads_count(all_done);
    const bool quit = num_done == NT;
    if (quit) {
      if (r + 1 < rounds) mbar_wait(&s_bar[(r + 1) & 1], ((r + 1) >> 1) & 1);  // drain the copy in flight
      break;
    }
    if (tid == 0 && r + 2 < rounds) {
      const uint32_t bytes = (uint32_t)min(FWD_BATCH, n - (r + 2) * FWD_BATCH) * sizeof(Record);
      mbar_expect_tx(&s_bar[st], bytes);
      bulk_g2s(&s_rec[st][0], src + (r + 2) * FWD_BATCH, bytes, &s_bar[st]);
    }
  }

  uint32_t my_max = 0;
  const size_t HW = (size_t)H * W;
#pragma unroll
  for (int p = 0; p < PPT; p++) {
    if (!inside[p]) continue;
    const size_t pix = (size_t)py[p] * W + px[p];
    final_T[view * HW + pix] = T[p];
    n_contrib[view * HW + pix] = last[p];
    my_max = max(my_max, lastc);
    float* oc = out_color + (size_t)view * 3 * HW + pix;
    oc[0] = C0[p] + T[p] * bg[0];
    oc[HW] = C1[p] + T[p] * bg[1];
    oc[2 * HW] = C2[p] + T[p] * bg[2];
    if (out_depth) out_depth[view * HW + pix] = D[p];
    if (out_opacity) out_opacity[view * HW + pix] = 1.f - T[p];
  }
  my_max = __reduce_max_sync(0xffffffffu, my_max);
  if (lane == 0 && my_max) atomicMax(&s_max, my_max);
  __syncthreads();
  if (tid == 0) max_contrib[(size_t)view * tiles_per_view + tile] = s_max;
}

// The production C == 3 kernel: the scalar kernel above with the per-(pixel, Gaussian) arithmetic PACKED two pixels
// per instruction (sm_100 FFMA2 / FMUL2 / FADD2).  A thread owns PPT pixels of ONE column (same x, rows 4 apart), so
//   log2 G = qa dx^2 + qb dx dy + qc dy^2 = (qc dy + qb dx) dy + qa dx^2
// costs 1 FADD + 3 FMUL once per thread (the dx terms) and 1 FADD2 + 2 FFMA2 per pixel PAIR; the record's scalars
// enter the packed instructions as broadcast operands.  The blend itself (1 - alpha, T (1 - alpha), alpha T, the three
// colour accumulations) is packed the same way.  A finished (or out-of-image) pixel is marked by a NaN row coordinate:
// its power is NaN, `power <= 0` is false, and it needs no predicate of its own in the test.
template <int PPT>
__global__ void __launch_bounds__(TILE_PIX / PPT) render_forward_c3_kernel(
    int W, int H, const uint2* __restrict__ ranges /* culled lists */, const Record* __restrict__ records,
    const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    uint32_t* __restrict__ max_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_opacity) {
  pdl_enter();
  static_assert(PPT == 2 || PPT == 4, "pixel pairs");
  constexpr int NT = TILE_PIX / PPT;
  constexpr int NQ = PPT / 2;
  __shared__ __align__(128) Record s_rec[2][FWD_BATCH];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_max;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x, tiles_per_view = gridDim.x * gridDim.y;
  const int view = blockIdx.z;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const uint2 range = ranges[(size_t)view * tiles_per_view + tile];
  const int n = (int)(range.y - range.x);
  const int rounds = (n + FWD_BATCH - 1) / FWD_BATCH;
  const Record* src = records + range.x;

  // pixel column of this thread: x = px, rows py0 + 4 p
  const int px = blockIdx.x * TILE + (warp & 1) * 8 + (lane & 7);
  const int py0 = blockIdx.y * TILE + (warp >> 1) * (4 * PPT) + (lane >> 3);
  const float fx = (float)px;
  float2 nfy[NQ];  // minus the row coordinate of the pair's pixels; NaN = pixel finished / outside the image
  const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int y0 = py0 + 8 * q, y1 = y0 + 4;
    nfy[q].x = (px < W && y0 < H) ? -(float)y0 : qnan;
    nfy[q].y = (px < W && y1 < H) ? -(float)y1 : qnan;
  }

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
    s_max = 0;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (r < rounds) {
        const uint32_t bytes = (uint32_t)min(FWD_BATCH, n - r * FWD_BATCH) * sizeof(Record);
        mbar_expect_tx(&s_bar[r], bytes);
        bulk_g2s(&s_rec[r][0], src + r * FWD_BATCH, bytes, &s_bar[r]);
      }
  }

  float2 T[NQ], C0[NQ], C1[NQ], C2[NQ], D[NQ];
  uint32_t last[PPT];   // reference numbering (position in the unculled list), stored as n_contrib
  uint32_t lastc = 0;   // position in the culled list of the last record any of my pixels blended
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    T[q] = make_float2(1.f, 1.f);
    D[q] = make_float2(15.f, 15.f);
    C0[q] = C1[q] = C2[q] = make_float2(0.f, 0.f);
    last[2 * q] = last[2 * q + 1] = 0;
  }
  auto all_finished = [&]() {
    bool f = true;
#pragma unroll
    for (int q = 0; q < NQ; q++) f = f && (nfy[q].x != nfy[q].x) && (nfy[q].y != nfy[q].y);
    return f;
  };

  for (int r = 0; r < rounds; r++) {
    const int st = r & 1;
    mbar_wait(&s_bar[st], (r >> 1) & 1);
    const int cnt = min(FWD_BATCH, n - r * FWD_BATCH);
    const float4* rec4 = reinterpret_cast<const float4*>(&s_rec[st][0]);
    // Control flow inside the batch is WARP-UNIFORM (votes over the full warp), see the scalar kernel.
    bool all_done = all_finished();
    if (!__all_sync(0xffffffffu, all_done)) {
      for (int j = 0; j < cnt; j++) {
        const float4 a = rec4[3 * j], b = rec4[3 * j + 1];  // x, y, qa, qb | qc, op, orig, r
        const float dx = a.x - fx;
        const float u = a.z * dx * dx;  // qa dx^2
        const float w = a.w * dx;       // qb dx
        float2 araw[NQ];  // opacity * G, not yet clamped: the clamp and the zeroing of non-participants wait for the blend
        bool ok[PPT];
        bool any_blend = false;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const float2 dy = fadd2(bcast2(a.y), nfy[q]);
          const float2 t = ffma2(bcast2(b.x), dy, bcast2(w));
          const float2 pw = ffma2(t, dy, bcast2(u));
          const float2 g = make_float2(ex2_approx(pw.x), ex2_approx(pw.y));
          araw[q] = fmul2(bcast2(b.y), g);
          // min(0.99, alpha) >= 1/255 <=> alpha >= 1/255; the NaN power of a finished pixel compares false
          ok[2 * q] = pw.x <= 0.0f && araw[q].x >= 1.0f / 255.0f;
          ok[2 * q + 1] = pw.y <= 0.0f && araw[q].y >= 1.0f / 255.0f;
          any_blend = any_blend || ok[2 * q] || ok[2 * q + 1];
        }
        if (!__any_sync(0xffffffffu, any_blend)) continue;  // nobody in the warp blends this Gaussian
        const float4 c = rec4[3 * j + 2];  // g, b, id, depth
        bool newly_done = false;
        bool blended = false;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const bool on0 = ok[2 * q], on1 = ok[2 * q + 1];
          float2 alq;
          alq.x = on0 ? fminf(0.99f, araw[q].x) : 0.f;
          alq.y = on1 ? fminf(0.99f, araw[q].y) : 0.f;
          const float2 om = fadd2(bcast2(1.f), make_float2(-alq.x, -alq.y));
          const float2 tT = fmul2(T[q], om);
          float2 wgt = fmul2(alq, T[q]);  // 0 for a pixel that does not take part (alpha == 0)
          const bool stop0 = on0 && tT.x < 0.0001f, stop1 = on1 && tT.y < 0.0001f;
          const bool bl0 = on0 && !stop0, bl1 = on1 && !stop1;
          wgt.x = stop0 ? 0.f : wgt.x;
          wgt.y = stop1 ? 0.f : wgt.y;
          C0[q] = ffma2(bcast2(b.w), wgt, C0[q]);
          C1[q] = ffma2(bcast2(c.x), wgt, C1[q]);
          C2[q] = ffma2(bcast2(c.y), wgt, C2[q]);
          if (bl0 && T[q].x > 0.5f && tT.x < 0.5f) D[q].x = c.w;
          if (bl1 && T[q].y > 0.5f && tT.y < 0.5f) D[q].y = c.w;
          T[q].x = bl0 ? tT.x : T[q].x;
          T[q].y = bl1 ? tT.y : T[q].y;
          last[2 * q] = bl0 ? __float_as_uint(b.z) : last[2 * q];
          last[2 * q + 1] = bl1 ? __float_as_uint(b.z) : last[2 * q + 1];
          blended = blended || bl0 || bl1;
          nfy[q].x = stop0 ? qnan : nfy[q].x;
          nfy[q].y = stop1 ? qnan : nfy[q].y;
          newly_done = newly_done || stop0 || stop1;
        }
        if (blended) lastc = (uint32_t)(r * FWD_BATCH + j + 1);
        if (__any_sync(0xffffffffu, newly_done)) {
          all_done = all_finished();
          if (__all_sync(0xffffffffu, all_done)) break;
        }
      }
    }
    // everyone is finished with stage `st`; leave early once the whole tile is saturated
    const int num_done = __syncthreads_count(all_done);
    if (num_done == NT) {
      if (r + 1 < rounds) mbar_wait(&s_bar[(r + 1) & 1], ((r + 1) >> 1) & 1);  // drain the copy in flight
      break;
    }
    if (tid == 0 && r + 2 < rounds) {
      const uint32_t bytes = (uint32_t)min(FWD_BATCH, n - (r + 2) * FWD_BATCH) * sizeof(Record);
      mbar_expect_tx(&s_bar[st], bytes);
      bulk_g2s(&s_rec[st][0], src + (r + 2) * FWD_BATCH, bytes, &s_bar[st]);
    }
  }

  const size_t HW = (size_t)H * W;
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
#pragma unroll
  for (int p = 0; p < PPT; p++) {
    const int q = p >> 1;
    const int py = py0 + 8 * q + 4 * (p & 1);
    if (!(px < W && py < H)) continue;
    const float Tp = (p & 1) ? T[q].y : T[q].x;
    const size_t pix = (size_t)py * W + px;
    final_T[view * HW + pix] = Tp;
    n_contrib[view * HW + pix] = last[p];
    float* oc = out_color + (size_t)view * 3 * HW + pix;
    oc[0] = ((p & 1) ? C0[q].y : C0[q].x) + Tp * bg0;
    oc[HW] = ((p & 1) ? C1[q].y : C1[q].x) + Tp * bg1;
    oc[2 * HW] = ((p & 1) ? C2[q].y : C2[q].x) + Tp * bg2;
    if (out_depth) out_depth[view * HW + pix] = (p & 1) ? D[q].y : D[q].x;
    if (out_opacity) out_opacity[view * HW + pix] = 1.f - Tp;
  }
  const uint32_t my_max = __reduce_max_sync(0xffffffffu, lastc);
  if (lane == 0 && my_max) atomicMax(&s_max, my_max);
  __syncthreads();
  if (tid == 0) max_contrib[(size_t)view * tiles_per_view + tile] = s_max;
}

// Generic channel count: 32-byte records, features gathered by id, CK channels per traversal.
constexpr int FWD_CK = 16;
constexpr int FWDG_BATCH = 128;

__global__ void __launch_bounds__(TILE_PIX) render_forward_generic_kernel(
    int W, int H, int C, int P, int views_per_sample, const uint2* __restrict__ ranges,
    const Record* __restrict__ records, const float* __restrict__ feats, const float* __restrict__ bg,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ max_contrib,
    float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_opacity) {
  __shared__ __align__(16) Record s_rec[FWDG_BATCH];
  __shared__ __align__(16) float s_feat[FWDG_BATCH][FWD_CK];
  __shared__ uint32_t s_max;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x, tiles_per_view = gridDim.x * gridDim.y;
  const int view = blockIdx.z;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const uint2 range = ranges[(size_t)view * tiles_per_view + tile];
  const int n = (int)(range.y - range.x);
  const int rounds = (n + FWDG_BATCH - 1) / FWDG_BATCH;
  const float* fbase = feats + (size_t)(view / views_per_sample) * P * C;

  const int px = blockIdx.x * TILE + (warp & 1) * 8 + (lane & 7);
  const int py = blockIdx.y * TILE + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float fx = (float)px, fy = (float)py;
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)py * W + px;
  if (tid == 0) s_max = 0;

  for (int c0 = 0; c0 < C; c0 += FWD_CK) {
    const int ck = min(FWD_CK, C - c0);
    bool done = !inside;
    float T = 1.f, D = 15.f;
    uint32_t last = 0, lastc = 0;
    float acc[FWD_CK];
#pragma unroll
    for (int k = 0; k < FWD_CK; k++) acc[k] = 0.f;
    for (int r = 0; r < rounds; r++) {
      const int num_done = __syncthreads_count(done);
      if (num_done == TILE_PIX) break;
      const int cnt = min(FWDG_BATCH, n - r * FWDG_BATCH);
      if (tid < cnt) s_rec[tid] = records[range.x + r * FWDG_BATCH + tid];
      for (int e = tid; e < cnt * FWD_CK; e += TILE_PIX) {
        const int j = e / FWD_CK, k = e - j * FWD_CK;
        const uint32_t id = records[range.x + r * FWDG_BATCH + j].id;
        s_feat[j][k] = k < ck ? __ldg(fbase + (size_t)id * C + c0 + k) : 0.f;
      }
      __syncthreads();
      if (__all_sync(0xffffffffu, done)) continue;
      for (int j = 0; j < cnt; j++) {  // warp-uniform control flow, see render_forward_c3_kernel
        const float4 a = reinterpret_cast<const float4*>(&s_rec[j])[0];
        const float4 b = reinterpret_cast<const float4*>(&s_rec[j])[1];
        const float dx = a.x - fx, dy = a.y - fy;
        const float power = a.z * dx * dx + b.x * dy * dy + a.w * dx * dy;  // log2 domain (scaled conic)
        const float alpha = fminf(0.99f, b.y * ex2_approx(power));
        const bool ok = !done && power <= 0.0f && alpha >= 1.0f / 255.0f;
        if (!__any_sync(0xffffffffu, ok)) continue;
        const float test_T = T * (1.f - alpha);
        const bool blend = ok && test_T >= 0.0001f;
        const bool stop = ok && test_T < 0.0001f;
        const float w = blend ? alpha * T : 0.f;
#pragma unroll
        for (int k = 0; k < FWD_CK; k += 4) {
          const float4 f = *reinterpret_cast<const float4*>(&s_feat[j][k]);
          acc[k] = fmaf(f.x, w, acc[k]); acc[k + 1] = fmaf(f.y, w, acc[k + 1]);
          acc[k + 2] = fmaf(f.z, w, acc[k + 2]); acc[k + 3] = fmaf(f.w, w, acc[k + 3]);
        }
        if (blend) {
          if (T > 0.5f && test_T < 0.5f) D = s_rec[j].depth;
          T = test_T;
          last = s_rec[j].orig;
          lastc = (uint32_t)(r * FWDG_BATCH + j + 1);
        }
        done = done || stop;
        if (__any_sync(0xffffffffu, stop) && __all_sync(0xffffffffu, done)) break;
      }
    }
    if (inside) {
#pragma unroll
      for (int k = 0; k < FWD_CK; k++)
        if (k < ck) out_color[((size_t)view * C + c0 + k) * HW + pix] = acc[k] + T * bg[c0 + k];
      if (c0 == 0) {
        final_T[view * HW + pix] = T;
        n_contrib[view * HW + pix] = last;
        if (out_depth) out_depth[view * HW + pix] = D;
        if (out_opacity) out_opacity[view * HW + pix] = 1.f - T;
        atomicMax(&s_max, lastc);
      }
    }
    __syncthreads();
  }
  if (tid == 0) max_contrib[(size_t)view * tiles_per_view + tile] = s_max;
}


// Generic channel count, single traversal (C <= 96): all CP channel accumulators of a pixel live in registers, so the
// list is walked -- and every alpha evaluated -- once instead of once per 16-channel chunk (C = 80: 5 traversals).
template <int CP>
__global__ void __launch_bounds__(TILE_PIX, CP <= 80 ? 2 : 1) render_forward_wide_kernel(
    int W, int H, int C, int P, int views_per_sample, const uint2* __restrict__ ranges,
    const Record* __restrict__ records, const float* __restrict__ feats, const float* __restrict__ bg,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ max_contrib,
    float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_opacity) {
  extern __shared__ __align__(16) unsigned char smem_w[];
  Record* s_rec = reinterpret_cast<Record*>(smem_w);                                  // [FWDG_BATCH]
  float* s_feat = reinterpret_cast<float*>(smem_w + FWDG_BATCH * sizeof(Record));     // [FWDG_BATCH][CP]
  __shared__ uint32_t s_max;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x, tiles_per_view = gridDim.x * gridDim.y;
  const int view = blockIdx.z;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const uint2 range = ranges[(size_t)view * tiles_per_view + tile];
  const int n = (int)(range.y - range.x);
  const int rounds = (n + FWDG_BATCH - 1) / FWDG_BATCH;
  const float* fbase = feats + (size_t)(view / views_per_sample) * P * C;

  const int px = blockIdx.x * TILE + (warp & 1) * 8 + (lane & 7);
  const int py = blockIdx.y * TILE + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float fx = (float)px, fy = (float)py;
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)py * W + px;
  if (tid == 0) s_max = 0;

  bool done = !inside;
  float T = 1.f, D = 15.f;
  uint32_t last = 0, lastc = 0;
  float acc[CP];
#pragma unroll
  for (int k = 0; k < CP; k++) acc[k] = 0.f;
  for (int r = 0; r < rounds; r++) {
    const int num_done = __syncthreads_count(done);
    if (num_done == TILE_PIX) break;
    const int cnt = min(FWDG_BATCH, n - r * FWDG_BATCH);
    if (tid < cnt) s_rec[tid] = records[range.x + r * FWDG_BATCH + tid];
    for (int e = tid; e < cnt * (CP / 4); e += TILE_PIX) {  // feature rows, 16 bytes per thread
      const int j = e / (CP / 4), k4 = (e - j * (CP / 4)) * 4;
      const uint32_t id = records[range.x + r * FWDG_BATCH + j].id;
      float4 f;
      f.x = k4 + 0 < C ? __ldg(fbase + (size_t)id * C + k4 + 0) : 0.f;
      f.y = k4 + 1 < C ? __ldg(fbase + (size_t)id * C + k4 + 1) : 0.f;
      f.z = k4 + 2 < C ? __ldg(fbase + (size_t)id * C + k4 + 2) : 0.f;
      f.w = k4 + 3 < C ? __ldg(fbase + (size_t)id * C + k4 + 3) : 0.f;
      *reinterpret_cast<float4*>(s_feat + j * CP + k4) = f;
    }
    __syncthreads();
    if (__all_sync(0xffffffffu, done)) continue;
    for (int j = 0; j < cnt; j++) {  // warp-uniform control flow, see render_forward_c3_kernel
      const float4 a = reinterpret_cast<const float4*>(&s_rec[j])[0];
      const float4 b = reinterpret_cast<const float4*>(&s_rec[j])[1];
      const float dx = a.x - fx, dy = a.y - fy;
      const float power = a.z * dx * dx + b.x * dy * dy + a.w * dx * dy;  // log2 domain (scaled conic)
      const float alpha = fminf(0.99f, b.y * ex2_approx(power));
      const bool ok = !done && power <= 0.0f && alpha >= 1.0f / 255.0f;
      if (!__any_sync(0xffffffffu, ok)) continue;
      const float test_T = T * (1.f - alpha);
      const bool blend = ok && test_T >= 0.0001f;
      const bool stop = ok && test_T < 0.0001f;
      const float w = blend ? alpha * T : 0.f;
      const float* fj = s_feat + j * CP;
      const float2 w2 = bcast2(w);
#pragma unroll
      for (int k = 0; k < CP; k += 4) {  // two channels per FFMA2: the feature pair comes packed out of the LDS.128
        const float4 f = *reinterpret_cast<const float4*>(fj + k);
        const float2 a01 = ffma2(make_float2(f.x, f.y), w2, make_float2(acc[k], acc[k + 1]));
        const float2 a23 = ffma2(make_float2(f.z, f.w), w2, make_float2(acc[k + 2], acc[k + 3]));
        acc[k] = a01.x; acc[k + 1] = a01.y; acc[k + 2] = a23.x; acc[k + 3] = a23.y;
      }
      if (blend) {
        if (T > 0.5f && test_T < 0.5f) D = s_rec[j].depth;
        T = test_T;
        last = s_rec[j].orig;
        lastc = (uint32_t)(r * FWDG_BATCH + j + 1);
      }
      done = done || stop;
      if (__any_sync(0xffffffffu, stop) && __all_sync(0xffffffffu, done)) break;
    }
  }
  if (inside) {
#pragma unroll
    for (int k = 0; k < CP; k++)
      if (k < C) out_color[((size_t)view * C + k) * HW + pix] = acc[k] + T * bg[k];
    final_T[view * HW + pix] = T;
    n_contrib[view * HW + pix] = last;
    if (out_depth) out_depth[view * HW + pix] = D;
    if (out_opacity) out_opacity[view * HW + pix] = 1.f - T;
    atomicMax(&s_max, lastc);
  }
  __syncthreads();
  if (tid == 0) max_contrib[(size_t)view * tiles_per_view + tile] = s_max;
}

template <int CP>
static int launch_forward_wide(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                               const float* colors, const float* bg, float* fT, uint32_t* nc, uint32_t* mc,
                               float* out_color, float* out_depth, float* out_opacity) {
  const size_t dyn = FWDG_BATCH * sizeof(Record) + (size_t)FWDG_BATCH * CP * 4;
  cudaError_t e = cudaFuncSetAttribute(render_forward_wide_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return (int)e;
  render_forward_wide_kernel<CP><<<grid, TILE_PIX, dyn, st>>>(sh->W, sh->H, sh->C, sh->P, sh->views_per_sample, ranges, rec,
                                                               colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
  return 0;
}

}  // namespace ocrf

using namespace ocrf;

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

extern "C" int ocrf_render_forward(void* stream, const OcrfShape* sh, uint64_t pair_capacity, const float* colors,
                                   int use_sh, const float* bg, const void* geom_ws, const void* bin_ws,
                                   void* image_ws, float* out_color, float* out_depth, float* out_opacity) {
  if (!sh || !bg || !bin_ws || !image_ws || !out_color) return OCRF_EINVAL;
  if (sh->C <= 0 || (sh->C != 3 && !colors)) return OCRF_EINVAL;
  (void)geom_ws;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OcrfBinLayout B;
  OcrfImageLayout I;
  int rc = ocrf_bin_layout(sh, pair_capacity, &B);
  if (rc) return rc;
  ocrf_image_layout(sh, &I);
  const dim3 grid(tiles_x(*sh), tiles_y(*sh), sh->V);
  const uint2* ranges = at<uint2>(image_ws, I.ranges_render);
  float* fT = at<float>(image_ws, I.final_T);
  uint32_t* nc = at<uint32_t>(image_ws, I.n_contrib);
  uint32_t* mc = at<uint32_t>(image_ws, I.max_contrib);
  if (sh->C == 3) {
    static const int ppt = env_int("OCRF_FWD_PPT", 2);
    static const int packed = env_int("OCRF_FWD_PACKED", 1);  // 0: the scalar kernel (A/B measurements)
    const Record* rec = at<Record>(bin_ws, B.records);
#define OCRF_FWD_ARGS sh->W, sh->H, ranges, rec, bg, fT, nc, mc, out_color, out_depth, out_opacity
    if (packed && ppt == 4) OCRF_LAUNCH(render_forward_c3_kernel<4>, dim3(grid), dim3(TILE_PIX / 4), 0, st, OCRF_FWD_ARGS);
    else if (packed) OCRF_LAUNCH(render_forward_c3_kernel<2>, dim3(grid), dim3(TILE_PIX / 2), 0, st, OCRF_FWD_ARGS);
    else if (ppt == 4) OCRF_LAUNCH(render_forward_c3_scalar_kernel<4>, dim3(grid), dim3(TILE_PIX / 4), 0, st, OCRF_FWD_ARGS);
    else if (ppt == 2) OCRF_LAUNCH(render_forward_c3_scalar_kernel<2>, dim3(grid), dim3(TILE_PIX / 2), 0, st, OCRF_FWD_ARGS);
    else OCRF_LAUNCH(render_forward_c3_scalar_kernel<1>, dim3(grid), dim3(TILE_PIX), 0, st, OCRF_FWD_ARGS);
#undef OCRF_FWD_ARGS
  } else {
    (void)use_sh;
    const Record* rec = at<Record>(bin_ws, B.records);
    int rcw = 0;
    if (tc::forward_tc_supported(sh->C)) rcw = tc::launch_forward_tc(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    else if (sh->C <= 16) rcw = launch_forward_wide<16>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    else if (sh->C <= 32) rcw = launch_forward_wide<32>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    else if (sh->C <= 48) rcw = launch_forward_wide<48>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    else if (sh->C <= 64) rcw = launch_forward_wide<64>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    else if (sh->C <= 80) rcw = launch_forward_wide<80>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    else if (sh->C <= 96) rcw = launch_forward_wide<96>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    else  // wider than the register-resident variants: 16 channels per traversal
      render_forward_generic_kernel<<<grid, TILE_PIX, 0, st>>>(sh->W, sh->H, sh->C, sh->P, sh->views_per_sample, ranges, rec,
                                                               colors, bg, fT, nc, mc, out_color, out_depth, out_opacity);
    if (rcw) return rcw;
  }
  OCRF_CHECK_LAST();
  return 0;
}
