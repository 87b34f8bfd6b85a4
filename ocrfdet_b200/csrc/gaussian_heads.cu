// Scope row a12 ("next" f-1): OcRF Gaussian construction as ONE pass over the voxel features.
//
// Replaces the four tiny MLP heads of the OcRF head
// (/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:272-320, called at :1130-1133):
//     opacity  = sigmoid (A.fc2(relu(A.fc1(feat))))                 [n,1]
//     scaling  = softplus(S.fc2(relu(S.fc1(feat))))                 [n,3]
//     rotation = normalize(R.fc2(relu(R.fc1(feat))))                [n,4]
//     color    = sigmoid (C.fc2(relu(C.fc1(cat(feat, rgb)))))       [n,3]
// The reference reads voxel_feat [212992, 80] (68 MB per sample) four times and launches ~16 small
// kernels; here the 16 hidden units of all four heads are one [n, F+3] x [F+3, 16] product evaluated
// while the feature tile sits in shared memory, so the features are read from HBM exactly once
// (forward) / once more (backward).  Algorithmic bytes per Gaussian (F = 80): forward 332 read + 108 written
// (44 outputs + 64 hidden kept for backward); backward 332 + 64 + 44 read, 320 written.
//
// Data movement: every 128-Gaussian tile is fetched with one 1-D bulk copy per row (TMA unit, SASS UBLKCP) into
// rows padded to an odd number of 16-byte words, so that "thread t reads row t" with LDS.128 is bank-conflict
// free; backward sends dL/dfeat back with bulk stores from the same rows.  Each thread owns TWO rows, so one
// broadcast LDS.128 of weights feeds 8 FFMAs: the inner loop is FP32-issue bound, not shared-memory bound.
//
// Packed parameters (built by ocrfdet_b200/gaussian_heads.py from the reference's parameter tensors):
//   w1t [F+3][16]  input-major: columns 0-3 S.fc1, 4-7 R.fc1, 8-11 A.fc1 (their 3 rgb rows zero), 12-15 C.fc1;  b1 [16]
//   w2  [11][4]    rows 0-2 S.fc2, 3-6 R.fc2, 7 A.fc2, 8-10 C.fc2;                                               b2 [11]
#include "common.cuh"

namespace ocrf {

constexpr int GH_ROWS = 128;     // Gaussians per tile
constexpr int GH_THREADS = 64;   // two rows per thread
constexpr int GH_HID = 16;
constexpr int GH_OUT = 11;
constexpr int GH_MAXK = 128;     // F + 3 <= 128: every input column has an owner thread slot in backward

__host__ __device__ inline int gh_row_stride(int F) {  // floats; [feat F | rgb 3 | pad], an odd number of float4s
  int f4 = (F + 3 + 3) / 4;
  if ((f4 & 1) == 0) f4++;
  return 4 * f4;
}

__device__ __forceinline__ int gh_hidden_base(int i) { return i < 3 ? 0 : (i < 7 ? 4 : (i < 8 ? 8 : 12)); }

__device__ __forceinline__ void gh_second_layer(const float* h, const float* s_w2, const float* s_b2, float* z) {
  // S rows 0-2 use hidden 0-3, R rows 3-6 hidden 4-7, A row 7 hidden 8-11, C rows 8-10 hidden 12-15
#pragma unroll
  for (int i = 0; i < GH_OUT; i++) {
    const int hb = gh_hidden_base(i);
    float acc = s_b2[i];
#pragma unroll
    for (int j = 0; j < 4; j++) acc = fmaf(s_w2[i * 4 + j], h[hb + j], acc);
    z[i] = acc;
  }
}

__device__ __forceinline__ float gh_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float gh_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch: beta 1, threshold 20

// Fetch one tile: rows [base, base+rows) of feat -> s_x rows (stride XS); rgb -> columns F..F+2.
// fast: feat 16-byte aligned and F % 4 == 0 (bulk copies, completion on `bar`); otherwise plain loads.
__device__ __forceinline__ void gh_fetch_tile(const float* __restrict__ feat, const float* __restrict__ rgb,
                                              long long base, int rows, int F, int XS, float* s_x, uint64_t* bar,
                                              bool fast) {
  const int tid = threadIdx.x;
  if (fast) {
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)rows * F * 4u);
#pragma unroll
    for (int r = tid; r < GH_ROWS; r += GH_THREADS)
      if (r < rows) bulk_g2s(s_x + r * XS, feat + (base + r) * F, (uint32_t)F * 4u, bar);
  } else {
    for (int r = 0; r < rows; r++)
      for (int k = tid; k < F; k += GH_THREADS) s_x[r * XS + k] = __ldg(feat + (base + r) * F + k);
  }
#pragma unroll
  for (int r = tid; r < GH_ROWS; r += GH_THREADS) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (r < rows) {
      c0 = __ldg(rgb + (base + r) * 3 + 0);
      c1 = __ldg(rgb + (base + r) * 3 + 1);
      c2 = __ldg(rgb + (base + r) * 3 + 2);
    }
    s_x[r * XS + F + 0] = c0;
    s_x[r * XS + F + 1] = c1;
    s_x[r * XS + F + 2] = c2;
  }
}

#define GH_FMA16(acc, XV, WA, WB, WC, WD)                                                                \
  acc[0] = fmaf(XV, WA.x, acc[0]); acc[1] = fmaf(XV, WA.y, acc[1]); acc[2] = fmaf(XV, WA.z, acc[2]);       \
  acc[3] = fmaf(XV, WA.w, acc[3]); acc[4] = fmaf(XV, WB.x, acc[4]); acc[5] = fmaf(XV, WB.y, acc[5]);       \
  acc[6] = fmaf(XV, WB.z, acc[6]); acc[7] = fmaf(XV, WB.w, acc[7]); acc[8] = fmaf(XV, WC.x, acc[8]);       \
  acc[9] = fmaf(XV, WC.y, acc[9]); acc[10] = fmaf(XV, WC.z, acc[10]); acc[11] = fmaf(XV, WC.w, acc[11]);   \
  acc[12] = fmaf(XV, WD.x, acc[12]); acc[13] = fmaf(XV, WD.y, acc[13]); acc[14] = fmaf(XV, WD.z, acc[14]); \
  acc[15] = fmaf(XV, WD.w, acc[15]);

__global__ void __launch_bounds__(GH_THREADS) gaussian_heads_forward_kernel(
    long long n, int F, int fast, const float* __restrict__ feat, const float* __restrict__ rgb,
    const float* __restrict__ w1t, const float* __restrict__ b1, const float* __restrict__ w2,
    const float* __restrict__ b2, float* __restrict__ opacity, float* __restrict__ scales,
    float* __restrict__ rotations, float* __restrict__ colors, float* __restrict__ hidden) {
  extern __shared__ __align__(128) float s_mem[];
  __shared__ __align__(8) uint64_t s_bar;
  const int XS = gh_row_stride(F);
  const int K = F + 3;
  float* s_x = s_mem;                       // [GH_ROWS][XS]
  float* s_w1 = s_x + GH_ROWS * XS;         // [K][16]
  float* s_b1 = s_w1 + K * GH_HID;          // [16]
  float* s_w2 = s_b1 + GH_HID;              // [11][4]
  float* s_b2 = s_w2 + GH_OUT * 4;          // [11]
  const int tid = threadIdx.x;
  for (int i = tid; i < K * GH_HID; i += GH_THREADS) s_w1[i] = w1t[i];
  if (tid < GH_HID) s_b1[tid] = b1[tid];
  if (tid < GH_OUT * 4) s_w2[tid] = w2[tid];
  if (tid < GH_OUT) s_b2[tid] = b2[tid];
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const long long tiles = (n + GH_ROWS - 1) / GH_ROWS;
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long base = tile * GH_ROWS;
    const int rows = (int)min((long long)GH_ROWS, n - base);
    gh_fetch_tile(feat, rgb, base, rows, F, XS, s_x, &s_bar, fast != 0);
    if (fast) {
      mbar_wait(&s_bar, phase);
      phase ^= 1;
    }
    __syncthreads();  // rgb columns (and the slow path's rows) are written by other threads

    float h0[GH_HID], h1[GH_HID];
#pragma unroll
    for (int j = 0; j < GH_HID; j++) h0[j] = h1[j] = s_b1[j];
    const float* x0p = s_x + tid * XS;
    const float* x1p = s_x + (tid + GH_THREADS) * XS;
    const int k4n = K >> 2;
    for (int k4 = 0; k4 < k4n; k4++) {
      const float4 xa = *reinterpret_cast<const float4*>(x0p + 4 * k4);
      const float4 xb = *reinterpret_cast<const float4*>(x1p + 4 * k4);
      const float xav[4] = {xa.x, xa.y, xa.z, xa.w};
      const float xbv[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4* w = reinterpret_cast<const float4*>(s_w1 + (4 * k4 + q) * GH_HID);
        const float4 wa = w[0], wb = w[1], wc = w[2], wd = w[3];
        GH_FMA16(h0, xav[q], wa, wb, wc, wd)
        GH_FMA16(h1, xbv[q], wa, wb, wc, wd)
      }
    }
    for (int k = k4n << 2; k < K; k++) {
      const float xa = x0p[k], xb = x1p[k];
      const float4* w = reinterpret_cast<const float4*>(s_w1 + k * GH_HID);
      const float4 wa = w[0], wb = w[1], wc = w[2], wd = w[3];
      GH_FMA16(h0, xa, wa, wb, wc, wd)
      GH_FMA16(h1, xb, wa, wb, wc, wd)
    }
#pragma unroll
    for (int half = 0; half < 2; half++) {
      float* h = half ? h1 : h0;
      const int r = tid + half * GH_THREADS;
      if (r < rows) {  // rows past the end of a partial tile hold stale shared memory: computed, never stored
        const long long g = base + r;
#pragma unroll
        for (int j = 0; j < GH_HID; j++) h[j] = fmaxf(h[j], 0.f);
        float z[GH_OUT];
        gh_second_layer(h, s_w2, s_b2, z);
        scales[g * 3 + 0] = gh_softplus(z[0]);
        scales[g * 3 + 1] = gh_softplus(z[1]);
        scales[g * 3 + 2] = gh_softplus(z[2]);
        const float nrm = fmaxf(sqrtf(z[3] * z[3] + z[4] * z[4] + z[5] * z[5] + z[6] * z[6]), 1e-12f);  // F.normalize eps
        reinterpret_cast<float4*>(rotations)[g] = make_float4(z[3] / nrm, z[4] / nrm, z[5] / nrm, z[6] / nrm);
        opacity[g] = gh_sigmoid(z[7]);
        colors[g * 3 + 0] = gh_sigmoid(z[8]);
        colors[g * 3 + 1] = gh_sigmoid(z[9]);
        colors[g * 3 + 2] = gh_sigmoid(z[10]);
        float4* hp = reinterpret_cast<float4*>(hidden + g * GH_HID);
        hp[0] = make_float4(h[0], h[1], h[2], h[3]);
        hp[1] = make_float4(h[4], h[5], h[6], h[7]);
        hp[2] = make_float4(h[8], h[9], h[10], h[11]);
        hp[3] = make_float4(h[12], h[13], h[14], h[15]);
      }
    }
    __syncthreads();  // every row has been consumed before the next tile lands in s_x
  }
}

// ---- forward, fast path (F % 4 == 0, 16-byte aligned features): k-sliced cp.async ring ----------------------
// A 128-row tile is consumed in slices of 16 feature columns.  One slice of a tile is 128 x 64 B = 8 KB (10 KB with the
// conflict-free 20-float row pitch), so a 3-deep ring costs 30 KB instead of the 43 KB a whole tile needs: six CTAs
// (12 warps) fit per SM, every CTA prefetches two slices ahead, and HBM latency is hidden inside the CTA instead of
// by neighbours that all stall at the same moment.  Each thread copies eight 16-byte chunks per slice (LDGSTS,
// fully coalesced: a slice row is two 32-byte sectors) and keeps two rows' 16 hidden units in registers across the
// slices of a tile.
constexpr int GH_KS = 16;                 // feature columns per slice
constexpr int GH_SS = GH_KS + 4;          // slice row pitch in floats (5 float4: odd -> LDS.128 conflict-free)
constexpr int GH_STAGES = 3;
constexpr int GH_STAGE_FLOATS = GH_ROWS * GH_SS;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(GH_THREADS) gaussian_heads_forward_sliced_kernel(
    long long n, int F, const float* __restrict__ feat, const float* __restrict__ rgb, const float* __restrict__ w1t,
    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
    float* __restrict__ opacity, float* __restrict__ scales, float* __restrict__ rotations, float* __restrict__ colors,
    float* __restrict__ hidden) {
  extern __shared__ __align__(128) float s_mem[];
  const int K = F + 3;
  const int NS = (F + GH_KS - 1) / GH_KS;   // slices per tile
  float* s_stage = s_mem;                                   // [GH_STAGES][GH_ROWS][GH_SS]
  float* s_w1 = s_stage + GH_STAGES * GH_STAGE_FLOATS;      // [max(K, NS*16)][16], rows >= K zero
  const int w_rows = max(K, NS * GH_KS);
  float* s_b1 = s_w1 + w_rows * GH_HID;                     // [16]
  float* s_w2 = s_b1 + GH_HID;                              // [11][4]
  float* s_b2 = s_w2 + GH_OUT * 4;                          // [11]
  const int tid = threadIdx.x;
  for (int i = tid; i < w_rows * GH_HID; i += GH_THREADS) s_w1[i] = i < K * GH_HID ? w1t[i] : 0.f;
  if (tid < GH_HID) s_b1[tid] = b1[tid];
  if (tid < GH_OUT * 4) s_w2[tid] = w2[tid];
  if (tid < GH_OUT) s_b2[tid] = b2[tid];

  const long long tiles = (n + GH_ROWS - 1) / GH_ROWS;
  const long long my_tiles = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long G = my_tiles * NS;  // stages this CTA consumes

  // producer state: the next stage to issue
  long long pf_g = 0, pf_tile = blockIdx.x;
  int pf_slice = 0;
  auto issue = [&]() {
    if (pf_g < G) {
      const long long base = pf_tile * GH_ROWS;
      const int rows = (int)min((long long)GH_ROWS, n - base);
      float* dst = s_stage + (int)(pf_g % GH_STAGES) * GH_STAGE_FLOATS;
      const int kcol = pf_slice * GH_KS + (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < GH_ROWS * 4 / GH_THREADS; q++) {
        const int row = (tid >> 2) + q * (GH_THREADS / 4);
        const bool ok = row < rows && kcol < F;
        const float* src = ok ? feat + (base + row) * F + kcol : feat;
        cp_async16(dst + row * GH_SS + (tid & 3) * 4, src, ok ? 16u : 0u);  // src-size 0: zero fill
      }
      pf_g++;
      if (++pf_slice == NS) {
        pf_slice = 0;
        pf_tile += gridDim.x;
      }
    }
    cp_async_commit();  // (an empty group keeps the wait arithmetic uniform at the end)
  };
#pragma unroll
  for (int st = 0; st < GH_STAGES - 1; st++) issue();

  float h0[GH_HID], h1[GH_HID];
  float c0[3] = {0.f, 0.f, 0.f}, c1[3] = {0.f, 0.f, 0.f};
  long long tile = blockIdx.x;
  int slice = 0;
  for (long long g = 0; g < G; g++) {
    cp_async_wait<GH_STAGES - 2>();
    __syncthreads();  // stage g has landed for every thread; stage g-1's buffer is free for the next issue
    issue();
    const long long base = tile * GH_ROWS;
    const int rows = (int)min((long long)GH_ROWS, n - base);
    if (slice == 0) {
#pragma unroll
      for (int j = 0; j < GH_HID; j++) h0[j] = h1[j] = s_b1[j];
#pragma unroll
      for (int c = 0; c < 3; c++) {  // the colour head's rgb inputs: in flight until the last slice
        c0[c] = tid < rows ? __ldg(rgb + (base + tid) * 3 + c) : 0.f;
        c1[c] = tid + GH_THREADS < rows ? __ldg(rgb + (base + tid + GH_THREADS) * 3 + c) : 0.f;
      }
    }
    const float* st = s_stage + (int)(g % GH_STAGES) * GH_STAGE_FLOATS;
    const float* x0p = st + tid * GH_SS;
    const float* x1p = st + (tid + GH_THREADS) * GH_SS;
    const float* wp = s_w1 + slice * GH_KS * GH_HID;
#pragma unroll
    for (int k4 = 0; k4 < GH_KS / 4; k4++) {
      const float4 xa = *reinterpret_cast<const float4*>(x0p + 4 * k4);
      const float4 xb = *reinterpret_cast<const float4*>(x1p + 4 * k4);
      const float xav[4] = {xa.x, xa.y, xa.z, xa.w};
      const float xbv[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4* w = reinterpret_cast<const float4*>(wp + (4 * k4 + q) * GH_HID);
        const float4 wa = w[0], wb = w[1], wc = w[2], wd = w[3];
        GH_FMA16(h0, xav[q], wa, wb, wc, wd)
        GH_FMA16(h1, xbv[q], wa, wb, wc, wd)
      }
    }
    if (++slice == NS) {
      // ---- the tile is complete: rgb columns, second layers, activations, stores ----
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4* w = reinterpret_cast<const float4*>(s_w1 + (F + c) * GH_HID);
        const float4 wa = w[0], wb = w[1], wc = w[2], wd = w[3];
        GH_FMA16(h0, c0[c], wa, wb, wc, wd)
        GH_FMA16(h1, c1[c], wa, wb, wc, wd)
      }
#pragma unroll
      for (int half = 0; half < 2; half++) {
        float* h = half ? h1 : h0;
        const int r = tid + half * GH_THREADS;
        if (r < rows) {
          const long long gi = base + r;
#pragma unroll
          for (int j = 0; j < GH_HID; j++) h[j] = fmaxf(h[j], 0.f);
          float z[GH_OUT];
          gh_second_layer(h, s_w2, s_b2, z);
          scales[gi * 3 + 0] = gh_softplus(z[0]);
          scales[gi * 3 + 1] = gh_softplus(z[1]);
          scales[gi * 3 + 2] = gh_softplus(z[2]);
          const float nrm = fmaxf(sqrtf(z[3] * z[3] + z[4] * z[4] + z[5] * z[5] + z[6] * z[6]), 1e-12f);  // F.normalize eps
          reinterpret_cast<float4*>(rotations)[gi] = make_float4(z[3] / nrm, z[4] / nrm, z[5] / nrm, z[6] / nrm);
          opacity[gi] = gh_sigmoid(z[7]);
          colors[gi * 3 + 0] = gh_sigmoid(z[8]);
          colors[gi * 3 + 1] = gh_sigmoid(z[9]);
          colors[gi * 3 + 2] = gh_sigmoid(z[10]);
          float4* hp = reinterpret_cast<float4*>(hidden + gi * GH_HID);
          hp[0] = make_float4(h[0], h[1], h[2], h[3]);
          hp[1] = make_float4(h[4], h[5], h[6], h[7]);
          hp[2] = make_float4(h[8], h[9], h[10], h[11]);
          hp[3] = make_float4(h[12], h[13], h[14], h[15]);
        }
      }
      slice = 0;
      tile += gridDim.x;
    }
  }
  cp_async_wait<0>();
}

// Backward.  Persistent CTAs: each keeps its share of the weight gradients in registers across all its tiles and
// adds them to the global sums once at the end.  Per tile:
//   (1) per row: z2 from the saved hidden, activation derivatives -> dL/dz2 [11] -> dL/dhidden [16] (ReLU-masked);
//       runs while the feature tile is still in flight
//   (2) weight gradients: thread u owns input columns k = u and u + 64 (16 sums each); the second warp also owns
//       b1 (16 sums) and w2/b2 (11 rows)
//   (3) dL/dfeat rows overwrite the feature rows in shared memory and leave through bulk stores.
__global__ void __launch_bounds__(GH_THREADS) gaussian_heads_backward_kernel(
    long long n, int F, int fast, const float* __restrict__ feat, const float* __restrict__ rgb,
    const float* __restrict__ w1t, const float* __restrict__ w2, const float* __restrict__ b2,
    const float* __restrict__ hidden, const float* __restrict__ g_opacity, const float* __restrict__ g_scales,
    const float* __restrict__ g_rotations, const float* __restrict__ g_colors, float* __restrict__ g_feat,
    float* __restrict__ g_w1t, float* __restrict__ g_b1, float* __restrict__ g_w2, float* __restrict__ g_b2) {
  extern __shared__ __align__(128) float s_mem[];
  __shared__ __align__(8) uint64_t s_bar;
  const int XS = gh_row_stride(F);
  const int K = F + 3;
  float* s_x = s_mem;                        // [GH_ROWS][XS]   features (+rgb), later dL/dfeat
  float* s_gh = s_x + GH_ROWS * XS;          // [GH_ROWS][16]   dL/dhidden
  float* s_hh = s_gh + GH_ROWS * GH_HID;     // [GH_ROWS][16]   hidden
  float* s_gz = s_hh + GH_ROWS * GH_HID;     // [GH_ROWS][12]   dL/dz2
  float* s_w1 = s_gz + GH_ROWS * 12;         // [K][16]
  float* s_w2 = s_w1 + K * GH_HID;           // [11][4]
  float* s_b2 = s_w2 + GH_OUT * 4;           // [11]
  const int tid = threadIdx.x;
  for (int i = tid; i < K * GH_HID; i += GH_THREADS) s_w1[i] = w1t[i];
  if (tid < GH_OUT * 4) s_w2[tid] = w2[tid];
  if (tid < GH_OUT) s_b2[tid] = b2[tid];
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  float acc_a[GH_HID], acc_b[GH_HID];  // dL/dw1t rows k = tid and k = tid + 64
#pragma unroll
  for (int j = 0; j < GH_HID; j++) acc_a[j] = acc_b[j] = 0.f;
  float acc_small[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // second warp: lane < 16 -> b1[lane]; 16 <= lane < 27 -> w2 row, b2
  const int ka = tid, kb = tid + GH_THREADS;
  const bool dual = ((tid & ~31) + GH_THREADS) < K;  // warp-uniform: does this warp own any second column?

  const long long tiles = (n + GH_ROWS - 1) / GH_ROWS;
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long base = tile * GH_ROWS;
    const int rows = (int)min((long long)GH_ROWS, n - base);
    gh_fetch_tile(feat, rgb, base, rows, F, XS, s_x, &s_bar, fast != 0);

    // ---- (1) ----
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const int r = tid + half * GH_THREADS;
      float gh[GH_HID], h[GH_HID], gz[12];
#pragma unroll
      for (int j = 0; j < GH_HID; j++) gh[j] = h[j] = 0.f;
#pragma unroll
      for (int i = 0; i < 12; i++) gz[i] = 0.f;
      if (r < rows) {
        const long long g = base + r;
        const float4* hp = reinterpret_cast<const float4*>(hidden + g * GH_HID);
        const float4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2), h3 = __ldg(hp + 3);
        h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
        h[8] = h2.x; h[9] = h2.y; h[10] = h2.z; h[11] = h2.w; h[12] = h3.x; h[13] = h3.y; h[14] = h3.z; h[15] = h3.w;
        float z[GH_OUT];
        gh_second_layer(h, s_w2, s_b2, z);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float sg = z[c] > 20.f ? 1.f : gh_sigmoid(z[c]);  // d softplus
          gz[c] = __ldg(g_scales + g * 3 + c) * sg;
        }
        {
          const float4 gr = __ldg(reinterpret_cast<const float4*>(g_rotations) + g);
          const float n2 = z[3] * z[3] + z[4] * z[4] + z[5] * z[5] + z[6] * z[6];
          const float nrm = sqrtf(n2);
          if (nrm > 1e-12f) {
            const float inv = 1.f / nrm;
            const float dotv = (gr.x * z[3] + gr.y * z[4] + gr.z * z[5] + gr.w * z[6]) * inv * inv;
            gz[3] = (gr.x - z[3] * dotv) * inv;
            gz[4] = (gr.y - z[4] * dotv) * inv;
            gz[5] = (gr.z - z[5] * dotv) * inv;
            gz[6] = (gr.w - z[6] * dotv) * inv;
          } else {  // clamped norm: y = z / eps
            gz[3] = gr.x * 1e12f; gz[4] = gr.y * 1e12f; gz[5] = gr.z * 1e12f; gz[6] = gr.w * 1e12f;
          }
        }
        {
          const float y = gh_sigmoid(z[7]);
          gz[7] = __ldg(g_opacity + g) * y * (1.f - y);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float y = gh_sigmoid(z[8 + c]);
          gz[8 + c] = __ldg(g_colors + g * 3 + c) * y * (1.f - y);
        }
#pragma unroll
        for (int i = 0; i < GH_OUT; i++) {
          const int hb = gh_hidden_base(i);
#pragma unroll
          for (int j = 0; j < 4; j++) gh[hb + j] = fmaf(s_w2[i * 4 + j], gz[i], gh[hb + j]);
        }
#pragma unroll
        for (int j = 0; j < GH_HID; j++) gh[j] = h[j] > 0.f ? gh[j] : 0.f;
      }
      float4* ghp = reinterpret_cast<float4*>(s_gh + r * GH_HID);
      float4* hhp = reinterpret_cast<float4*>(s_hh + r * GH_HID);
      float4* gzp = reinterpret_cast<float4*>(s_gz + r * 12);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        ghp[q] = make_float4(gh[4 * q], gh[4 * q + 1], gh[4 * q + 2], gh[4 * q + 3]);
        hhp[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
#pragma unroll
      for (int q = 0; q < 3; q++) gzp[q] = make_float4(gz[4 * q], gz[4 * q + 1], gz[4 * q + 2], gz[4 * q + 3]);
    }
    if (fast) {
      mbar_wait(&s_bar, phase);
      phase ^= 1;
    }
    __syncthreads();

    // ---- (2) weight gradients over the tile's valid rows ----
    if (ka < K) {
      if (dual) {
        const bool vb = kb < K;
        for (int t = 0; t < rows; t++) {
          const float xa = s_x[t * XS + ka];
          const float xb = vb ? s_x[t * XS + kb] : 0.f;
          const float4* gp = reinterpret_cast<const float4*>(s_gh + t * GH_HID);
          const float4 wa = gp[0], wb = gp[1], wc = gp[2], wd = gp[3];
          GH_FMA16(acc_a, xa, wa, wb, wc, wd)
          GH_FMA16(acc_b, xb, wa, wb, wc, wd)
        }
      } else {
        for (int t = 0; t < rows; t++) {
          const float xa = s_x[t * XS + ka];
          const float4* gp = reinterpret_cast<const float4*>(s_gh + t * GH_HID);
          const float4 wa = gp[0], wb = gp[1], wc = gp[2], wd = gp[3];
          GH_FMA16(acc_a, xa, wa, wb, wc, wd)
        }
      }
    }
    if (tid >= 32) {
      const int lane = tid - 32;
      if (lane < GH_HID) {
        float s = 0.f;
        for (int t = 0; t < rows; t++) s += s_gh[t * GH_HID + lane];
        acc_small[0] += s;
      } else if (lane < GH_HID + GH_OUT) {
        const int i = lane - GH_HID;
        const int hb = gh_hidden_base(i);
        for (int t = 0; t < rows; t++) {
          const float gzi = s_gz[t * 12 + i];
          const float4 hv = *reinterpret_cast<const float4*>(s_hh + t * GH_HID + hb);
          acc_small[0] = fmaf(gzi, hv.x, acc_small[0]);
          acc_small[1] = fmaf(gzi, hv.y, acc_small[1]);
          acc_small[2] = fmaf(gzi, hv.z, acc_small[2]);
          acc_small[3] = fmaf(gzi, hv.w, acc_small[3]);
          acc_small[4] += gzi;
        }
      }
    }
    __syncthreads();

    // ---- (3) dL/dfeat = w1t . dL/dhidden, written over this thread's two rows ----
    {
      float g0[GH_HID], g1[GH_HID];
      const float4* p0 = reinterpret_cast<const float4*>(s_gh + tid * GH_HID);
      const float4* p1 = reinterpret_cast<const float4*>(s_gh + (tid + GH_THREADS) * GH_HID);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4 a = p0[q], b = p1[q];
        g0[4 * q] = a.x; g0[4 * q + 1] = a.y; g0[4 * q + 2] = a.z; g0[4 * q + 3] = a.w;
        g1[4 * q] = b.x; g1[4 * q + 1] = b.y; g1[4 * q + 2] = b.z; g1[4 * q + 3] = b.w;
      }
      float* x0p = s_x + tid * XS;
      float* x1p = s_x + (tid + GH_THREADS) * XS;
      const int k4n = K >> 2;  // 4 * k4n >= F: covers every feature column; the rgb columns are scratch
      for (int k4 = 0; k4 < k4n; k4++) {
        float o0[4], o1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4* w = reinterpret_cast<const float4*>(s_w1 + (4 * k4 + q) * GH_HID);
          const float4 wa = w[0], wb = w[1], wc = w[2], wd = w[3];
          const float wv[16] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w, wd.x, wd.y, wd.z, wd.w};
          float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;  // two partial chains per row
#pragma unroll
          for (int j = 0; j < 8; j++) {
            a0 = fmaf(wv[j], g0[j], a0);
            c0 = fmaf(wv[8 + j], g0[8 + j], c0);
            a1 = fmaf(wv[j], g1[j], a1);
            c1 = fmaf(wv[8 + j], g1[8 + j], c1);
          }
          o0[q] = a0 + c0;
          o1[q] = a1 + c1;
        }
        *reinterpret_cast<float4*>(x0p + 4 * k4) = make_float4(o0[0], o0[1], o0[2], o0[3]);
        *reinterpret_cast<float4*>(x1p + 4 * k4) = make_float4(o1[0], o1[1], o1[2], o1[3]);
      }
    }
    if (fast) {
      fence_proxy_async();
      __syncthreads();
#pragma unroll
      for (int r = tid; r < GH_ROWS; r += GH_THREADS)
        if (r < rows) bulk_s2g(g_feat + (base + r) * F, s_x + r * XS, (uint32_t)F * 4u);
      bulk_commit();
      bulk_wait_read();
    } else {
      __syncthreads();
      for (int r = 0; r < rows; r++)
        for (int k = tid; k < F; k += GH_THREADS) g_feat[(base + r) * F + k] = s_x[r * XS + k];
    }
    __syncthreads();  // s_x, s_gh, s_hh, s_gz are free again
  }

  // ---- add this CTA's sums to the global gradients (lane-contiguous atomics through shared memory) ----
  float* s_out = s_x;  // [K][16]
  if (ka < K) {
#pragma unroll
    for (int j = 0; j < GH_HID; j++) s_out[ka * GH_HID + j] = acc_a[j];
    if (kb < K) {
#pragma unroll
      for (int j = 0; j < GH_HID; j++) s_out[kb * GH_HID + j] = acc_b[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < K * GH_HID; i += GH_THREADS) {
    const float v = s_out[i];
    const bool structural_zero = (i >> 4) >= F && (i & 15) < 12;  // S/R/A heads do not see rgb: keep those weights' gradient 0
    if (v != 0.f && !structural_zero) atomicAdd(g_w1t + i, v);
  }
  if (tid >= 32) {
    const int lane = tid - 32;
    if (lane < GH_HID) {
      atomicAdd(g_b1 + lane, acc_small[0]);
    } else if (lane < GH_HID + GH_OUT) {
      const int i = lane - GH_HID;
#pragma unroll
      for (int j = 0; j < 4; j++) atomicAdd(g_w2 + i * 4 + j, acc_small[j]);
      atomicAdd(g_b2 + i, acc_small[4]);
    }
  }
}


// ---- backward, fast path (F % 4 == 0, aligned): two kernels --------------------------------------------------
// (1) rows:    per Gaussian, hidden + upstream gradients -> dL/dz2 -> dL/dhidden (kept in registers, also written to
//              a [n,16] scratch for kernel 2) -> dL/dfeat = w1t . dL/dhidden, produced in 16-column slices that leave
//              through a double-buffered shared-memory slice as coalesced 16-byte stores.  The features themselves are
//              not needed here, so there is no 43 KB tile and 8 CTAs (16 warps) fit per SM.  The 71 small parameter
//              sums (w2, b2, b1) are reduce-scattered over the warp once per tile and kept in two registers per lane.
// (2) weights: dL/dw1t[k][j] = sum_n x[n][k] * dL/dhidden[n][j], a [F+3, n] x [n, 16] product streamed through a 2-deep
//              cp.async ring of 64-row tiles; thread (k-block, j-quad) owns a 4 x 4 block of the result in registers for
//              the whole kernel (two LDS.128 feed 16 FFMAs), one atomicAdd per parameter per CTA at the end.
constexpr int GHB_PITCH = 57;  // 44 (w2) + 11 (b2) per thread, odd pitch: conflict-free rows and columns

__global__ void __launch_bounds__(GH_THREADS) gaussian_heads_backward_rows_kernel(
    long long n, int F, const float* __restrict__ w1t, const float* __restrict__ w2, const float* __restrict__ b2,
    const float* __restrict__ hidden, const float* __restrict__ g_opacity, const float* __restrict__ g_scales,
    const float* __restrict__ g_rotations, const float* __restrict__ g_colors, float* __restrict__ g_hidden /*[n,16]*/,
    float* __restrict__ g_feat, float* __restrict__ g_w2, float* __restrict__ g_b2) {
  extern __shared__ __align__(128) float s_mem[];
  const int NS = (F + GH_KS - 1) / GH_KS;
  float* s_slice = s_mem;                              // [2][GH_ROWS][GH_SS]
  float* s_w1 = s_slice + 2 * GH_STAGE_FLOATS;         // [NS*16][16], rows >= F zero
  float* s_w2 = s_w1 + NS * GH_KS * GH_HID;            // [11][4]
  float* s_b2 = s_w2 + GH_OUT * 4;                     // [11] (+1 pad)
  float* s_small = s_slice;                            // [GH_THREADS][GHB_PITCH]: per-thread dL/dw2 (44), dL/db2 (11);
                                                       // aliases the slice buffers (used before them, barrier between)
  const int tid = threadIdx.x;
  for (int i = tid; i < NS * GH_KS * GH_HID; i += GH_THREADS) s_w1[i] = i < F * GH_HID ? w1t[i] : 0.f;
  if (tid < GH_OUT * 4) s_w2[tid] = w2[tid];
  if (tid < GH_OUT) s_b2[tid] = b2[tid];
  __syncthreads();

  float small = 0.f;  // thread t < 55 keeps parameter sum t (44 dL/dw2, then 11 dL/db2) across all its tiles
  const long long tiles = (n + GH_ROWS - 1) / GH_ROWS;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long base = tile * GH_ROWS;
    const int rows = (int)min((long long)GH_ROWS, n - base);
    float gh[2][GH_HID];
    float* my_small = s_small + tid * GHB_PITCH;
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const int r = tid + half * GH_THREADS;
      float h[GH_HID], gz[GH_OUT];
#pragma unroll
      for (int j = 0; j < GH_HID; j++) gh[half][j] = h[j] = 0.f;
#pragma unroll
      for (int i = 0; i < GH_OUT; i++) gz[i] = 0.f;
      if (r < rows) {
        const long long g = base + r;
        const float4* hp = reinterpret_cast<const float4*>(hidden + g * GH_HID);
        const float4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2), h3 = __ldg(hp + 3);
        h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
        h[8] = h2.x; h[9] = h2.y; h[10] = h2.z; h[11] = h2.w; h[12] = h3.x; h[13] = h3.y; h[14] = h3.z; h[15] = h3.w;
        float z[GH_OUT];
        gh_second_layer(h, s_w2, s_b2, z);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float sg = z[c] > 20.f ? 1.f : gh_sigmoid(z[c]);  // d softplus
          gz[c] = __ldg(g_scales + g * 3 + c) * sg;
        }
        {
          const float4 gr = __ldg(reinterpret_cast<const float4*>(g_rotations) + g);
          const float n2 = z[3] * z[3] + z[4] * z[4] + z[5] * z[5] + z[6] * z[6];
          const float nrm = sqrtf(n2);
          if (nrm > 1e-12f) {
            const float inv = 1.f / nrm;
            const float dotv = (gr.x * z[3] + gr.y * z[4] + gr.z * z[5] + gr.w * z[6]) * inv * inv;
            gz[3] = (gr.x - z[3] * dotv) * inv;
            gz[4] = (gr.y - z[4] * dotv) * inv;
            gz[5] = (gr.z - z[5] * dotv) * inv;
            gz[6] = (gr.w - z[6] * dotv) * inv;
          } else {  // clamped norm: y = z / eps
            gz[3] = gr.x * 1e12f; gz[4] = gr.y * 1e12f; gz[5] = gr.z * 1e12f; gz[6] = gr.w * 1e12f;
          }
        }
        {
          const float y = gh_sigmoid(z[7]);
          gz[7] = __ldg(g_opacity + g) * y * (1.f - y);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float y = gh_sigmoid(z[8 + c]);
          gz[8 + c] = __ldg(g_colors + g * 3 + c) * y * (1.f - y);
        }
#pragma unroll
        for (int i = 0; i < GH_OUT; i++) {
          const int hb = gh_hidden_base(i);
#pragma unroll
          for (int j = 0; j < 4; j++) gh[half][hb + j] = fmaf(s_w2[i * 4 + j], gz[i], gh[half][hb + j]);
        }
#pragma unroll
        for (int j = 0; j < GH_HID; j++) gh[half][j] = h[j] > 0.f ? gh[half][j] : 0.f;
        float4* gp = reinterpret_cast<float4*>(g_hidden + g * GH_HID);
#pragma unroll
        for (int q = 0; q < 4; q++)
          gp[q] = make_float4(gh[half][4 * q], gh[half][4 * q + 1], gh[half][4 * q + 2], gh[half][4 * q + 3]);
      }
      // this thread's share of dL/dw2, dL/db2 goes to its own shared-memory row (h = gz = 0 for rows past the end)
#pragma unroll
      for (int i = 0; i < GH_OUT; i++) {
        const int hb = gh_hidden_base(i);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float v = gz[i] * h[hb + j];
          my_small[i * 4 + j] = half ? my_small[i * 4 + j] + v : v;
        }
        my_small[44 + i] = half ? my_small[44 + i] + gz[i] : gz[i];
      }
    }
    __syncthreads();
    if (tid < 55) {
      float acc = 0.f;
#pragma unroll 8
      for (int t = 0; t < GH_THREADS; t++) acc += s_small[t * GHB_PITCH + tid];
      small += acc;
    }
    __syncthreads();  // s_small is dead: its storage becomes the first dL/dfeat slice
    // dL/dfeat in slices of 16 columns through the double-buffered slice
    for (int s = 0; s < NS; s++) {
      float* buf = s_slice + (s & 1) * GH_STAGE_FLOATS;
      const float* wp = s_w1 + s * GH_KS * GH_HID;
#pragma unroll
      for (int k4 = 0; k4 < GH_KS / 4; k4++) {
        float o0[4], o1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4* w = reinterpret_cast<const float4*>(wp + (4 * k4 + q) * GH_HID);
          const float4 wa = w[0], wb = w[1], wc = w[2], wd = w[3];
          const float wv[16] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w, wd.x, wd.y, wd.z, wd.w};
          float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            a0 = fmaf(wv[j], gh[0][j], a0);
            c0 = fmaf(wv[8 + j], gh[0][8 + j], c0);
            a1 = fmaf(wv[j], gh[1][j], a1);
            c1 = fmaf(wv[8 + j], gh[1][8 + j], c1);
          }
          o0[q] = a0 + c0;
          o1[q] = a1 + c1;
        }
        *reinterpret_cast<float4*>(buf + tid * GH_SS + 4 * k4) = make_float4(o0[0], o0[1], o0[2], o0[3]);
        *reinterpret_cast<float4*>(buf + (tid + GH_THREADS) * GH_SS + 4 * k4) = make_float4(o1[0], o1[1], o1[2], o1[3]);
      }
      __syncthreads();  // slice complete; the other buffer's readers finished before the previous barrier
      const int kcol = s * GH_KS + (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < GH_ROWS * 4 / GH_THREADS; q++) {
        const int row = (tid >> 2) + q * (GH_THREADS / 4);
        if (row < rows && kcol < F)
          *reinterpret_cast<float4*>(g_feat + (base + row) * F + kcol) =
              *reinterpret_cast<const float4*>(buf + row * GH_SS + (tid & 3) * 4);
      }
    }
    __syncthreads();
  }
  if (tid < 44) atomicAdd(g_w2 + tid, small);
  else if (tid < 55) atomicAdd(g_b2 + (tid - 44), small);
}

constexpr int GHW_MAX_THREADS = 160;  // owner warps for the 4 * ceil((F+3)/4) (k-block, j-quad) pairs + one warp for dL/db1
#ifndef OCRF_GHW_ROWS
#define OCRF_GHW_ROWS 64
#endif
#ifndef OCRF_GHW_STAGES
#define OCRF_GHW_STAGES 2
#endif
constexpr int GHW_ROWS = OCRF_GHW_ROWS;
constexpr int GHW_STAGES = OCRF_GHW_STAGES;

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}

__global__ void __launch_bounds__(GHW_MAX_THREADS) gaussian_heads_backward_weights_kernel(
    long long n, int F, const float* __restrict__ feat, const float* __restrict__ rgb,
    const float* __restrict__ g_hidden /*[n,16]*/, float* __restrict__ g_w1t, float* __restrict__ g_b1) {
  extern __shared__ __align__(128) float s_mem[];
  const int XP = F + 4;                                   // row pitch: [feat F | rgb 3 | 0]
  const int stage_floats = GHW_ROWS * (XP + GH_HID);
  const int tid = threadIdx.x;
  const int kb = tid >> 2, jq = tid & 3;
  const int kblocks = (F + 3 + 3) / 4;
  const bool owner = kb < kblocks;
  // the pad column F+3 of every x row is never written by the copies: zero it once
  for (int i = tid; i < GHW_STAGES * GHW_ROWS; i += (int)blockDim.x)
    s_mem[(i / GHW_ROWS) * stage_floats + (i % GHW_ROWS) * XP + F + 3] = 0.f;

  const long long tiles = (n + GHW_ROWS - 1) / GHW_ROWS;
  const long long my_tiles = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  long long pf = 0;
  const int f4 = F / 4;
  auto issue = [&]() {
    if (pf < my_tiles) {
      const long long base = (blockIdx.x + pf * gridDim.x) * GHW_ROWS;
      const int rows = (int)min((long long)GHW_ROWS, n - base);
      float* sx = s_mem + (int)(pf % GHW_STAGES) * stage_floats;
      float* sg = sx + GHW_ROWS * XP;
      // features: lane = 16-byte column chunk, one row per warp per iteration (no integer division in the loop)
      for (int c0 = 0; c0 < f4; c0 += 32) {
        const int col = c0 + (tid & 31);
        if (col < f4) {
#pragma unroll 4
          for (int row = tid >> 5; row < GHW_ROWS; row += (int)(blockDim.x >> 5)) {
            const bool ok = row < rows;
            cp_async16(sx + row * XP + 4 * col, ok ? feat + (base + row) * F + 4 * col : feat, ok ? 16u : 0u);
          }
        }
      }
#pragma unroll
      for (int c = tid; c < GHW_ROWS * 4; c += (int)blockDim.x) {           // dL/dhidden: 4 chunks per row
        const int row = c >> 2, col = c & 3;
        const bool ok = row < rows;
        cp_async16(sg + row * GH_HID + 4 * col, ok ? g_hidden + (base + row) * GH_HID + 4 * col : g_hidden, ok ? 16u : 0u);
        if (col < 3)                                                     // rgb: 4-byte copies into columns F..F+2
          cp_async4(sx + row * XP + F + col, ok ? rgb + (base + row) * 3 + col : rgb, ok ? 4u : 0u);
      }
      pf++;
    }
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < GHW_STAGES - 1; st++) issue();

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
  float bsum = 0.f;  // dL/db1[lane], kept by the last warp of the CTA (the host adds it after the owner warps)

  for (long long t = 0; t < my_tiles; t++) {
    cp_async_wait<GHW_STAGES - 2>();
    __syncthreads();
    issue();
    const float* sx = s_mem + (int)(t % GHW_STAGES) * stage_floats;
    const float* sg = sx + GHW_ROWS * XP;
    if (owner) {
      const float* xp = sx + 4 * kb;
      const float* gp = sg + 4 * jq;
#pragma unroll 4
      for (int r = 0; r < GHW_ROWS; r++) {   // rows past the end of the last tile are zero-filled
        const float4 x = *reinterpret_cast<const float4*>(xp + r * XP);
        const float4 g = *reinterpret_cast<const float4*>(gp + r * GH_HID);
        acc[0][0] = fmaf(x.x, g.x, acc[0][0]); acc[0][1] = fmaf(x.x, g.y, acc[0][1]);
        acc[0][2] = fmaf(x.x, g.z, acc[0][2]); acc[0][3] = fmaf(x.x, g.w, acc[0][3]);
        acc[1][0] = fmaf(x.y, g.x, acc[1][0]); acc[1][1] = fmaf(x.y, g.y, acc[1][1]);
        acc[1][2] = fmaf(x.y, g.z, acc[1][2]); acc[1][3] = fmaf(x.y, g.w, acc[1][3]);
        acc[2][0] = fmaf(x.z, g.x, acc[2][0]); acc[2][1] = fmaf(x.z, g.y, acc[2][1]);
        acc[2][2] = fmaf(x.z, g.z, acc[2][2]); acc[2][3] = fmaf(x.z, g.w, acc[2][3]);
        acc[3][0] = fmaf(x.w, g.x, acc[3][0]); acc[3][1] = fmaf(x.w, g.y, acc[3][1]);
        acc[3][2] = fmaf(x.w, g.z, acc[3][2]); acc[3][3] = fmaf(x.w, g.w, acc[3][3]);
      }
    } else if (tid >= (int)blockDim.x - 32 && (tid & 31) < GH_HID) {  // dL/db1: one hidden unit per lane of the last warp
      const float* gp = sg + (tid & 31);
      float b = 0.f;
#pragma unroll 8
      for (int r = 0; r < GHW_ROWS; r++) b += gp[r * GH_HID];
      bsum += b;
    }
  }
  cp_async_wait<0>();
  if (owner) {
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int k = 4 * kb + a;
      if (k >= F + 3) continue;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int j = 4 * jq + b;
        const bool structural_zero = k >= F && j < 12;  // the S/R/A heads do not see rgb
        if (!structural_zero && acc[a][b] != 0.f) atomicAdd(g_w1t + k * GH_HID + j, acc[a][b]);
      }
    }
  } else if (tid >= (int)blockDim.x - 32 && (tid & 31) < GH_HID) {
    atomicAdd(g_b1 + (tid & 31), bsum);
  }
}

}  // namespace ocrf

using namespace ocrf;

static size_t gh_fwd_smem(int F) {
  return ((size_t)GH_ROWS * gh_row_stride(F) + (size_t)(F + 3) * GH_HID + GH_HID + GH_OUT * 4 + 12) * 4;
}
static size_t gh_bwd_smem(int F) {
  return ((size_t)GH_ROWS * gh_row_stride(F) + (size_t)GH_ROWS * (GH_HID * 2 + 12) + (size_t)(F + 3) * GH_HID +
          GH_OUT * 4 + 12) * 4;
}
static bool gh_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int gh_grid(long long n, size_t smem) {
  const long long tiles = (n + GH_ROWS - 1) / GH_ROWS;
  const int per_sm = (int)((220 * 1024) / (smem + 1024));
  const long long cap = (long long)num_sms() * (per_sm < 1 ? 1 : per_sm);
  return (int)(tiles < cap ? tiles : cap);
}

extern "C" int ocrf_gaussian_heads_forward(void* stream, int64_t n, int32_t F, const float* feat, const float* rgb,
                                           const float* w1t, const float* b1, const float* w2, const float* b2,
                                           float* opacity, float* scales, float* rotations, float* colors,
                                           float* hidden) {
  if (n < 0 || F <= 0 || F + 3 > GH_MAXK) return OCRF_EINVAL;
  if (n == 0) return 0;
  if (!feat || !rgb || !w1t || !b1 || !w2 || !b2 || !opacity || !scales || !rotations || !colors || !hidden) return OCRF_EINVAL;
  if (!gh_aligned16(rotations) || !gh_aligned16(hidden)) return OCRF_EINVAL;
  if ((F % 4 == 0) && gh_aligned16(feat)) {
    const int ns = (F + GH_KS - 1) / GH_KS;
    const int w_rows = (F + 3) > ns * GH_KS ? (F + 3) : ns * GH_KS;
    const size_t smem = ((size_t)GH_STAGES * GH_STAGE_FLOATS + (size_t)w_rows * GH_HID + GH_HID + GH_OUT * 4 + 12) * 4;
    cudaError_t e = cudaFuncSetAttribute(gaussian_heads_forward_sliced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    gaussian_heads_forward_sliced_kernel<<<gh_grid(n, smem), GH_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        n, F, feat, rgb, w1t, b1, w2, b2, opacity, scales, rotations, colors, hidden);
  } else {  // odd channel counts / unaligned views: whole-row tiles filled with plain loads
    const size_t smem = gh_fwd_smem(F);
    cudaError_t e = cudaFuncSetAttribute(gaussian_heads_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gaussian_heads_forward_kernel<<<gh_grid(n, smem), GH_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        n, F, 0, feat, rgb, w1t, b1, w2, b2, opacity, scales, rotations, colors, hidden);
  }
  OCRF_CHECK_LAST();
  return 0;
}

extern "C" size_t ocrf_gaussian_heads_backward_workspace_bytes(int64_t n) {
  return (size_t)(n > 0 ? n : 1) * GH_HID * sizeof(float) + 128;
}

extern "C" int ocrf_gaussian_heads_backward(void* stream, int64_t n, int32_t F, const float* feat, const float* rgb,
                                            const float* w1t, const float* w2, const float* b2, const float* hidden,
                                            const float* g_opacity, const float* g_scales, const float* g_rotations,
                                            const float* g_colors, float* g_feat, float* g_w1t, float* g_b1,
                                            float* g_w2, float* g_b2, void* ws) {
  if (n < 0 || F <= 0 || F + 3 > GH_MAXK) return OCRF_EINVAL;
  if (n == 0) return 0;
  if (!feat || !rgb || !w1t || !w2 || !b2 || !hidden || !g_opacity || !g_scales || !g_rotations || !g_colors || !g_feat ||
      !g_w1t || !g_b1 || !g_w2 || !g_b2)
    return OCRF_EINVAL;
  if (!gh_aligned16(g_rotations) || !gh_aligned16(hidden)) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((F % 4 == 0) && gh_aligned16(feat) && gh_aligned16(g_feat) && ws != nullptr && gh_aligned16(ws)) {
    float* g_hidden = static_cast<float*>(ws);
    const int ns = (F + GH_KS - 1) / GH_KS;
    const size_t smem1 = ((size_t)2 * GH_STAGE_FLOATS + (size_t)ns * GH_KS * GH_HID + GH_OUT * 4 + 12) * 4;
    cudaError_t e = cudaFuncSetAttribute(gaussian_heads_backward_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem1);
    if (e != cudaSuccess) return (int)e;
    const long long tiles1 = (n + GH_ROWS - 1) / GH_ROWS;
    const long long cap1 = (long long)num_sms() * 8;  // residency of the rows kernel: 128 registers x 64 threads, 26 KB
    gaussian_heads_backward_rows_kernel<<<(unsigned)(tiles1 < cap1 ? tiles1 : cap1), GH_THREADS, smem1, st>>>(
        n, F, w1t, w2, b2, hidden, g_opacity, g_scales, g_rotations, g_colors, g_hidden, g_feat, g_w2, g_b2);
    const size_t smem2 = (size_t)GHW_STAGES * GHW_ROWS * (F + 4 + GH_HID) * 4;
    e = cudaFuncSetAttribute(gaussian_heads_backward_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return (int)e;
    const long long tiles2 = (n + GHW_ROWS - 1) / GHW_ROWS;
    const int per_sm = (int)((220 * 1024) / (smem2 + 1024));
    const long long cap = (long long)num_sms() * (per_sm < 1 ? 1 : per_sm);
    const int owner_threads = 4 * ((F + 3 + 3) / 4);
    const int threads2 = ((owner_threads + 31) / 32) * 32 + 32;  // + the dL/db1 warp
    gaussian_heads_backward_weights_kernel<<<(unsigned)(tiles2 < cap ? tiles2 : cap), threads2, smem2, st>>>(
        n, F, feat, rgb, g_hidden, g_w1t, g_b1);
  } else {  // odd channel counts / unaligned views / no workspace: the single fused kernel with whole-row tiles
    const size_t smem = gh_bwd_smem(F);
    cudaError_t e = cudaFuncSetAttribute(gaussian_heads_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int fast = (F % 4 == 0) && gh_aligned16(feat) && gh_aligned16(g_feat);
    gaussian_heads_backward_kernel<<<gh_grid(n, smem), GH_THREADS, smem, st>>>(
        n, F, fast, feat, rgb, w1t, w2, b2, hidden, g_opacity, g_scales, g_rotations, g_colors, g_feat, g_w1t, g_b1,
        g_w2, g_b2);
  }
  OCRF_CHECK_LAST();
  return 0;
}
