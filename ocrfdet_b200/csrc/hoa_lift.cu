// Stage 5a: the height-aware opacity lift -- per-voxel opacity and alpha_lidar -> opacity_alpha.
//
// Replaces view_transformer_ocrf.py:1159-1161 of the reference,
//     opacity_up = interpolate(opacity, (W/6, L/6), bilinear, align_corners=True)      [B,13,128,128] -> [B,13,21,21]
//     alpha_up   = interpolate(alpha_lidar, ...)
//     opacity_alpha = interpolate(DeformableAttention2D(opacity_up, alpha_up), (W, L)) + opacity
// with DeformableAttention2D of mmdet3d/ops/cross_attention_2d.py:93-220 in OcRFDet's configuration
// (view_transformer_ocrf.py:639-648: dim 13, one head of 8, one offset group, 6x6 stride-4 offset conv, offset scale
// 4, continuous position bias MLP 2 -> 3 -> 3 -> 1).  In torch this is ~45 launches of kernels that each touch a few
// kilobytes; the whole attention works on 441 queries x 25 keys.
//
// Forward : lift_coarse_kernel   one CTA per sample: both down-samplings, q, the offset network, the bilinear key/value
//                                sampling, logits + position bias + softmax (+ dropout keep-mask) and the output
//                                projection, all in shared memory; writes the 13 x 21 x 21 attention output
//           lift_upsample_kernel streams opacity once: out = bilinear(att) + opacity (float4)
// Backward: lift_grad_up_kernel  adjoint of the up-sampling (one CTA per (sample, plane), plane staged in shared memory)
//           lift_grad_coarse_kernel  one CTA per sample: recomputes the forward state, then every gradient of the
//                                attention in closed form (the reductions over queries are loops of the thread that owns
//                                the output: no shared-memory float atomics); parameter gradients are accumulated
//                                into the packed vector with one global reduction per parameter per sample
//           d/d opacity = g + adjoint-downsample(d/d opacity_up): a copy plus 4 x 13 x 441 sparse taps (the taps of
//                                different coarse points never coincide), same for d/d alpha on a zero fill.
// Resize taps follow ATen's float32 expressions (UpSample.h) so the sampling weights carry the reference's rounding.
#include "common.cuh"

namespace ocrf {
namespace hoa {

constexpr int DIM = 13, INNER = 8, HID = 3, KS = 6, DS = 4, PAD = 1;
constexpr float OFFSET_SCALE = 4.f;
constexpr int THREADS = 512, WARPS = THREADS / 32;
constexpr int MAXK = 36;  // keys: offset-map positions (5 x 5 = 25 at the reference size)
// packed parameters = the reference's named_parameters() order
constexpr int P_WDW = 0, P_BDW = 288, P_WPW = 296, P_W1 = 312, P_B1 = 318, P_W2 = 321, P_B2 = 330, P_W3 = 333,
              P_B3 = 336, P_WQ = 337, P_WK = 441, P_WV = 545, P_WO = 649, P_BO = 753, P_TOTAL = 766;

struct Dims {
  int H, W;    // full resolution
  int ch, cw;  // coarse (query) resolution: int(H / 6), int(W / 6)
  int hk, wk;  // offset map = key resolution
  int nq, nk;
};

__host__ __device__ inline Dims make_dims(int H, int W) {
  Dims d;
  d.H = H; d.W = W;
  d.ch = H / 6; d.cw = W / 6;
  d.hk = (d.ch + 2 * PAD - KS) / DS + 1;
  d.wk = (d.cw + 2 * PAD - KS) / DS + 1;
  d.nq = d.ch * d.cw;
  d.nk = d.hk * d.wk;
  return d;
}

// ATen area_pixel_compute_source_index for align_corners=True, float32
__device__ __forceinline__ void ac_tap(int dst, int n_in, int n_out, int& i0, int& i1, float& w0, float& w1) {
  const float scale = n_out > 1 ? __fdiv_rn((float)(n_in - 1), (float)(n_out - 1)) : 0.f;
  const float src = __fmul_rn(scale, (float)dst);
  i0 = min((int)src, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  w1 = __fsub_rn(src, (float)i0);
  w0 = __fsub_rn(1.f, w1);
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * __expf(-0.5f * x * x) * 0.3989422804014327f;
}

// Shared-memory state of one sample (float offsets into the dynamic array)
struct Smem {
  float* p;     // [P_TOTAL] parameters
  float* xq;    // [DIM][nq]
  float* xkv;   // [DIM][nq]
  float* q;     // [INNER][nq]  (unscaled)
  float* dwv;   // [INNER][nk]  offset network before the GELU
  float* ge;    // [INNER][nk]
  float* th;    // [2][nk] tanh
  float* vn;    // [nk][2] normalised sampling positions
  float* kvf;   // [DIM][nk]
  float* k;     // [INNER][nk]
  float* v;     // [INNER][nk]
};

__host__ __device__ inline size_t smem_forward_floats(const Dims& d) {
  return (size_t)P_TOTAL + 2 + 2 * DIM * d.nq + INNER * d.nq + 2 * INNER * d.nk + 4 * d.nk + DIM * d.nk + 2 * INNER * d.nk;
}

__device__ inline float* carve(Smem& s, float* base, const Dims& d) {
  float* f = base;
  s.p = f;   f += (P_TOTAL + 2);
  s.xq = f;  f += DIM * d.nq;
  s.xkv = f; f += DIM * d.nq;
  s.q = f;   f += INNER * d.nq;
  s.dwv = f; f += INNER * d.nk;
  s.ge = f;  f += INNER * d.nk;
  s.th = f;  f += 2 * d.nk;
  s.vn = f;  f += 2 * d.nk;
  s.kvf = f; f += DIM * d.nk;
  s.k = f;   f += INNER * d.nk;
  s.v = f;   f += INNER * d.nk;
  return f;
}

// bilinear taps of grid_sample (align_corners=False, zero padding) at normalised (x, y)
struct Taps {
  int x0, y0;
  float fx, fy;
};
__device__ __forceinline__ Taps sample_taps(float nx, float ny, int h, int w) {
  const float ix = ((nx + 1.f) * w - 1.f) * 0.5f, iy = ((ny + 1.f) * h - 1.f) * 0.5f;
  Taps t;
  const float flx = floorf(ix), fly = floorf(iy);
  t.x0 = (int)flx; t.y0 = (int)fly;
  t.fx = ix - flx; t.fy = iy - fly;
  return t;
}

// Everything of the forward up to (k, v): called by all threads of the CTA.
__device__ void lift_core(const Smem& s, const Dims& d, const float* __restrict__ opacity,
                          const float* __restrict__ alpha, const float* __restrict__ params) {
  const int tid = threadIdx.x;
  for (int i = tid; i < P_TOTAL; i += THREADS) s.p[i] = params[i];
  // (1) both down-samplings (view_transformer_ocrf.py:1159-1160)
  const size_t plane = (size_t)d.H * d.W;
  for (int e = tid; e < 2 * DIM * d.nq; e += THREADS) {
    const int which = e / (DIM * d.nq), r = e - which * DIM * d.nq;
    const int c = r / d.nq, i = r - c * d.nq;
    const int cy = i / d.cw, cx = i - cy * d.cw;
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    ac_tap(cy, d.H, d.ch, y0, y1, wy0, wy1);
    ac_tap(cx, d.W, d.cw, x0, x1, wx0, wx1);
    const float* src = (which ? alpha : opacity) + c * plane;
    const float v = wy0 * (wx0 * __ldg(src + (size_t)y0 * d.W + x0) + wx1 * __ldg(src + (size_t)y0 * d.W + x1)) +
                    wy1 * (wx0 * __ldg(src + (size_t)y1 * d.W + x0) + wx1 * __ldg(src + (size_t)y1 * d.W + x1));
    (which ? s.xkv : s.xq)[r] = v;
  }
  __syncthreads();
  // (2) q = to_q(x_q)  (cross_attention_2d.py:155)
  for (int e = tid; e < INNER * d.nq; e += THREADS) {
    const int o = e / d.nq, i = e - o * d.nq;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < DIM; c++) acc = fmaf(s.p[P_WQ + o * DIM + c], s.xq[c * d.nq + i], acc);
    s.q[e] = acc;
  }
  __syncthreads();
  // (3) to_offsets (:133-139): depthwise 6x6 stride-4 conv (+bias), GELU
  for (int e = tid; e < INNER * d.nk; e += THREADS) {
    const int c = e / d.nk, j = e - c * d.nk;
    const int oy = j / d.wk, ox = j - oy * d.wk;
    float acc = s.p[P_BDW + c];
    for (int ky = 0; ky < KS; ky++) {
      const int y = oy * DS - PAD + ky;
      if (y < 0 || y >= d.ch) continue;
      for (int kx = 0; kx < KS; kx++) {
        const int x = ox * DS - PAD + kx;
        if (x < 0 || x >= d.cw) continue;
        acc = fmaf(s.p[P_WDW + c * KS * KS + ky * KS + kx], s.q[c * d.nq + y * d.cw + x], acc);
      }
    }
    s.dwv[e] = acc;
    s.ge[e] = gelu_f(acc);
  }
  __syncthreads();
  //     1x1 conv to 2, tanh, * offset_scale; grid + offsets; normalize_grid (:166-174)
  for (int e = tid; e < 2 * d.nk; e += THREADS) {
    const int o = e / d.nk, j = e - o * d.nk;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < INNER; c++) acc = fmaf(s.p[P_WPW + o * INNER + c], s.ge[c * d.nk + j], acc);
    const float t = tanhf(acc);
    s.th[e] = t;
    const int oy = j / d.wk, ox = j - oy * d.wk;
    const float vg = (o == 0 ? (float)ox : (float)oy) + OFFSET_SCALE * t;
    const int denom = max((o == 0 ? d.hk : d.wk) - 1, 1);  // the reference divides x by (h - 1), y by (w - 1)
    s.vn[j * 2 + o] = 2.f * vg / (float)denom - 1.f;
  }
  __syncthreads();
  // (4) kv_feats = grid_sample(x_kv, vgrid_scaled) (:176-179)
  for (int e = tid; e < DIM * d.nk; e += THREADS) {
    const int c = e / d.nk, j = e - c * d.nk;
    const Taps t = sample_taps(s.vn[2 * j], s.vn[2 * j + 1], d.ch, d.cw);
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dx = 0; dx < 2; dx++) {
        const int y = t.y0 + dy, x = t.x0 + dx;
        if (y < 0 || y >= d.ch || x < 0 || x >= d.cw) continue;
        acc = fmaf((dy ? t.fy : 1.f - t.fy) * (dx ? t.fx : 1.f - t.fx), s.xkv[c * d.nq + y * d.cw + x], acc);
      }
    s.kvf[e] = acc;
  }
  __syncthreads();
  //     k, v = to_k(kv_feats), to_v(kv_feats) (:185)
  for (int e = tid; e < 2 * INNER * d.nk; e += THREADS) {
    const int which = e / (INNER * d.nk), r = e - which * INNER * d.nk;
    const int o = r / d.nk, j = r - o * d.nk;
    const float* wgt = s.p + (which ? P_WV : P_WK) + o * DIM;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < DIM; c++) acc = fmaf(wgt[c], s.kvf[c * d.nk + j], acc);
    (which ? s.v : s.k)[r] = acc;
  }
  __syncthreads();
}

// continuous position bias of (query i, key j) (:52-88); optionally returns the hidden state for the backward
struct BiasState {
  float pos[2], bb[2], z1[HID], z2[HID];
};
__device__ __forceinline__ float cpb_bias(const float* p, float gqx, float gqy, float knx, float kny, BiasState* st) {
  float pos[2] = {gqx - knx, gqy - kny}, bb[2];
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const float l = log1pf(fabsf(pos[a]));
    bb[a] = pos[a] > 0.f ? l : (pos[a] < 0.f ? -l : 0.f);
  }
  float h1[HID], h2[HID], z1[HID], z2[HID];
#pragma unroll
  for (int u = 0; u < HID; u++) {
    z1[u] = fmaf(p[P_W1 + u * 2], bb[0], fmaf(p[P_W1 + u * 2 + 1], bb[1], p[P_B1 + u]));
    h1[u] = fmaxf(z1[u], 0.f);
  }
  float out = p[P_B3];
#pragma unroll
  for (int u = 0; u < HID; u++) {
    float z = p[P_B2 + u];
#pragma unroll
    for (int t = 0; t < HID; t++) z = fmaf(p[P_W2 + u * HID + t], h1[t], z);
    z2[u] = z;
    h2[u] = fmaxf(z, 0.f);
    out = fmaf(p[P_W3 + u], h2[u], out);
  }
  if (st) {
#pragma unroll
    for (int a = 0; a < 2; a++) { st->pos[a] = pos[a]; st->bb[a] = bb[a]; }
#pragma unroll
    for (int u = 0; u < HID; u++) { st->z1[u] = z1[u]; st->z2[u] = z2[u]; }
  }
  return out;
}

// softmax row of query i over the keys; a[] receives the probabilities (before dropout)
__device__ __forceinline__ void attention_row(const Smem& s, const Dims& d, int i, float* a) {
  const int qy = i / d.cw, qx = i - qy * d.cw;
  // normalize_grid(grid, dim=0): x by (h - 1), y by (w - 1)
  const float gqx = 2.f * (float)qx / (float)max(d.ch - 1, 1) - 1.f, gqy = 2.f * (float)qy / (float)max(d.cw - 1, 1) - 1.f;
  const float scale = rsqrtf((float)INNER);
  float qs[INNER];
#pragma unroll
  for (int o = 0; o < INNER; o++) qs[o] = s.q[o * d.nq + i] * scale;
  float mx = -INFINITY;
  for (int j = 0; j < d.nk; j++) {
    float l = cpb_bias(s.p, gqx, gqy, s.vn[2 * j], s.vn[2 * j + 1], nullptr);
#pragma unroll
    for (int o = 0; o < INNER; o++) l = fmaf(qs[o], s.k[o * d.nk + j], l);
    a[j] = l;
    mx = fmaxf(mx, l);
  }
  float sum = 0.f;
  for (int j = 0; j < d.nk; j++) {
    a[j] = __expf(a[j] - mx);
    sum += a[j];
  }
  const float inv = 1.f / sum;
  for (int j = 0; j < d.nk; j++) a[j] *= inv;
}

__global__ void __launch_bounds__(THREADS) lift_coarse_kernel(Dims d, const float* __restrict__ opacity,
                                                              const float* __restrict__ alpha,
                                                              const float* __restrict__ params,
                                                              const float* __restrict__ keep, float* __restrict__ att) {
  extern __shared__ __align__(16) float smem_l[];
  Smem s;
  carve(s, smem_l, d);
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t vol = (size_t)DIM * d.H * d.W;
  lift_core(s, d, opacity + b * vol, alpha + b * vol, params);
  // (5) logits + position bias, softmax, dropout keep-mask, attn . v, to_out (:195-218)
  for (int i = tid; i < d.nq; i += THREADS) {
    float a[MAXK];
    attention_row(s, d, i, a);
    float out[INNER];
#pragma unroll
    for (int o = 0; o < INNER; o++) out[o] = 0.f;
    for (int j = 0; j < d.nk; j++) {
      const float w = keep ? a[j] * keep[((size_t)b * d.nq + i) * d.nk + j] : a[j];
#pragma unroll
      for (int o = 0; o < INNER; o++) out[o] = fmaf(w, s.v[o * d.nk + j], out[o]);
    }
#pragma unroll
    for (int c = 0; c < DIM; c++) {
      float y = s.p[P_BO + c];
#pragma unroll
      for (int o = 0; o < INNER; o++) y = fmaf(s.p[P_WO + c * INNER + o], out[o], y);
      att[((size_t)b * DIM + c) * d.nq + i] = y;
    }
  }
}

// out = bilinear(att, align_corners=True) + opacity; one thread per 4 consecutive x
__global__ void __launch_bounds__(256) lift_upsample_kernel(Dims d, int planes, const float* __restrict__ att,
                                                            const float* __restrict__ opacity, float* __restrict__ out) {
  const int w4 = (d.W + 3) / 4;
  const size_t total = (size_t)planes * d.H * w4;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int xq4 = (int)(e % w4);
    const size_t r = e / w4;
    const int y = (int)(r % d.H);
    const size_t pl = r / d.H;
    int y0, y1;
    float wy0, wy1;
    ac_tap(y, d.ch, d.H, y0, y1, wy0, wy1);
    const float* a = att + pl * d.nq;
    float res[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int x = xq4 * 4 + k;
      int x0, x1;
      float wx0, wx1;
      ac_tap(min(x, d.W - 1), d.cw, d.W, x0, x1, wx0, wx1);
      res[k] = wy0 * (wx0 * __ldg(a + y0 * d.cw + x0) + wx1 * __ldg(a + y0 * d.cw + x1)) +
               wy1 * (wx0 * __ldg(a + y1 * d.cw + x0) + wx1 * __ldg(a + y1 * d.cw + x1));
    }
    const size_t base = (pl * d.H + y) * d.W + (size_t)xq4 * 4;
    if ((d.W & 3) == 0) {
      const float4 o = __ldg(reinterpret_cast<const float4*>(opacity + base));
      *reinterpret_cast<float4*>(out + base) = make_float4(res[0] + o.x, res[1] + o.y, res[2] + o.z, res[3] + o.w);
    } else {
      for (int k = 0; k < 4 && xq4 * 4 + k < d.W; k++) out[base + k] = res[k] + opacity[base + k];
    }
  }
}

// ---- backward ------------------------------------------------------------------------------------------------------
// adjoint of the up-sampling: g_att[pl][cy][cx] = sum over the fine pixels whose taps include (cy, cx)
__global__ void __launch_bounds__(256) lift_grad_up_kernel(Dims d, const float* __restrict__ g_out,
                                                           float* __restrict__ g_att) {
  extern __shared__ __align__(16) float s_rows[];  // [H][cw]: the plane reduced along x
  const size_t pl = blockIdx.x;
  const float* g = g_out + pl * d.H * d.W;
  // pass 1: along x.  Coarse column cx collects the fine columns x with tap x0 == cx (weight w0) or x1 == cx (w1).
  for (int e = threadIdx.x; e < d.H * d.cw; e += blockDim.x) {
    const int y = e / d.cw, cx = e - y * d.cw;
    // fine columns whose source position lies in (cx - 1, cx + 1)
    const float inv = d.cw > 1 ? (float)(d.W - 1) / (float)(d.cw - 1) : 0.f;
    const int lo = max(0, (int)floorf((cx - 1) * inv)), hi = min(d.W - 1, (int)ceilf((cx + 1) * inv));
    float acc = 0.f;
    for (int x = lo; x <= hi; x++) {
      int x0, x1;
      float w0, w1;
      ac_tap(x, d.cw, d.W, x0, x1, w0, w1);
      const float gv = g[(size_t)y * d.W + x];
      if (x0 == cx) acc = fmaf(w0, gv, acc);
      if (x1 == cx) acc = fmaf(w1, gv, acc);
    }
    s_rows[e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < d.nq; e += blockDim.x) {
    const int cy = e / d.cw, cx = e - cy * d.cw;
    const float inv = d.ch > 1 ? (float)(d.H - 1) / (float)(d.ch - 1) : 0.f;
    const int lo = max(0, (int)floorf((cy - 1) * inv)), hi = min(d.H - 1, (int)ceilf((cy + 1) * inv));
    float acc = 0.f;
    for (int y = lo; y <= hi; y++) {
      int y0, y1;
      float w0, w1;
      ac_tap(y, d.ch, d.H, y0, y1, w0, w1);
      const float gv = s_rows[y * d.cw + cx];
      if (y0 == cy) acc = fmaf(w0, gv, acc);
      if (y1 == cy) acc = fmaf(w1, gv, acc);
    }
    g_att[pl * d.nq + e] = acc;
  }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CTA-wide sum of one value per thread into out[0] (smem scratch [WARPS]); all threads call
__device__ __forceinline__ void cta_sum_to(float v, float* scratch, float* dst_global) {
  v = warp_sum_f(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; w++) t += scratch[w];
    atomicAdd(dst_global, t);
  }
}

__global__ void __launch_bounds__(THREADS) lift_grad_coarse_kernel(
    Dims d, const float* __restrict__ opacity, const float* __restrict__ alpha, const float* __restrict__ params,
    const float* __restrict__ keep, const float* __restrict__ g_att, float* __restrict__ g_xq_out,
    float* __restrict__ g_xkv_out, float* __restrict__ g_params) {
  extern __shared__ __align__(16) float smem_l[];
  Smem s;
  float* f = carve(s, smem_l, d);
  float* A = f;       f += (size_t)d.nq * d.nk;  // attention probabilities after dropout
  float* GS = f;      f += (size_t)d.nq * d.nk;  // dL/d logits
  float* gout = f;    f += INNER * d.nq;         // dL/d (attn . v)
  float* gq = f;      f += INNER * d.nq;         // dL/d q
  float* gk = f;      f += INNER * d.nk;
  float* gv = f;      f += INNER * d.nk;
  float* gkvf = f;    f += DIM * d.nk;
  float* gvn = f;     f += 2 * d.nk;             // dL/d normalised sampling positions
  float* gvn_w = f;   f += (size_t)WARPS * 2 * d.nk;
  float* gdw = f;     f += INNER * d.nk;
  float* go2 = f;     f += 2 * d.nk;
  float* scratch = f; f += 32 * WARPS;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t vol = (size_t)DIM * d.H * d.W;
  lift_core(s, d, opacity + b * vol, alpha + b * vol, params);
  const float* gy = g_att + (size_t)b * DIM * d.nq;
  const float scale = rsqrtf((float)INNER);

  for (int e = tid; e < WARPS * 2 * d.nk; e += THREADS) gvn_w[e] = 0.f;
  __syncthreads();

  // ---- phase A: per query: to_out backward, softmax backward, position-bias MLP backward ----
  float pg[25];  // partial parameter gradients of the bias MLP: W1[6] b1[3] W2[9] b2[3] W3[3] b3[1]
#pragma unroll
  for (int u = 0; u < 25; u++) pg[u] = 0.f;
  const int rounds = (d.nq + THREADS - 1) / THREADS;
  for (int r = 0; r < rounds; r++) {
    const int i = r * THREADS + tid;
    const bool live = i < d.nq;
    float a[MAXK], go[INNER];
    if (live) {
      attention_row(s, d, i, a);
      float gyi[DIM];
#pragma unroll
      for (int c = 0; c < DIM; c++) gyi[c] = gy[c * d.nq + i];
#pragma unroll
      for (int o = 0; o < INNER; o++) {
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < DIM; c++) t = fmaf(s.p[P_WO + c * INNER + o], gyi[c], t);
        go[o] = t;
        gout[o * d.nq + i] = t;
      }
      float rowdot = 0.f;
      float ga[MAXK];
      for (int j = 0; j < d.nk; j++) {
        const float kp = keep ? keep[((size_t)b * d.nq + i) * d.nk + j] : 1.f;
        float t = 0.f;
#pragma unroll
        for (int o = 0; o < INNER; o++) t = fmaf(go[o], s.v[o * d.nk + j], t);
        t *= kp;  // dL/d attn (before dropout)
        ga[j] = t;
        rowdot = fmaf(a[j], t, rowdot);
        A[(size_t)i * d.nk + j] = a[j] * kp;
      }
      for (int j = 0; j < d.nk; j++) GS[(size_t)i * d.nk + j] = a[j] * (ga[j] - rowdot);
    }
    // bias MLP backward; the per-key sums over the queries go through warp reductions into per-warp slots
    const int qy = live ? i / d.cw : 0, qx = live ? i - qy * d.cw : 0;
    const float gqx = 2.f * (float)qx / (float)max(d.ch - 1, 1) - 1.f, gqy = 2.f * (float)qy / (float)max(d.cw - 1, 1) - 1.f;
    for (int j = 0; j < d.nk; j++) {
      float gp0 = 0.f, gp1 = 0.f;
      if (live) {
        BiasState st;
        cpb_bias(s.p, gqx, gqy, s.vn[2 * j], s.vn[2 * j + 1], &st);
        const float g = GS[(size_t)i * d.nk + j];
        float gz2[HID], gh1[HID] = {0.f, 0.f, 0.f}, gz1[HID];
        pg[24] += g;
#pragma unroll
        for (int u = 0; u < HID; u++) {
          pg[21 + u] = fmaf(g, fmaxf(st.z2[u], 0.f), pg[21 + u]);
          gz2[u] = st.z2[u] > 0.f ? g * s.p[P_W3 + u] : 0.f;
          pg[18 + u] += gz2[u];
#pragma unroll
          for (int t = 0; t < HID; t++) {
            pg[9 + u * HID + t] = fmaf(gz2[u], fmaxf(st.z1[t], 0.f), pg[9 + u * HID + t]);
            gh1[t] = fmaf(gz2[u], s.p[P_W2 + u * HID + t], gh1[t]);
          }
        }
        float gbb[2] = {0.f, 0.f};
#pragma unroll
        for (int u = 0; u < HID; u++) {
          gz1[u] = st.z1[u] > 0.f ? gh1[u] : 0.f;
          pg[6 + u] += gz1[u];
          pg[u * 2] = fmaf(gz1[u], st.bb[0], pg[u * 2]);
          pg[u * 2 + 1] = fmaf(gz1[u], st.bb[1], pg[u * 2 + 1]);
          gbb[0] = fmaf(gz1[u], s.p[P_W1 + u * 2], gbb[0]);
          gbb[1] = fmaf(gz1[u], s.p[P_W1 + u * 2 + 1], gbb[1]);
        }
        gp0 = gbb[0] / (fabsf(st.pos[0]) + 1.f);
        gp1 = gbb[1] / (fabsf(st.pos[1]) + 1.f);
      }
      gp0 = warp_sum_f(gp0);
      gp1 = warp_sum_f(gp1);
      if (lane == 0) {  // pos = query - key: the key position receives the negative
        gvn_w[(warp * d.nk + j) * 2] -= gp0;
        gvn_w[(warp * d.nk + j) * 2 + 1] -= gp1;
      }
    }
  }
  __syncthreads();
  {  // parameter gradients of the bias MLP: W1, b1, W2, b2, W3, b3 are contiguous from P_W1
#pragma unroll
    for (int u = 0; u < 25; u++) cta_sum_to(pg[u], scratch, g_params + P_W1 + u);
  }
  for (int e = tid; e < 2 * d.nk; e += THREADS) {
    float t = 0.f;
    for (int w = 0; w < WARPS; w++) t += gvn_w[w * 2 * d.nk + e];
    gvn[e] = t;
  }
  // to_out parameters: dWo[c][o] = sum_i gy[c][i] out[o][i] with out = A . v recomputed; dbo[c] = sum_i gy[c][i]
  for (int e = tid; e < DIM * INNER + DIM; e += THREADS) {
    float acc = 0.f;
    if (e < DIM * INNER) {
      const int c = e / INNER, o = e - c * INNER;
      for (int i = 0; i < d.nq; i++) {
        float out = 0.f;
        for (int j = 0; j < d.nk; j++) out = fmaf(A[(size_t)i * d.nk + j], s.v[o * d.nk + j], out);
        acc = fmaf(gy[c * d.nq + i], out, acc);
      }
      atomicAdd(g_params + P_WO + e, acc);
    } else {
      const int c = e - DIM * INNER;
      for (int i = 0; i < d.nq; i++) acc += gy[c * d.nq + i];
      atomicAdd(g_params + P_BO + c, acc);
    }
  }
  __syncthreads();
  // ---- phase B: reductions over the queries, owned by the thread of the output ----
  for (int e = tid; e < 2 * INNER * d.nk; e += THREADS) {
    const int which = e / (INNER * d.nk), r = e - which * INNER * d.nk;
    const int o = r / d.nk, j = r - o * d.nk;
    float acc = 0.f;
    if (which == 0) {  // dL/dv[o][j] = sum_i A[i][j] gout[o][i]
      for (int i = 0; i < d.nq; i++) acc = fmaf(A[(size_t)i * d.nk + j], gout[o * d.nq + i], acc);
      gv[r] = acc;
    } else {           // dL/dk[o][j] = sum_i GS[i][j] q[o][i] scale
      for (int i = 0; i < d.nq; i++) acc = fmaf(GS[(size_t)i * d.nk + j], s.q[o * d.nq + i], acc);
      gk[r] = acc * scale;
    }
  }
  for (int e = tid; e < INNER * d.nq; e += THREADS) {  // dL/dq (through the logits)
    const int o = e / d.nq, i = e - o * d.nq;
    float acc = 0.f;
    for (int j = 0; j < d.nk; j++) acc = fmaf(GS[(size_t)i * d.nk + j], s.k[o * d.nk + j], acc);
    gq[e] = acc * scale;
  }
  __syncthreads();
  // ---- phase C: key / value path ----
  for (int e = tid; e < 2 * INNER * DIM; e += THREADS) {  // dWk, dWv
    const int which = e / (INNER * DIM), r = e - which * INNER * DIM;
    const int o = r / DIM, c = r - o * DIM;
    const float* gsrc = which ? gv : gk;
    float acc = 0.f;
    for (int j = 0; j < d.nk; j++) acc = fmaf(gsrc[o * d.nk + j], s.kvf[c * d.nk + j], acc);
    atomicAdd(g_params + (which ? P_WV : P_WK) + r, acc);
  }
  for (int e = tid; e < DIM * d.nk; e += THREADS) {
    const int c = e / d.nk, j = e - c * d.nk;
    float acc = 0.f;
#pragma unroll
    for (int o = 0; o < INNER; o++)
      acc = fmaf(s.p[P_WK + o * DIM + c], gk[o * d.nk + j], fmaf(s.p[P_WV + o * DIM + c], gv[o * d.nk + j], acc));
    gkvf[e] = acc;
  }
  __syncthreads();
  // grid_sample backward: positions (one thread per key), then the scatter into x_kv (one thread per channel)
  for (int j = tid; j < d.nk; j += THREADS) {
    const Taps t = sample_taps(s.vn[2 * j], s.vn[2 * j + 1], d.ch, d.cw);
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < DIM; c++) {
      const float g = gkvf[c * d.nk + j];
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          const int y = t.y0 + dy, x = t.x0 + dx;
          if (y < 0 || y >= d.ch || x < 0 || x >= d.cw) continue;
          const float val = s.xkv[c * d.nq + y * d.cw + x] * g;
          gix += val * (dy ? t.fy : 1.f - t.fy) * (dx ? 1.f : -1.f);
          giy += val * (dx ? t.fx : 1.f - t.fx) * (dy ? 1.f : -1.f);
        }
    }
    const float gnx = gvn[2 * j] + gix * (0.5f * d.cw), gny = gvn[2 * j + 1] + giy * (0.5f * d.ch);
    // vn = 2 vgrid / (n - 1) - 1; vgrid = grid + 4 tanh(o2)
    const float gvx = gnx * 2.f / (float)max(d.hk - 1, 1), gvy = gny * 2.f / (float)max(d.wk - 1, 1);
    go2[j] = gvx * OFFSET_SCALE * (1.f - s.th[j] * s.th[j]);
    go2[d.nk + j] = gvy * OFFSET_SCALE * (1.f - s.th[d.nk + j] * s.th[d.nk + j]);
  }
  __syncthreads();
  if (tid < DIM) {  // dL/dx_kv: thread = channel, keys and taps in sequence (taps of different keys may coincide)
    const int c = tid;
    float* gx = g_xkv_out + ((size_t)b * DIM + c) * d.nq;
    for (int i = 0; i < d.nq; i++) gx[i] = 0.f;
    for (int j = 0; j < d.nk; j++) {
      const Taps t = sample_taps(s.vn[2 * j], s.vn[2 * j + 1], d.ch, d.cw);
      const float g = gkvf[c * d.nk + j];
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          const int y = t.y0 + dy, x = t.x0 + dx;
          if (y < 0 || y >= d.ch || x < 0 || x >= d.cw) continue;
          gx[y * d.cw + x] += (dy ? t.fy : 1.f - t.fy) * (dx ? t.fx : 1.f - t.fx) * g;
        }
    }
  }
  // offset network backward
  for (int e = tid; e < 2 * INNER; e += THREADS) {  // dWpw[o][c] = sum_j go2[o][j] ge[c][j]
    const int o = e / INNER, c = e - o * INNER;
    float acc = 0.f;
    for (int j = 0; j < d.nk; j++) acc = fmaf(go2[o * d.nk + j], s.ge[c * d.nk + j], acc);
    atomicAdd(g_params + P_WPW + e, acc);
  }
  for (int e = tid; e < INNER * d.nk; e += THREADS) {
    const int c = e / d.nk, j = e - c * d.nk;
    const float gge = s.p[P_WPW + c] * go2[j] + s.p[P_WPW + INNER + c] * go2[d.nk + j];
    gdw[e] = gge * gelu_grad_f(s.dwv[e]);
  }
  __syncthreads();
  for (int e = tid; e < INNER * KS * KS + INNER; e += THREADS) {  // dWdw[c][ky][kx], dbdw[c]
    float acc = 0.f;
    if (e < INNER * KS * KS) {
      const int c = e / (KS * KS), kk = e - c * KS * KS, ky = kk / KS, kx = kk - ky * KS;
      for (int j = 0; j < d.nk; j++) {
        const int oy = j / d.wk, ox = j - oy * d.wk;
        const int y = oy * DS - PAD + ky, x = ox * DS - PAD + kx;
        if (y < 0 || y >= d.ch || x < 0 || x >= d.cw) continue;
        acc = fmaf(gdw[c * d.nk + j], s.q[c * d.nq + y * d.cw + x], acc);
      }
      atomicAdd(g_params + P_WDW + e, acc);
    } else {
      const int c = e - INNER * KS * KS;
      for (int j = 0; j < d.nk; j++) acc += gdw[c * d.nk + j];
      atomicAdd(g_params + P_BDW + c, acc);
    }
  }
  // dL/dq += offset-conv path (gather over the <= 2 x 2 output positions whose window covers the pixel)
  for (int e = tid; e < INNER * d.nq; e += THREADS) {
    const int c = e / d.nq, i = e - c * d.nq;
    const int y = i / d.cw, x = i - y * d.cw;
    float acc = gq[e];
    for (int oy = max(0, (y + PAD - KS + DS) / DS); oy < d.hk && oy * DS - PAD <= y; oy++) {
      const int ky = y - (oy * DS - PAD);
      if (ky < 0 || ky >= KS) continue;
      for (int ox = max(0, (x + PAD - KS + DS) / DS); ox < d.wk && ox * DS - PAD <= x; ox++) {
        const int kx = x - (ox * DS - PAD);
        if (kx < 0 || kx >= KS) continue;
        acc = fmaf(gdw[c * d.nk + oy * d.wk + ox], s.p[P_WDW + c * KS * KS + ky * KS + kx], acc);
      }
    }
    gq[e] = acc;
  }
  __syncthreads();
  // ---- phase D: to_q backward ----
  for (int e = tid; e < INNER * DIM; e += THREADS) {
    const int o = e / DIM, c = e - o * DIM;
    float acc = 0.f;
    for (int i = 0; i < d.nq; i++) acc = fmaf(gq[o * d.nq + i], s.xq[c * d.nq + i], acc);
    atomicAdd(g_params + P_WQ + e, acc);
  }
  for (int e = tid; e < DIM * d.nq; e += THREADS) {
    const int c = e / d.nq, i = e - c * d.nq;
    float acc = 0.f;
#pragma unroll
    for (int o = 0; o < INNER; o++) acc = fmaf(s.p[P_WQ + o * DIM + c], gq[o * d.nq + i], acc);
    g_xq_out[((size_t)b * DIM + c) * d.nq + i] = acc;
  }
}

// adjoint of the down-sampling, added into a dense gradient: 4 taps per coarse point, never shared between points
__global__ void __launch_bounds__(256) lift_grad_down_kernel(Dims d, int planes, const float* __restrict__ g_coarse,
                                                             float* __restrict__ g_fine) {
  const size_t total = (size_t)planes * d.nq;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t pl = e / d.nq;
    const int i = (int)(e - pl * d.nq);
    const int cy = i / d.cw, cx = i - cy * d.cw;
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    ac_tap(cy, d.H, d.ch, y0, y1, wy0, wy1);
    ac_tap(cx, d.W, d.cw, x0, x1, wx0, wx1);
    const float g = g_coarse[e];
    float* dst = g_fine + pl * d.H * d.W;
    // (at the border the clamped second tap coincides with the first: accumulate in sequence)
    dst[(size_t)y0 * d.W + x0] += wy0 * wx0 * g;
    dst[(size_t)y0 * d.W + x1] += wy0 * wx1 * g;
    dst[(size_t)y1 * d.W + x0] += wy1 * wx0 * g;
    dst[(size_t)y1 * d.W + x1] += wy1 * wx1 * g;
  }
}

static size_t backward_smem_floats(const Dims& d) {
  return smem_forward_floats(d) + 2 * (size_t)d.nq * d.nk + 2 * INNER * d.nq + 2 * INNER * d.nk + DIM * d.nk + 2 * d.nk +
         (size_t)WARPS * 2 * d.nk + INNER * d.nk + 2 * d.nk + 32 * WARPS;
}

}  // namespace hoa
}  // namespace ocrf

using namespace ocrf;
using namespace ocrf::hoa;

static int check_lift_shape(int32_t B, int32_t dim, int32_t H, int32_t W, Dims& d) {
  if (B <= 0 || dim != DIM || H < 36 || W < 36) return OCRF_EINVAL;
  d = make_dims(H, W);
  if (d.hk < 1 || d.wk < 1 || d.nk > MAXK) return OCRF_ECAPACITY;
  return 0;
}

extern "C" size_t ocrf_hoa_lift_workspace_floats(int32_t B, int32_t dim, int32_t H, int32_t W) {
  Dims d;
  if (check_lift_shape(B, dim, H, W, d)) return 0;
  return (size_t)3 * B * DIM * d.nq;  // att | g_xq | g_xkv
}

extern "C" int ocrf_hoa_lift_forward(void* stream, int32_t B, int32_t dim, int32_t H, int32_t W, const float* opacity,
                                     const float* alpha, const float* params, const float* keep, float* out,
                                     float* workspace) {
  if (!opacity || !alpha || !params || !out || !workspace) return OCRF_EINVAL;
  Dims d;
  int rc = check_lift_shape(B, dim, H, W, d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = smem_forward_floats(d) * sizeof(float);
  if (smem > 200 * 1024) return OCRF_ECAPACITY;
  static unsigned long long attr = 0;
  cudaError_t e = ensure_dynamic_smem(lift_coarse_kernel, smem, attr);
  if (e != cudaSuccess) return (int)e;
  float* att = workspace;
  lift_coarse_kernel<<<B, THREADS, smem, st>>>(d, opacity, alpha, params, keep, att);
  const size_t work = (size_t)B * DIM * H * ((W + 3) / 4);
  const unsigned grid = (unsigned)min((size_t)num_sms() * 8, (work + 255) / 256);
  lift_upsample_kernel<<<grid, 256, 0, st>>>(d, B * DIM, att, opacity, out);
  OCRF_CHECK_LAST();
  return 0;
}

extern "C" int ocrf_hoa_lift_backward(void* stream, int32_t B, int32_t dim, int32_t H, int32_t W, const float* opacity,
                                      const float* alpha, const float* params, const float* keep, const float* g_out,
                                      float* g_opacity, float* g_alpha, float* g_params, float* workspace) {
  if (!opacity || !alpha || !params || !g_out || !g_opacity || !g_alpha || !g_params || !workspace) return OCRF_EINVAL;
  Dims d;
  int rc = check_lift_shape(B, dim, H, W, d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = backward_smem_floats(d) * sizeof(float);
  if (smem > 227 * 1024) return OCRF_ECAPACITY;
  static unsigned long long attr = 0, attr_up = 0;
  cudaError_t e = ensure_dynamic_smem(lift_grad_coarse_kernel, smem, attr);
  if (e != cudaSuccess) return (int)e;
  const size_t smem_up = (size_t)H * d.cw * sizeof(float);
  e = ensure_dynamic_smem(lift_grad_up_kernel, smem_up, attr_up);
  if (e != cudaSuccess) return (int)e;
  const size_t n_coarse = (size_t)B * DIM * d.nq;
  float* g_att = workspace;
  float* g_xq = workspace + n_coarse;
  float* g_xkv = workspace + 2 * n_coarse;
  const size_t bytes = (size_t)B * DIM * H * W * sizeof(float);
  lift_grad_up_kernel<<<B * DIM, 256, smem_up, st>>>(d, g_out, g_att);
  lift_grad_coarse_kernel<<<B, THREADS, smem, st>>>(d, opacity, alpha, params, keep, g_att, g_xq, g_xkv, g_params);
  // residual: d/d opacity starts as g_out, d/d alpha as zero; then the sparse taps of the two down-samplings
  e = cudaMemcpyAsync(g_opacity, g_out, bytes, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(g_alpha, 0, bytes, st);
  if (e != cudaSuccess) return (int)e;
  const unsigned grid = (unsigned)((n_coarse + 255) / 256);
  lift_grad_down_kernel<<<grid, 256, 0, st>>>(d, B * DIM, g_xq, g_opacity);
  lift_grad_down_kernel<<<grid, 256, 0, st>>>(d, B * DIM, g_xkv, g_alpha);
  OCRF_CHECK_LAST();
  return 0;
}
