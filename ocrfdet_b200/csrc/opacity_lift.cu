// Stage 5: the memory-bound part of the height-aware opacity (HOA) lift -- the opacity mask that
// gates the BEV geometry feature.
//
// Replaces `ObatinOpacityMask.forward` and its application
// (/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:230-242, 1197-1199):
//     s = cat(mean_c(x), max_c(x));  mask = sigmoid(conv7x7(s) + opacity_bev);  out = x * mask
// which in the reference is five torch ops (mean, max, cat, conv2d, sigmoid-add, mul), each a full
// pass over [B,80,128,128].  Here: one pass for the channel statistics, one fused pass for
// conv + sigmoid + gating (x comes from L2 the second time: 42 MB at B=8 fits the 126 MB L2).
// All accesses are pixel-contiguous (coalesced 128-byte lines per channel plane).
#include "common.cuh"

namespace ocrf {

__global__ void __launch_bounds__(256) opacity_stats_kernel(int C, int HW, const float* __restrict__ x,
                                                            float* __restrict__ stats) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* xb = x + (size_t)b * C * HW + p;
  float sum = 0.f, mx = -INFINITY;
#pragma unroll 8
  for (int c = 0; c < C; c++) {
    const float v = __ldg(xb + (size_t)c * HW);
    sum += v;
    mx = fmaxf(mx, v);
  }
  stats[((size_t)b * 2) * HW + p] = sum / (float)C;
  stats[((size_t)b * 2 + 1) * HW + p] = mx;
}

__global__ void __launch_bounds__(256) opacity_mask_apply_kernel(int C, int H, int W, int K,
                                                                 const float* __restrict__ x,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ opacity_bev,
                                                                 const float* __restrict__ stats,
                                                                 float* __restrict__ out, float* __restrict__ mask) {
  extern __shared__ float s_w[];  // [2*K*K]
  for (int i = threadIdx.x; i < 2 * K * K; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int HW = H * W, b = blockIdx.y, pad = K / 2;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int yy = p / W, xx = p - yy * W;
  const float* st = stats + (size_t)b * 2 * HW;
  float acc = 0.f;
  for (int ch = 0; ch < 2; ch++)
    for (int ky = 0; ky < K; ky++) {
      const int sy = yy + ky - pad;
      if (sy < 0 || sy >= H) continue;
      for (int kx = 0; kx < K; kx++) {
        const int sx = xx + kx - pad;
        if (sx < 0 || sx >= W) continue;
        acc += s_w[(ch * K + ky) * K + kx] * __ldg(st + (size_t)ch * HW + sy * W + sx);
      }
    }
  const float m = 1.f / (1.f + __expf(-(acc + opacity_bev[(size_t)b * HW + p])));
  mask[(size_t)b * HW + p] = m;
  const float* xb = x + (size_t)b * C * HW + p;
  float* ob = out + (size_t)b * C * HW + p;
#pragma unroll 8
  for (int c = 0; c < C; c++) ob[(size_t)c * HW] = __ldg(xb + (size_t)c * HW) * m;
}

// backward 1: gz = (sum_c g_out * x) * m (1 - m)  -> scratch[0], also g_opacity_bev
__global__ void __launch_bounds__(256) opacity_bwd_gz_kernel(int C, int HW, const float* __restrict__ x,
                                                             const float* __restrict__ mask,
                                                             const float* __restrict__ g_out,
                                                             float* __restrict__ gz, float* __restrict__ g_opacity_bev) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* xb = x + (size_t)b * C * HW + p;
  const float* gb = g_out + (size_t)b * C * HW + p;
  float gm = 0.f;
#pragma unroll 8
  for (int c = 0; c < C; c++) gm += __ldg(gb + (size_t)c * HW) * __ldg(xb + (size_t)c * HW);
  const float m = mask[(size_t)b * HW + p];
  const float g = gm * m * (1.f - m);
  gz[(size_t)b * HW + p] = g;
  g_opacity_bev[(size_t)b * HW + p] = g;
}

// backward 2: gradient of the conv wrt its input (g_stats) and its weights (g_w, accumulated)
__global__ void __launch_bounds__(256) opacity_bwd_conv_kernel(int H, int W, int K, const float* __restrict__ w,
                                                               const float* __restrict__ stats,
                                                               const float* __restrict__ gz,
                                                               float* __restrict__ g_stats, float* __restrict__ g_w) {
  extern __shared__ float s_mem[];  // [2*K*K] weights, [2*K*K] weight-gradient partials
  float* s_w = s_mem;
  float* s_gw = s_mem + 2 * K * K;
  for (int i = threadIdx.x; i < 2 * K * K; i += blockDim.x) {
    s_w[i] = w[i];
    s_gw[i] = 0.f;
  }
  __syncthreads();
  const int HW = H * W, b = blockIdx.y, pad = K / 2;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < HW;
  const int yy = valid ? p / W : 0, xx = valid ? p - yy * W : 0;
  const float* st = stats + (size_t)b * 2 * HW;
  const float* gzb = gz + (size_t)b * HW;
  const float g_here = valid ? gzb[p] : 0.f;
  const int lane = threadIdx.x & 31;
  for (int ch = 0; ch < 2; ch++) {
    float gs = 0.f;
    for (int ky = 0; ky < K; ky++)
      for (int kx = 0; kx < K; kx++) {
        // d out(p) / d w[ch,ky,kx] = stats[ch, p + (k - pad)]
        const int sy = yy + ky - pad, sx = xx + kx - pad;
        float contrib = 0.f;
        if (valid && sy >= 0 && sy < H && sx >= 0 && sx < W) contrib = g_here * __ldg(st + (size_t)ch * HW + sy * W + sx);
        contrib = warp_sum(contrib);
        if (lane == 0 && contrib != 0.f) atomicAdd(&s_gw[(ch * K + ky) * K + kx], contrib);
        // d out(q) / d stats[ch, p] with q = p - (k - pad)
        const int qy = yy - (ky - pad), qx = xx - (kx - pad);
        if (valid && qy >= 0 && qy < H && qx >= 0 && qx < W) gs += __ldg(gzb + qy * W + qx) * s_w[(ch * K + ky) * K + kx];
      }
    if (valid) g_stats[((size_t)b * 2 + ch) * HW + p] = gs;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * K * K; i += blockDim.x)
    if (s_gw[i] != 0.f) atomicAdd(&g_w[i], s_gw[i]);
}

// backward 3: g_x = g_out * m + g_mean / C + [c == argmax] g_max
__global__ void __launch_bounds__(256) opacity_bwd_gx_kernel(int C, int HW, const float* __restrict__ x,
                                                             const float* __restrict__ mask,
                                                             const float* __restrict__ g_out,
                                                             const float* __restrict__ g_stats,
                                                             float* __restrict__ g_x) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* xb = x + (size_t)b * C * HW + p;
  int arg = 0;
  float mx = -INFINITY;
  for (int c = 0; c < C; c++) {
    const float v = __ldg(xb + (size_t)c * HW);
    if (v > mx) { mx = v; arg = c; }
  }
  const float m = mask[(size_t)b * HW + p];
  const float gmean = g_stats[((size_t)b * 2) * HW + p] / (float)C;
  const float gmax = g_stats[((size_t)b * 2 + 1) * HW + p];
  const float* gb = g_out + (size_t)b * C * HW + p;
  float* gx = g_x + (size_t)b * C * HW + p;
#pragma unroll 4
  for (int c = 0; c < C; c++) gx[(size_t)c * HW] = __ldg(gb + (size_t)c * HW) * m + gmean + (c == arg ? gmax : 0.f);
}

}  // namespace ocrf

using namespace ocrf;

extern "C" int ocrf_opacity_mask_forward(void* stream, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K,
                                         const float* x, const float* w, const float* opacity_bev, float* out,
                                         float* mask, float* stats) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0 || (K & 1) == 0 || K > 31) return OCRF_EINVAL;
  if (!x || !w || !opacity_bev || !out || !mask || !stats) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W;
  const dim3 grid(ceil_div(HW, 256), B);
  opacity_stats_kernel<<<grid, 256, 0, st>>>(C, HW, x, stats);
  opacity_mask_apply_kernel<<<grid, 256, 2 * K * K * sizeof(float), st>>>(C, H, W, K, x, w, opacity_bev, stats, out, mask);
  OCRF_CHECK_LAST();
  return 0;
}

extern "C" int ocrf_opacity_mask_backward(void* stream, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K,
                                          const float* x, const float* w, const float* mask, const float* stats,
                                          const float* g_out, float* g_x, float* g_w, float* g_opacity_bev,
                                          float* scratch) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0 || (K & 1) == 0 || K > 31) return OCRF_EINVAL;
  if (!x || !w || !mask || !stats || !g_out || !g_x || !g_w || !g_opacity_bev || !scratch) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W;
  const dim3 grid(ceil_div(HW, 256), B);
  float* gz = scratch;
  float* g_stats = scratch + (size_t)B * HW;
  opacity_bwd_gz_kernel<<<grid, 256, 0, st>>>(C, HW, x, mask, g_out, gz, g_opacity_bev);
  opacity_bwd_conv_kernel<<<grid, 256, 4 * K * K * sizeof(float), st>>>(H, W, K, w, stats, gz, g_stats, g_w);
  opacity_bwd_gx_kernel<<<grid, 256, 0, st>>>(C, HW, x, mask, g_out, g_stats, g_x);
  OCRF_CHECK_LAST();
  return 0;
}
