// Scope row f-4 ("next", HOA remainder / the callers upstream of the Gaussian heads): voxel colouring and the sparse
// supervision image, both from the projected voxel centres.
//
// (1) ocrf_color_voxels replaces lidar_points_to_image_values + color_voxels
//     (/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:924-971, called at :1068-1072): every voxel's
//     colour = mean over the cameras that see it of a bilinear image sample (F.grid_sample, align_corners=True, zero
//     padding).  The reference materialises ~15 tensors of [B,6,13,16384,3] per step (clone, two normalisations,
//     grid_sample, permute, mask multiply, where, sum, count, where, divide, zeros, two slice copies); here one thread
//     per voxel walks its cameras, reads the projected coordinate and the mask once (coalesced), gathers the 4 x C taps
//     and writes the mean: 10 B read per (camera, voxel) + the taps, 12 B written per voxel.
// (2) ocrf_retain_valid_pixels replaces retain_valid_pixels (:1004-1022, called at :1077), which loops
//     B x 6 x 13 times in Python over boolean-mask indexing (~4 000 kernel launches per step): the result is 255
//     everywhere except at the pixels some visible voxel centre truncates to, which keep the image value.  Two
//     launches: fill, then one thread per (camera, voxel) copies its pixel's channels (writes of equal values commute).
// Neither op has a gradient in the reference (images and projected coordinates are inputs).
#include "common.cuh"

namespace ocrf {

// F.grid_sample(bilinear, zeros, align_corners=True) of one point, with the caller's normalisation in front of it:
// xn = x / (W - 1) * 2 - 1 (view_transformer_ocrf.py:929-930), ix = (xn + 1) / 2 * (W - 1) (ATen GridSampler.h
// grid_sampler_unnormalize), corner weights and accumulation order as in ATen's grid_sampler_2d kernel.
template <int MAXC>
__device__ __forceinline__ void bilinear_sample(const float* __restrict__ img, int C, int H, int W, float x, float y,
                                                float* acc) {
  const float xn = __fsub_rn(__fmul_rn(__fdiv_rn(x, (float)(W - 1)), 2.f), 1.f);
  const float yn = __fsub_rn(__fmul_rn(__fdiv_rn(y, (float)(H - 1)), 2.f), 1.f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(xn, 1.f), 2.f), (float)(W - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(yn, 1.f), 2.f), (float)(H - 1));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
  const float w_nw = __fmul_rn(fx1 - ix, fy1 - iy), w_ne = __fmul_rn(ix - fx0, fy1 - iy);
  const float w_sw = __fmul_rn(fx1 - ix, iy - fy0), w_se = __fmul_rn(ix - fx0, iy - fy0);
  // out-of-range float -> int conversions are guarded by the comparisons on the float values
  const bool x0_ok = fx0 >= 0.f && fx0 <= (float)(W - 1), x1_ok = fx1 >= 0.f && fx1 <= (float)(W - 1);
  const bool y0_ok = fy0 >= 0.f && fy0 <= (float)(H - 1), y1_ok = fy1 >= 0.f && fy1 <= (float)(H - 1);
  const int x0 = x0_ok ? (int)fx0 : 0, x1 = x1_ok ? (int)fx1 : 0, y0 = y0_ok ? (int)fy0 : 0, y1 = y1_ok ? (int)fy1 : 0;
  const size_t HW = (size_t)H * W;
#pragma unroll
  for (int c = 0; c < MAXC; c++) {
    if (c >= C) break;
    const float* pl = img + c * HW;
    float v = 0.f;
    if (x0_ok && y0_ok) v = __fmul_rn(__ldg(pl + (size_t)y0 * W + x0), w_nw);
    if (x1_ok && y0_ok) v = __fadd_rn(v, __fmul_rn(__ldg(pl + (size_t)y0 * W + x1), w_ne));
    if (x0_ok && y1_ok) v = __fadd_rn(v, __fmul_rn(__ldg(pl + (size_t)y1 * W + x0), w_sw));
    if (x1_ok && y1_ok) v = __fadd_rn(v, __fmul_rn(__ldg(pl + (size_t)y1 * W + x1), w_se));
    acc[c] = v;
  }
}

constexpr int VC_MAXC = 4;

__global__ void __launch_bounds__(256) color_voxels_kernel(int B, int N, long long M, int C, int H, int W,
                                                           const float2* __restrict__ coords /*[B,N,M]*/,
                                                           const uint8_t* __restrict__ mask /*[B,N,M]*/,
                                                           const float* __restrict__ imgs /*[B,N,C,H,W]*/, float divisor,
                                                           float* __restrict__ avg /*[B,M,C]*/,
                                                           uint8_t* __restrict__ valid /*[B,M] or null*/) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (m >= M) return;
  float sum[VC_MAXC] = {0.f, 0.f, 0.f, 0.f};
  int count = 0;
  for (int n = 0; n < N; n++) {
    const size_t o = ((size_t)b * N + n) * M + m;
    if (!__ldg(mask + o)) continue;
    const float2 p = __ldg(coords + o);
    float s[VC_MAXC] = {0.f, 0.f, 0.f, 0.f};
    bilinear_sample<VC_MAXC>(imgs + ((size_t)b * N + n) * C * H * W, C, H, W, p.x, p.y, s);
#pragma unroll
    for (int c = 0; c < VC_MAXC; c++) sum[c] = __fadd_rn(sum[c], s[c]);  // cameras in index order
    count++;
  }
  const float denom = (float)(count > 0 ? count : 1);
#pragma unroll
  for (int c = 0; c < VC_MAXC; c++)
    if (c < C) {
      float v = __fdiv_rn(sum[c], denom);
      if (divisor != 1.f) v = __fdiv_rn(v, divisor);
      avg[((size_t)b * M + m) * C + c] = v;
    }
  if (valid) valid[(size_t)b * M + m] = count > 0;
}

__global__ void __launch_bounds__(256) fill_value_kernel(size_t n4, size_t n, float value, float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4* o4 = reinterpret_cast<float4*>(out);
  for (size_t i = t0; i < n4; i += stride) o4[i] = make_float4(value, value, value, value);
  for (size_t i = n4 * 4 + t0; i < n; i += stride) out[i] = value;
}

__global__ void __launch_bounds__(256) retain_valid_pixels_kernel(long long M, int C, int H, int W,
                                                                  const float2* __restrict__ coords /*[V,M]*/,
                                                                  const uint8_t* __restrict__ mask /*[V,M]*/,
                                                                  const float* __restrict__ img /*[V,C,H,W]*/,
                                                                  float* __restrict__ out) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (m >= M) return;
  const size_t o = (size_t)v * M + m;
  if (!__ldg(mask + o)) return;
  const float2 p = __ldg(coords + o);
  if (p.x == -1.f) return;  // the reference marks dropped points with -1 and tests x only (:1011,1017)
  // .long() truncates toward zero; both coordinates are clamped to [0, max(W, H) - 1] (:1018)
  const long long lim = (long long)max(W, H) - 1;
  long long xi = (long long)p.x, yi = (long long)p.y;
  xi = xi < 0 ? 0 : (xi > lim ? lim : xi);
  yi = yi < 0 ? 0 : (yi > lim ? lim : yi);
  if (xi >= W || yi >= H) return;  // torch would raise an index error here; such points are not produced by the caller
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)yi * W + xi;
  for (int c = 0; c < C; c++) {
    const size_t idx = ((size_t)v * C + c) * HW + pix;
    out[idx] = __ldg(img + idx);
  }
}

}  // namespace ocrf

using namespace ocrf;

extern "C" int ocrf_color_voxels(void* stream, int32_t B, int32_t N, int64_t M, int32_t C, int32_t H, int32_t W,
                                 const float* coords, const uint8_t* mask, const float* imgs, float divisor, float* avg,
                                 uint8_t* valid) {
  if (B < 0 || N <= 0 || M < 0 || C <= 0 || C > VC_MAXC || H < 2 || W < 2) return OCRF_EINVAL;
  if (B == 0 || M == 0) return 0;
  if (!coords || !mask || !imgs || !avg) return OCRF_EINVAL;
  if (reinterpret_cast<uintptr_t>(coords) & 7) return OCRF_EINVAL;
  const dim3 grid((unsigned)((M + 255) / 256), (unsigned)B);
  color_voxels_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      B, N, M, C, H, W, reinterpret_cast<const float2*>(coords), mask, imgs, divisor, avg, valid);
  OCRF_CHECK_LAST();
  return 0;
}

extern "C" int ocrf_retain_valid_pixels(void* stream, int32_t V, int64_t M, int32_t C, int32_t H, int32_t W,
                                        const float* coords, const uint8_t* mask, const float* img, float fill,
                                        float* out) {
  if (V < 0 || M < 0 || C <= 0 || H <= 0 || W <= 0) return OCRF_EINVAL;
  if (V == 0) return 0;
  if (!img || !out || (M > 0 && (!coords || !mask))) return OCRF_EINVAL;
  if ((reinterpret_cast<uintptr_t>(coords) & 7) || (reinterpret_cast<uintptr_t>(out) & 15)) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)V * C * H * W;
  const size_t want = (n / 4 + 255) / 256;
  fill_value_kernel<<<(unsigned)(want < (size_t)num_sms() * 8 ? (want ? want : 1) : (size_t)num_sms() * 8), 256, 0, st>>>(
      n / 4, n, fill, out);
  if (M > 0) {
    const dim3 grid((unsigned)((M + 255) / 256), (unsigned)V);
    retain_valid_pixels_kernel<<<grid, 256, 0, st>>>(M, C, H, W, reinterpret_cast<const float2*>(coords), mask, img, out);
  }
  OCRF_CHECK_LAST();
  return 0;
}
