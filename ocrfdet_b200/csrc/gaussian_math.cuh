// Per-Gaussian projection math shared by the forward and backward preprocess kernels.
// Everything that feeds the sort keys is written with explicit round-to-nearest intrinsics in the
// contraction pattern nvcc produces for the reference's expressions (see preprocess.cu header).
#pragma once
#include "common.cuh"

namespace ocrf {

__device__ __forceinline__ float dot3(float a, float b, float c, float d, float e, float f) {
  return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
  return __fadd_rn(__fmaf_rn(m[8 + r], z, __fmaf_rn(m[r], x, __fmul_rn(m[4 + r], y))), m[12 + r]);
}

__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod, float4 q, float* out) {
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
  const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  float R[3][3];  // R[c][k]: column c, row k
  R[0][0] = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(yy, zz)));
  R[0][1] = __fmul_rn(2.f, __fmaf_rn(x, y, -rz));
  R[0][2] = __fmul_rn(2.f, __fmaf_rn(r, y, xz));
  R[1][0] = __fmul_rn(2.f, __fmaf_rn(x, y, rz));
  R[1][1] = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, zz)));
  R[1][2] = __fmul_rn(2.f, __fmaf_rn(y, z, -rx));
  R[2][0] = __fmul_rn(2.f, __fmaf_rn(-r, y, xz));
  R[2][1] = __fmul_rn(2.f, __fmaf_rn(y, z, rx));
  R[2][2] = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, yy)));
  const float s[3] = {__fmul_rn(sx, mod), __fmul_rn(sy, mod), __fmul_rn(sz, mod)};
  float M[3][3];
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) M[c][k] = __fmul_rn(s[k], R[c][k]);
#define OCRF_SIG(c, k) dot3(M[k][0], M[c][0], M[k][1], M[c][1], M[k][2], M[c][2])
  out[0] = OCRF_SIG(0, 0);
  out[1] = OCRF_SIG(0, 1);
  out[2] = OCRF_SIG(0, 2);
  out[3] = OCRF_SIG(1, 1);
  out[4] = OCRF_SIG(1, 2);
  out[5] = OCRF_SIG(2, 2);
#undef OCRF_SIG
}

__device__ __forceinline__ void cov2d_ewa(float tx, float ty, float tz, float fx, float fy, float tanx, float tany,
                                          const float* c6, const float* v, float& a, float& b, float& c) {
  const float limx = __fmul_rn(1.3f, tanx), limy = __fmul_rn(1.3f, tany);
  const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
  const float cx = fminf(limx, fmaxf(-limx, txtz)), cy = fminf(limy, fmaxf(-limy, tytz));
  const float J00 = __fdiv_rn(fx, tz), J11 = __fdiv_rn(fy, tz);
  const float tz2 = __fmul_rn(tz, tz);
  const float J02 = __fdiv_rn(__fmul_rn(-__fmul_rn(tz, cx), fx), tz2);
  const float J12 = __fdiv_rn(__fmul_rn(-__fmul_rn(tz, cy), fy), tz2);
  float T0[3], T1[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    T0[k] = __fmaf_rn(v[4 * k + 2], J02, __fmaf_rn(v[4 * k + 0], J00, __fmul_rn(0.f, v[4 * k + 1])));
    T1[k] = __fmaf_rn(v[4 * k + 2], J12, __fmaf_rn(0.f, v[4 * k + 0], __fmul_rn(v[4 * k + 1], J11)));
  }
  const float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
  float A0[3], A1[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    A0[k] = dot3(T0[0], S[0][k], T0[1], S[1][k], T0[2], S[2][k]);
    A1[k] = dot3(T1[0], S[0][k], T1[1], S[1][k], T1[2], S[2][k]);
  }
  a = __fadd_rn(dot3(T0[0], A0[0], T0[1], A0[1], T0[2], A0[2]), 0.3f);
  b = dot3(T0[0], A1[0], T0[1], A1[1], T0[2], A1[2]);
  c = __fadd_rn(dot3(T1[0], A1[0], T1[1], A1[1], T1[2], A1[2]), 0.3f);
}

__device__ __forceinline__ float ndc_to_pix(float v, int S) {
  return __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5));
}

__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1,
                                          int& y1) {
  const float r = (float)radius;
  x0 = min(gx, max(0, (int)__fmul_rn(__fsub_rn(px, r), 0.0625f)));
  y0 = min(gy, max(0, (int)__fmul_rn(__fsub_rn(py, r), 0.0625f)));
  x1 = min(gx, max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(px, r), 16.f), 1.f), 0.0625f)));
  y1 = min(gy, max(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(py, r), 16.f), 1.f), 0.0625f)));
}

// SH basis constants (auxiliary.h:20-38 of the reference)
#define OCRF_SH_C0 0.28209479177387814f
#define OCRF_SH_C1 0.4886025119029199f
#define OCRF_SH_C2_0 1.0925484305920792f
#define OCRF_SH_C2_1 (-1.0925484305920792f)
#define OCRF_SH_C2_2 0.31539156525252005f
#define OCRF_SH_C2_3 (-1.0925484305920792f)
#define OCRF_SH_C2_4 0.5462742152960396f
#define OCRF_SH_C3_0 (-0.5900435899266435f)
#define OCRF_SH_C3_1 2.890611442640554f
#define OCRF_SH_C3_2 (-0.4570457994644658f)
#define OCRF_SH_C3_3 0.3731763325901154f
#define OCRF_SH_C3_4 (-0.4570457994644658f)
#define OCRF_SH_C3_5 1.445305721320277f
#define OCRF_SH_C3_6 (-0.5900435899266435f)

// SH basis up to degree 3 (forward.cu:31-60)
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
  const float C0 = OCRF_SH_C0, C1 = OCRF_SH_C1;
  b[0] = C0;
  if (deg < 1) return;
  b[1] = -C1 * y; b[2] = C1 * z; b[3] = -C1 * x;
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  b[4] = OCRF_SH_C2_0 * xy; b[5] = OCRF_SH_C2_1 * yz; b[6] = OCRF_SH_C2_2 * (2.f * zz - xx - yy);
  b[7] = OCRF_SH_C2_3 * xz; b[8] = OCRF_SH_C2_4 * (xx - yy);
  if (deg < 3) return;
  b[9] = OCRF_SH_C3_0 * y * (3.f * xx - yy); b[10] = OCRF_SH_C3_1 * xy * z; b[11] = OCRF_SH_C3_2 * y * (4.f * zz - xx - yy);
  b[12] = OCRF_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy); b[13] = OCRF_SH_C3_4 * x * (4.f * zz - xx - yy);
  b[14] = OCRF_SH_C3_5 * z * (xx - yy); b[15] = OCRF_SH_C3_6 * x * (xx - 3.f * yy);
}


}  // namespace ocrf
