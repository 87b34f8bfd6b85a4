// Backward of stage 1: screen-space gradients -> Gaussian parameters.
//
// Replaces computeCov2DCUDA (cuda_rasterizer/backward.cu:144-274), preprocessCUDA backward
// (:346-396), computeCov3D backward (:278-341) and the SH backward (:20-139), which the reference
// runs as two kernels per view.  Here ONE kernel handles a Gaussian for ALL views of its sample:
// the thread recomputes the 3D covariance once (instead of re-reading 24 B per view), walks the
// sample's cameras, accumulates dL/dmean3D, dL/dSigma and dL/dopacity in registers and writes each
// output exactly once -- no atomics, deterministic, and the sum over views that a per-view API
// leaves to autograd happens in registers.
//
// Precision: the conic -> covariance -> quaternion chain cancels heavily in float32 (the reference's
// (denom - a*c) is -b*b; two float32 builds of the same formulas differ by ~1e-4 of the largest
// gradient).  This kernel is P-proportional and B200 has FP64 to spare, so the chain runs in double:
// ~2x the ALU time of a kernel that is bound by its 100 B/Gaussian of HBM traffic anyway.
#include "common.cuh"
#include "gaussian_math.cuh"

namespace ocrf {

// d(basis)/d(x,y,z), matching sh_basis
__device__ __forceinline__ void sh_basis_grad(int deg, float x, float y, float z, float (*g)[3]) {
#pragma unroll
  for (int i = 0; i < 16; i++) g[i][0] = g[i][1] = g[i][2] = 0.f;
  if (deg < 1) return;
  g[1][1] = -OCRF_SH_C1; g[2][2] = OCRF_SH_C1; g[3][0] = -OCRF_SH_C1;
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  g[4][0] = OCRF_SH_C2_0 * y; g[4][1] = OCRF_SH_C2_0 * x;
  g[5][1] = OCRF_SH_C2_1 * z; g[5][2] = OCRF_SH_C2_1 * y;
  g[6][0] = OCRF_SH_C2_2 * -2.f * x; g[6][1] = OCRF_SH_C2_2 * -2.f * y; g[6][2] = OCRF_SH_C2_2 * 4.f * z;
  g[7][0] = OCRF_SH_C2_3 * z; g[7][2] = OCRF_SH_C2_3 * x;
  g[8][0] = OCRF_SH_C2_4 * 2.f * x; g[8][1] = OCRF_SH_C2_4 * -2.f * y;
  if (deg < 3) return;
  g[9][0] = OCRF_SH_C3_0 * 6.f * xy; g[9][1] = OCRF_SH_C3_0 * 3.f * (xx - yy);
  g[10][0] = OCRF_SH_C3_1 * yz; g[10][1] = OCRF_SH_C3_1 * xz; g[10][2] = OCRF_SH_C3_1 * xy;
  g[11][0] = OCRF_SH_C3_2 * -2.f * xy; g[11][1] = OCRF_SH_C3_2 * (-3.f * yy + 4.f * zz - xx); g[11][2] = OCRF_SH_C3_2 * 8.f * yz;
  g[12][0] = OCRF_SH_C3_3 * -6.f * xz; g[12][1] = OCRF_SH_C3_3 * -6.f * yz; g[12][2] = OCRF_SH_C3_3 * 3.f * (2.f * zz - xx - yy);
  g[13][0] = OCRF_SH_C3_4 * (-3.f * xx + 4.f * zz - yy); g[13][1] = OCRF_SH_C3_4 * -2.f * xy; g[13][2] = OCRF_SH_C3_4 * 8.f * xz;
  g[14][0] = OCRF_SH_C3_5 * 2.f * xz; g[14][1] = OCRF_SH_C3_5 * -2.f * yz; g[14][2] = OCRF_SH_C3_5 * (xx - yy);
  g[15][0] = OCRF_SH_C3_6 * 3.f * (xx - yy); g[15][1] = OCRF_SH_C3_6 * -6.f * xy;
}

// Sigma = A^T diag(s^2) A in double (the forward's float value is only needed bit-exactly for the keys)
__device__ __forceinline__ void cov3d_f64(const float* sc, float mod, float4 q, double* out) {
  const double r = q.x, x = q.y, y = q.z, z = q.w;
  const double A[3][3] = {{1. - 2. * (y * y + z * z), 2. * (x * y + r * z), 2. * (x * z - r * y)},
                          {2. * (x * y - r * z), 1. - 2. * (x * x + z * z), 2. * (y * z + r * x)},
                          {2. * (x * z + r * y), 2. * (y * z - r * x), 1. - 2. * (x * x + y * y)}};
  double s2[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { const double s = (double)mod * sc[k]; s2[k] = s * s; }
  int e = 0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) out[e++] = s2[0] * A[0][i] * A[0][j] + s2[1] * A[1][i] * A[1][j] + s2[2] * A[2][i] * A[2][j];
}

constexpr int PB_THREADS = 128;  // ~160 registers (fp64 chain): 3 CTAs of 128 per SM instead of 1 of 256

__global__ void __launch_bounds__(PB_THREADS) preprocess_backward_kernel(
    OcrfShape sh, const float* __restrict__ means3D, const float* __restrict__ scales,
    const float* __restrict__ rotations, const float* __restrict__ cov3D_precomp, const float* __restrict__ shs,
    const Camera* __restrict__ cams, float scale_modifier, const int32_t* __restrict__ radii,
    const uint8_t* __restrict__ clamped, const double* __restrict__ ggrad, const float* __restrict__ dL_dcolors_view,
    float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dopacities,
    float* __restrict__ dL_dscales, float* __restrict__ dL_drotations, float* __restrict__ dL_dcov3D,
    float* __restrict__ dL_dshs) {
  pdl_enter();
  extern __shared__ Camera s_cams[];  // the sample's cameras
  const int s = blockIdx.y;
  const int vps = sh.views_per_sample;
  for (int i = threadIdx.x; i < vps * OCRF_CAM_STRIDE; i += blockDim.x)
    reinterpret_cast<float*>(s_cams)[i] = reinterpret_cast<const float*>(cams + (size_t)s * vps)[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sh.P) return;
  const size_t gi = (size_t)s * sh.P + i;
  const double x = means3D[3 * gi], y = means3D[3 * gi + 1], z = means3D[3 * gi + 2];

  double c6[6];
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  float sc[3] = {0.f, 0.f, 0.f};
  if (cov3D_precomp != nullptr) {
#pragma unroll
    for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[gi * 6 + k];
  } else {
    q = *reinterpret_cast<const float4*>(rotations + gi * 4);
    sc[0] = scales[3 * gi]; sc[1] = scales[3 * gi + 1]; sc[2] = scales[3 * gi + 2];
    float c6f[6];  // the forward's own float32 covariance: gradients are taken at the point the forward evaluated
    cov3d_from_scale_rot(sc[0], sc[1], sc[2], scale_modifier, q, c6f);
#pragma unroll
    for (int k = 0; k < 6; k++) c6[k] = c6f[k];
  }
  const double S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};

  double gmean[3] = {0., 0., 0.}, gcov[6] = {0., 0., 0., 0., 0., 0.};
  double gop = 0.;
  const int nb = (sh.sh_degree + 1) * (sh.sh_degree + 1);
  if (shs != nullptr)
    for (int k = 0; k < sh.sh_M * 3; k++) dL_dshs[gi * (size_t)sh.sh_M * 3 + k] = 0.;

  // one round trip for the visibility of this Gaussian in every view of the sample, then only the visible ones
  unsigned long long vis_mask = 0ull;
#pragma unroll 8
  for (int lv = 0; lv < vps; lv++)
    vis_mask |= (unsigned long long)(__ldg(radii + (size_t)(s * vps + lv) * sh.P + i) > 0) << (lv & 63);

  for (int lv = 0; lv < vps; lv++) {
    const int v = s * vps + lv;
    const size_t o = (size_t)v * sh.P + i;
    double g2x = 0., g2y = 0.;
    if (vps <= 64 ? ((vis_mask >> lv) & 1ull) != 0ull : radii[o] > 0) {
      const Camera& cam = s_cams[lv];
      const float* vm = cam.view;
      const float* pm = cam.proj;
      const double* gr = ggrad + o * OCRF_GGRAD_STRIDE;  // dmx dmy dA dB dC dOp
      g2x = gr[0]; g2y = gr[1];
      gop += gr[5];
      const double fy = sh.H / (2.0 * cam.tanfovy), fx = sh.W / (2.0 * cam.tanfovx);
      // ---- conic -> cov2D -> cov3D and view-space mean ----
      double t[3];
#pragma unroll
      for (int r = 0; r < 3; r++) t[r] = vm[r] * x + vm[4 + r] * y + vm[8 + r] * z + vm[12 + r];
      const double limx = (double)1.3f * cam.tanfovx, limy = (double)1.3f * cam.tanfovy;
      const double txtz = t[0] / t[2], tytz = t[1] / t[2];
      t[0] = fmin(limx, fmax(-limx, txtz)) * t[2];
      t[1] = fmin(limy, fmax(-limy, tytz)) * t[2];
      const double xmul = (txtz < -limx || txtz > limx) ? 0. : 1.;
      const double ymul = (tytz < -limy || tytz > limy) ? 0. : 1.;
      const double itz = 1. / t[2], itz2 = itz * itz, itz3 = itz2 * itz;
      const double J00 = fx * itz, J11 = fy * itz, J02 = -(fx * t[0]) * itz2, J12 = -(fy * t[1]) * itz2;
      double T0[3], T1[3], ST0[3], ST1[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        T0[k] = vm[4 * k] * J00 + vm[4 * k + 2] * J02;
        T1[k] = vm[4 * k + 1] * J11 + vm[4 * k + 2] * J12;
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        ST0[k] = S[k][0] * T0[0] + S[k][1] * T0[1] + S[k][2] * T0[2];
        ST1[k] = S[k][0] * T1[0] + S[k][1] * T1[1] + S[k][2] * T1[2];
      }
      const double a = (T0[0] * ST0[0] + T0[1] * ST0[1] + T0[2] * ST0[2]) + (double)0.3f;
      const double b = T0[0] * ST1[0] + T0[1] * ST1[1] + T0[2] * ST1[2];
      const double c = (T1[0] * ST1[0] + T1[1] * ST1[1] + T1[2] * ST1[2]) + (double)0.3f;
      const double gA = gr[2], gB = gr[3], gC = gr[4];
      const double denom = a * c - b * b;
      const double d2i = 1.0 / ((denom * denom) + (double)0.0000001f);
      double da = 0., db = 0., dc = 0.;
      if (d2i != 0.) {
        da = d2i * (-c * c * gA + 2 * b * c * gB + (denom - a * c) * gC);
        dc = d2i * (-a * a * gC + 2 * a * b * gB + (denom - a * c) * gA);
        db = d2i * 2 * (b * c * gA - (denom + 2 * b * b) * gB + a * b * gC);
        gcov[0] += T0[0] * T0[0] * da + T0[0] * T1[0] * db + T1[0] * T1[0] * dc;
        gcov[3] += T0[1] * T0[1] * da + T0[1] * T1[1] * db + T1[1] * T1[1] * dc;
        gcov[5] += T0[2] * T0[2] * da + T0[2] * T1[2] * db + T1[2] * T1[2] * dc;
        gcov[1] += 2 * T0[0] * T0[1] * da + (T0[0] * T1[1] + T0[1] * T1[0]) * db + 2 * T1[0] * T1[1] * dc;
        gcov[2] += 2 * T0[0] * T0[2] * da + (T0[0] * T1[2] + T0[2] * T1[0]) * db + 2 * T1[0] * T1[2] * dc;
        gcov[4] += 2 * T0[2] * T0[1] * da + (T0[1] * T1[2] + T0[2] * T1[1]) * db + 2 * T1[1] * T1[2] * dc;
      }
      double dT0[3], dT1[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        dT0[k] = 2 * ST0[k] * da + ST1[k] * db;
        dT1[k] = 2 * ST1[k] * dc + ST0[k] * db;
      }
      const double dJ00 = vm[0] * dT0[0] + vm[4] * dT0[1] + vm[8] * dT0[2];
      const double dJ02 = vm[2] * dT0[0] + vm[6] * dT0[1] + vm[10] * dT0[2];
      const double dJ11 = vm[1] * dT1[0] + vm[5] * dT1[1] + vm[9] * dT1[2];
      const double dJ12 = vm[2] * dT1[0] + vm[6] * dT1[1] + vm[10] * dT1[2];
      const double dtx = xmul * -fx * itz2 * dJ02;
      const double dty = ymul * -fy * itz2 * dJ12;
      const double dtz = -fx * itz2 * dJ00 - fy * itz2 * dJ11 + (2 * fx * t[0]) * itz3 * dJ02 + (2 * fy * t[1]) * itz3 * dJ12;
#pragma unroll
      for (int j = 0; j < 3; j++) gmean[j] += vm[4 * j] * dtx + vm[4 * j + 1] * dty + vm[4 * j + 2] * dtz;
      // ---- screen-space mean -> 3D mean ----
      const double hw = pm[3] * x + pm[7] * y + pm[11] * z + pm[15];
      const double mw = 1.0 / (hw + (double)0.0000001f);
      const double mul1 = (pm[0] * x + pm[4] * y + pm[8] * z + pm[12]) * mw * mw;
      const double mul2 = (pm[1] * x + pm[5] * y + pm[9] * z + pm[13]) * mw * mw;
#pragma unroll
      for (int j = 0; j < 3; j++)
        gmean[j] += (pm[4 * j] * mw - pm[4 * j + 3] * mul1) * g2x + (pm[4 * j + 1] * mw - pm[4 * j + 3] * mul2) * g2y;
      // ---- SH colours (view dependent) ----
      if (shs != nullptr) {
        const double d0[3] = {x - cam.campos[0], y - cam.campos[1], z - cam.campos[2]};
        const double s2 = d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2];
        const double len = sqrt(s2);
        const double dir[3] = {d0[0] / len, d0[1] / len, d0[2] / len};
        double bas[16], bg3[16][3];
        { float bf[16]; sh_basis(sh.sh_degree, (float)dir[0], (float)dir[1], (float)dir[2], bf);
#pragma unroll
          for (int k = 0; k < 16; k++) bas[k] = bf[k]; }
        { float gf[16][3]; sh_basis_grad(sh.sh_degree, (float)dir[0], (float)dir[1], (float)dir[2], gf);
#pragma unroll
          for (int k = 0; k < 16; k++) { bg3[k][0] = gf[k][0]; bg3[k][1] = gf[k][1]; bg3[k][2] = gf[k][2]; } }
        double grgb[3], ddir[3] = {0., 0., 0.};
#pragma unroll
        for (int ch = 0; ch < 3; ch++) grgb[ch] = clamped[o * 3 + ch] ? 0. : dL_dcolors_view[o * 3 + ch];
        const float* coef = shs + gi * (size_t)sh.sh_M * 3;
        float* gsh = dL_dshs + gi * (size_t)sh.sh_M * 3;
        for (int k = 0; k < nb; k++)
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            gsh[3 * k + ch] += bas[k] * grgb[ch];
            const double cg = coef[3 * k + ch] * grgb[ch];
            ddir[0] += bg3[k][0] * cg; ddir[1] += bg3[k][1] * cg; ddir[2] += bg3[k][2] * cg;
          }
        const double inv32 = 1.0 / sqrt(s2 * s2 * s2);
        const double dotv = d0[0] * ddir[0] + d0[1] * ddir[1] + d0[2] * ddir[2];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) gmean[ax] += (s2 * ddir[ax] - d0[ax] * dotv) * inv32;
      }
    }
    if (dL_dmeans2D) {
      dL_dmeans2D[o * 3] = g2x;
      dL_dmeans2D[o * 3 + 1] = g2y;
      dL_dmeans2D[o * 3 + 2] = 0.;
    }
  }

#pragma unroll
  for (int j = 0; j < 3; j++) dL_dmeans3D[gi * 3 + j] = gmean[j];
  dL_dopacities[gi] = gop;
  if (cov3D_precomp != nullptr) {
    if (dL_dcov3D)
#pragma unroll
      for (int e = 0; e < 6; e++) dL_dcov3D[gi * 6 + e] = gcov[e];
    return;
  }
  // ---- cov3D -> scale, quaternion (no normalisation Jacobian, backward.cu:340) ----
  const double r = q.x, qx = q.y, qy = q.z, qz = q.w;
  const double Rg[3][3] = {{1. - 2. * (qy * qy + qz * qz), 2. * (qx * qy + r * qz), 2. * (qx * qz - r * qy)},
                          {2. * (qx * qy - r * qz), 1. - 2. * (qx * qx + qz * qz), 2. * (qy * qz + r * qx)},
                          {2. * (qx * qz + r * qy), 2. * (qy * qz - r * qx), 1. - 2. * (qx * qx + qy * qy)}};
  const double sm[3] = {(double)scale_modifier * sc[0], (double)scale_modifier * sc[1], (double)scale_modifier * sc[2]};
  const double Gs[3][3] = {{gcov[0], 0.5 * gcov[1], 0.5 * gcov[2]},
                          {0.5 * gcov[1], gcov[3], 0.5 * gcov[4]},
                          {0.5 * gcov[2], 0.5 * gcov[4], gcov[5]}};
  double D[3][3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    double ds = 0.;
#pragma unroll
    for (int c2 = 0; c2 < 3; c2++) {
      const double dM = 2.0 * sm[p] * (Rg[p][0] * Gs[0][c2] + Rg[p][1] * Gs[1][c2] + Rg[p][2] * Gs[2][c2]);
      ds += Rg[p][c2] * dM;
      D[p][c2] = sm[p] * dM;
    }
    dL_dscales[gi * 3 + p] = ds;
  }
  float4 gq;
  gq.x = (float)(2 * qz * (D[0][1] - D[1][0]) + 2 * qy * (D[2][0] - D[0][2]) + 2 * qx * (D[1][2] - D[2][1]));
  gq.y = (float)(2 * qy * (D[1][0] + D[0][1]) + 2 * qz * (D[2][0] + D[0][2]) + 2 * r * (D[1][2] - D[2][1]) - 4 * qx * (D[2][2] + D[1][1]));
  gq.z = (float)(2 * qx * (D[1][0] + D[0][1]) + 2 * r * (D[2][0] - D[0][2]) + 2 * qz * (D[1][2] + D[2][1]) - 4 * qy * (D[2][2] + D[0][0]));
  gq.w = (float)(2 * r * (D[0][1] - D[1][0]) + 2 * qx * (D[2][0] + D[0][2]) + 2 * qy * (D[1][2] + D[2][1]) - 4 * qz * (D[1][1] + D[0][0]));
  *reinterpret_cast<float4*>(dL_drotations + gi * 4) = gq;
}

}  // namespace ocrf

using namespace ocrf;

extern "C" int ocrf_preprocess_backward(void* stream, const OcrfShape* sh, const float* means3D, const float* scales,
                                        const float* rotations, const float* cov3D_precomp, const float* shs,
                                        const float* cams, float scale_modifier, const int32_t* radii,
                                        const void* geom_ws, const double* ggrad, const float* dL_dcolors_view,
                                        float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dopacities, float* dL_dscales,
                                        float* dL_drotations, float* dL_dcov3D, float* dL_dshs) {
  if (!sh || !means3D || !cams || !radii || !geom_ws || !ggrad || !dL_dmeans3D || !dL_dopacities) return OCRF_EINVAL;
  if (cov3D_precomp == nullptr && (!scales || !rotations || !dL_dscales || !dL_drotations)) return OCRF_EINVAL;
  if (shs != nullptr && (!dL_dshs || !dL_dcolors_view)) return OCRF_EINVAL;
  if (sh->P <= 0) return 0;
  OcrfGeomLayout G;
  ocrf_geom_layout(sh, shs != nullptr, &G);
  const size_t smem = (size_t)sh->views_per_sample * sizeof(Camera);
  if (smem > 48 * 1024) return OCRF_ECAPACITY;
  const dim3 grid(ceil_div(sh->P, PB_THREADS), sh->S);
  OCRF_LAUNCH(preprocess_backward_kernel, dim3(grid), dim3(PB_THREADS), smem, static_cast<cudaStream_t>(stream), 
      *sh, means3D, scales, rotations, cov3D_precomp, shs, reinterpret_cast<const Camera*>(cams), scale_modifier,
      radii, at<uint8_t>(geom_ws, G.clamped), ggrad, dL_dcolors_view, dL_dmeans3D, dL_dmeans2D, dL_dopacities,
      dL_dscales, dL_drotations, dL_dcov3D, dL_dshs);
  OCRF_CHECK_LAST();
  return 0;
}
