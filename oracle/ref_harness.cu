// C-ABI harness around the UNMODIFIED vendored Inria rasterizer
// (/root/reference/mmdet3d/models/necks/MVSGaussian/lib/submodules/diff-gaussian-rasterization/
//  cuda_rasterizer/{forward,backward,rasterizer_impl}.cu), compiled where those sources lie.
//
// TEST / BENCH INFRASTRUCTURE ONLY.  It plays the role of the reference's torch binding
// (rasterize_points.cu:36-196) without torch headers: it owns three grow-only device
// byte buffers (the reference grows torch tensors through callbacks, rasterize_points.cu:27-33),
// forwards raw pointers to CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (cuda_rasterizer/rasterizer.h:21-85) and lets the tests read the internal state
// (keys, point list, tile ranges, ...) back for bit-exact comparison.
//
// Nothing under ocrfdet_b200/ links or loads this file's output.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <cuda_runtime.h>

#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {

struct Buf {
  char* p = nullptr;
  size_t cap = 0;
  size_t used = 0;
  char* get(size_t n) {
    if (n > cap) {
      if (p) cudaFree(p);
      size_t want = n + n / 4 + 4096;
      if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cap = 0; return nullptr; }
      cap = want;
    }
    used = n;
    return p;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = used = 0; }
};

struct RefCtx {
  Buf geom, binning, img;
  int P = 0, W = 0, H = 0, R = 0;
};

}  // namespace

extern "C" {

void* ref_ctx_create() { return new RefCtx(); }

void ref_ctx_destroy(void* h) {
  RefCtx* c = static_cast<RefCtx*>(h);
  c->geom.release(); c->binning.release(); c->img.release();
  delete c;
}

int ref_num_channels() { return NUM_CHANNELS; }

// Mirrors RasterizeGaussiansCUDA (rasterize_points.cu:36-115).  All pointers are device
// pointers (or NULL for an absent optional input, as the binding passes for empty tensors).
// Returns num_rendered (>=0) or -1 on a CUDA error.
int ref_forward(void* h, int P, int D, int M, const float* background, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                int prefiltered, float* out_color, int* radii) {
  RefCtx* c = static_cast<RefCtx*>(h);
  c->P = P; c->W = W; c->H = H; c->R = 0;
  if (P == 0) return 0;
  std::function<char*(size_t)> g = [c](size_t n) { return c->geom.get(n); };
  std::function<char*(size_t)> b = [c](size_t n) { return c->binning.get(n); };
  std::function<char*(size_t)> i = [c](size_t n) { return c->img.get(n); };
  int rendered = CudaRasterizer::Rasterizer::forward(
      g, b, i, P, D, M, background, W, H, means3D, shs, colors_precomp, opacities, scales,
      scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx,
      tan_fovy, prefiltered != 0, out_color, radii, false);
  c->R = rendered;
  if (cudaGetLastError() != cudaSuccess) return -1;
  return rendered;
}

// Mirrors RasterizeGaussiansBackwardCUDA (rasterize_points.cu:118-196).  The caller zeroes
// the gradient outputs first, as the binding does with torch::zeros (:151-159).
int ref_backward(void* h, int P, int D, int M, int R, const float* background, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                 const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                 float* dL_dscale, float* dL_drot) {
  RefCtx* c = static_cast<RefCtx*>(h);
  if (P == 0) return 0;
  CudaRasterizer::Rasterizer::backward(
      P, D, M, R, background, W, H, means3D, shs, colors_precomp, scales, scale_modifier,
      rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii,
      c->geom.p, c->binning.p, c->img.p, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
      dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, false);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return 0;
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
  if (P == 0) return 0;
  CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Copies the reference's internal state of the LAST forward into caller-provided device
// arrays (any may be NULL).  Layout follows GeometryState/BinningState/ImageState::fromChunk
// (rasterizer_impl.cu:155-194).
int ref_get_state(void* h, float* depths, float* means2D, float* cov3D, float* conic_opacity,
                  uint32_t* tiles_touched, uint32_t* point_offsets, uint64_t* keys_unsorted,
                  uint32_t* values_unsorted, uint64_t* keys_sorted, uint32_t* point_list,
                  uint32_t* ranges, float* final_T, uint32_t* n_contrib) {
  RefCtx* c = static_cast<RefCtx*>(h);
  const size_t P = c->P, R = c->R, N = (size_t)c->W * c->H;
  if (P == 0) return 0;
  char* gp = c->geom.p;
  CudaRasterizer::GeometryState gs = CudaRasterizer::GeometryState::fromChunk(gp, P);
  char* ip = c->img.p;
  CudaRasterizer::ImageState is = CudaRasterizer::ImageState::fromChunk(ip, N);
  const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
  if (depths) cudaMemcpy(depths, gs.depths, P * 4, k);
  if (means2D) cudaMemcpy(means2D, gs.means2D, P * 8, k);
  if (cov3D) cudaMemcpy(cov3D, gs.cov3D, P * 24, k);
  if (conic_opacity) cudaMemcpy(conic_opacity, gs.conic_opacity, P * 16, k);
  if (tiles_touched) cudaMemcpy(tiles_touched, gs.tiles_touched, P * 4, k);
  if (point_offsets) cudaMemcpy(point_offsets, gs.point_offsets, P * 4, k);
  if (R > 0) {
    char* bp = c->binning.p;
    CudaRasterizer::BinningState bs = CudaRasterizer::BinningState::fromChunk(bp, R);
    if (keys_unsorted) cudaMemcpy(keys_unsorted, bs.point_list_keys_unsorted, R * 8, k);
    if (values_unsorted) cudaMemcpy(values_unsorted, bs.point_list_unsorted, R * 4, k);
    if (keys_sorted) cudaMemcpy(keys_sorted, bs.point_list_keys, R * 8, k);
    if (point_list) cudaMemcpy(point_list, bs.point_list, R * 4, k);
  }
  const size_t tiles = (size_t)((c->W + BLOCK_X - 1) / BLOCK_X) * ((c->H + BLOCK_Y - 1) / BLOCK_Y);
  if (ranges) cudaMemcpy(ranges, is.ranges, tiles * 8, k);
  if (final_T) cudaMemcpy(final_T, is.accum_alpha, N * 4, k);
  if (n_contrib) cudaMemcpy(n_contrib, is.n_contrib, N * 4, k);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
