// C-ABI harness around the reference's OWN bev_pool_v2 CUDA kernels, compiled from the source where it lies
// (/root/reference/mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu) by `make -C oracle refbev`.
// TEST / BENCH INFRASTRUCTURE ONLY.  The reference launches on the legacy default stream.
#include <cuda_runtime.h>

void bev_pool_v2(int c, int n_intervals, const float* depth, const float* feat, const int* ranks_depth,
                 const int* ranks_feat, const int* ranks_bev, const int* interval_starts, const int* interval_lengths,
                 float* out);
void bev_pool_v2_grad(int c, int n_intervals, const float* out_grad, const float* depth, const float* feat,
                      const int* ranks_depth, const int* ranks_feat, const int* ranks_bev, const int* interval_starts,
                      const int* interval_lengths, float* depth_grad, float* feat_grad);

extern "C" int ref_bev_pool_forward(int c, int n_intervals, const float* depth, const float* feat, const int* ranks_depth,
                                    const int* ranks_feat, const int* ranks_bev, const int* interval_starts,
                                    const int* interval_lengths, float* out) {
  bev_pool_v2(c, n_intervals, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths, out);
  return (int)cudaGetLastError();
}

extern "C" int ref_bev_pool_backward(int c, int n_intervals, const float* out_grad, const float* depth, const float* feat,
                                     const int* ranks_depth, const int* ranks_feat, const int* ranks_bev,
                                     const int* interval_starts, const int* interval_lengths, float* depth_grad,
                                     float* feat_grad) {
  bev_pool_v2_grad(c, n_intervals, out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
                   interval_lengths, depth_grad, feat_grad);
  return (int)cudaGetLastError();
}
