// Scope row f-3 ("next"): BEV pooling v2 -- the op that produces lss_feat / ht_feat immediately upstream of the
// Gaussian render path.
//
// Replaces bev_pool_v2_kernel / bev_pool_grad_kernel
// (/root/reference/mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-121) and the per-step regrouping of the point
// list that the reference does in Python before its backward (bev_pool_v2/bev_pool.py:47-60: argsort by ranks_feat,
// fancy-index the three rank arrays, boundary mask, where, diff).
//
//   forward : out[ranks_bev[s], c] = sum_{i in interval} depth[ranks_depth[i]] * feat[ranks_feat[i], c]
//   backward: depth_grad[ranks_depth[i]] = sum_c out_grad[ranks_bev[i], c] * feat[ranks_feat[i], c]
//             feat_grad[f, c]            = sum_{i: ranks_feat[i] = f} out_grad[ranks_bev[i], c] * depth[ranks_depth[i]]
//
// Forward: one thread per (interval, 4 channels) with 16-byte loads (the reference: one thread per channel, scalar
// loads); the per-(interval, channel) sum runs in the same order with the same fma contraction, so it is bit-exact
// against the reference kernel.  Eight points' rows are in flight per thread and the next eight points' indices
// travel with them; the FMAs stay in point order.
// Backward: the regrouping by feature pixel is a 2-3 pass onesweep sort of (ranks_feat, point) pairs (radix_sort.cu)
// instead of a torch argsort + 5 indexing kernels; one WARP then owns a feature pixel (binary search in the sorted keys,
// no interval arrays), fetches the indices / depth values of 32 points at a time in parallel, and walks them with the
// channels spread over the lanes: the feature row is read once, every out_grad row once (coalesced), the dot product
// for depth_grad is a warp reduction, feat_grad is written exactly once per pixel -- pixels without points get their
// zeros here, so only depth_grad needs a zero fill.  The reference walks one THREAD per pixel through two nested serial
// loops over points and channels with strided scalar loads.
#include "common.cuh"

namespace ocrf {

template <bool VEC4>
__global__ void __launch_bounds__(256) bev_pool_forward_kernel(int c, int cq, int n_intervals,
                                                               const float* __restrict__ depth,
                                                               const float* __restrict__ feat,
                                                               const int* __restrict__ ranks_depth,
                                                               const int* __restrict__ ranks_feat,
                                                               const int* __restrict__ ranks_bev,
                                                               const int* __restrict__ interval_starts,
                                                               const int* __restrict__ interval_lengths,
                                                               float* __restrict__ out) {
  // cq = work items per interval: c/4 (VEC4) or c
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int index = (int)(idx / cq);
  const int q = (int)(idx - (long long)index * cq);
  if (index >= n_intervals) return;
  const int start = __ldg(interval_starts + index);
  const int len = __ldg(interval_lengths + index);
  if (VEC4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int U = 8;  // rows in flight per thread; the FMAs below stay in point order (bit-exact with the reference)
    int i = 0;
    int rd[U], rf[U];
    if (len >= U) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        rd[u] = __ldg(ranks_depth + start + u);
        rf[u] = __ldg(ranks_feat + start + u);
      }
    }
    for (; i + U <= len; i += U) {
      float d[U];
      float4 f[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        d[u] = __ldg(depth + rd[u]);
        f[u] = __ldg(reinterpret_cast<const float4*>(feat + (size_t)rf[u] * c) + q);
      }
      if (i + 2 * U <= len) {  // the next batch's indices travel together with this batch's rows
#pragma unroll
        for (int u = 0; u < U; u++) {
          rd[u] = __ldg(ranks_depth + start + i + U + u);
          rf[u] = __ldg(ranks_feat + start + i + U + u);
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        acc.x = fmaf(f[u].x, d[u], acc.x);
        acc.y = fmaf(f[u].y, d[u], acc.y);
        acc.z = fmaf(f[u].z, d[u], acc.z);
        acc.w = fmaf(f[u].w, d[u], acc.w);
      }
    }
    for (; i < len; i++) {
      const float d = __ldg(depth + __ldg(ranks_depth + start + i));
      const float4 f = __ldg(reinterpret_cast<const float4*>(feat + (size_t)__ldg(ranks_feat + start + i) * c) + q);
      acc.x = fmaf(f.x, d, acc.x);
      acc.y = fmaf(f.y, d, acc.y);
      acc.z = fmaf(f.z, d, acc.z);
      acc.w = fmaf(f.w, d, acc.w);
    }
    reinterpret_cast<float4*>(out + (size_t)__ldg(ranks_bev + start) * c)[q] = acc;
  } else {
    float acc = 0.f;
    for (int i = 0; i < len; i++)
      acc = fmaf(__ldg(feat + (size_t)__ldg(ranks_feat + start + i) * c + q), __ldg(depth + __ldg(ranks_depth + start + i)), acc);
    out[(size_t)__ldg(ranks_bev + start) * c + q] = acc;
  }
}

__global__ void __launch_bounds__(256) bev_pool_keys_kernel(uint32_t n, const int* __restrict__ ranks_feat,
                                                            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                            uint32_t* __restrict__ n_dev) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_dev = n;
  if (i < n) {
    keys[i] = (uint64_t)(uint32_t)__ldg(ranks_feat + i);
    vals[i] = i;
  }
}

// first index in [a, b) whose key is >= target (b if none); keys ascending.  All 32 lanes call it with the same
// arguments; every round probes 32 positions at once, so 3 M keys take 5 rounds of one load each instead of 22.
__device__ __forceinline__ uint32_t warp_lower_bound(const uint64_t* __restrict__ keys, uint32_t a, uint32_t b,
                                                     uint64_t target) {
  const uint32_t lane = threadIdx.x & 31;
  while (a < b) {
    const uint32_t step = (b - a + 31) / 32;
    const uint32_t pos = a + lane * step;
    const bool below = pos < b && __ldg(keys + pos) < target;
    const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, below));  // the probes below the target form a prefix
    if (cnt == 0) return a;
    const uint32_t last_below = a + (cnt - 1) * step;
    const uint32_t first_not = a + cnt * step;  // may lie beyond b: then no probe was >= target
    a = last_below + 1;
    b = first_not < b ? first_not : b;
  }
  return a;
}

constexpr int BP_MAX_CPL = 4;  // channels per lane: C <= 128

template <int CPL>  // channels per lane = ceil(c / 32)
__global__ void __launch_bounds__(256) bev_pool_backward_kernel(int c, uint32_t n, int n_feat,
                                                                const float* __restrict__ out_grad,
                                                                const float* __restrict__ depth,
                                                                const float* __restrict__ feat,
                                                                const int* __restrict__ ranks_depth,
                                                                const int* __restrict__ ranks_bev,
                                                                const uint64_t* __restrict__ keys,  // sorted ranks_feat
                                                                const uint32_t* __restrict__ vals,  // point of each key
                                                                float* __restrict__ depth_grad,
                                                                float* __restrict__ feat_grad) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= n_feat) return;
  // [lo, hi) = the points of feature pixel f: lower_bound(f), lower_bound(f + 1), each a 32-ary search by the warp
  const uint32_t lo = warp_lower_bound(keys, 0u, n, (uint64_t)f);
  const uint32_t hi = warp_lower_bound(keys, lo, n, (uint64_t)f + 1);
  float frow[CPL], acc[CPL];
#pragma unroll
  for (int k = 0; k < CPL; k++) {
    const int ch = lane + 32 * k;
    frow[k] = (ch < c && lo < hi) ? __ldg(feat + (size_t)f * c + ch) : 0.f;
    acc[k] = 0.f;
  }
  // 32 points' metadata at a time, fetched in parallel one batch ahead of the walk
  int nbev = 0, nrd = 0;
  float nd = 0.f;
  auto fetch = [&](uint32_t base) {
    const uint32_t j = base + lane;
    nbev = 0; nrd = 0; nd = 0.f;
    if (j < hi) {
      const uint32_t p = __ldg(vals + j);
      nbev = __ldg(ranks_bev + p);
      nrd = __ldg(ranks_depth + p);
      nd = __ldg(depth + nrd);
    }
  };
  if (lo < hi) fetch(lo);
  for (uint32_t base = lo; base < hi; base += 32) {
    const uint32_t j = base + lane;
    const int bev = nbev, rd = nrd;
    const float d = nd;
    if (base + 32 < hi) fetch(base + 32);
    const int cnt = (int)min(32u, hi - base);
    float my_dot = 0.f;  // lane t ends up holding the dot product of point base + t
    constexpr int U = 4;  // out_grad rows in flight per warp
    for (int t0 = 0; t0 < cnt; t0 += U) {
      float gv[U][CPL], dt[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int t = min(t0 + u, cnt - 1);  // the tail repeats the last point (warp-uniform control flow, results unused)
        const int bt = __shfl_sync(0xffffffffu, bev, t);
        dt[u] = __shfl_sync(0xffffffffu, d, t);
        const float* g = out_grad + (size_t)bt * c;
#pragma unroll
        for (int k = 0; k < CPL; k++) {
          const int ch = lane + 32 * k;
          gv[u][k] = ch < c ? __ldg(g + ch) : 0.f;
        }
      }
      float part[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        part[u] = 0.f;
        if (t0 + u < cnt) {  // warp-uniform
#pragma unroll
          for (int k = 0; k < CPL; k++) {
            part[u] = fmaf(gv[u][k], frow[k], part[u]);
            acc[k] = fmaf(gv[u][k], dt[u], acc[k]);  // point order: bit-exact with the reference's serial sum
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < U; u++) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        if (lane == t0 + u) my_dot = part[u];
    }
    if (j < hi) depth_grad[rd] = my_dot;  // the reference stores too (bev_pool_cuda.cu:103): colliding ranks_depth race there as here
  }
#pragma unroll
  for (int k = 0; k < CPL; k++) {
    const int ch = lane + 32 * k;
    if (ch < c) feat_grad[(size_t)f * c + ch] = acc[k];
  }
}

}  // namespace ocrf

using namespace ocrf;

static int bp_key_bits(int n_feat) {
  int b = 1;
  while (b < 32 && ((uint64_t)n_feat >> b)) b++;
  return b;
}

extern "C" int ocrf_bev_pool_forward(void* stream, int32_t c, int32_t n_intervals, const float* depth, const float* feat,
                                     const int32_t* ranks_depth, const int32_t* ranks_feat, const int32_t* ranks_bev,
                                     const int32_t* interval_starts, const int32_t* interval_lengths, float* out) {
  if (c <= 0 || n_intervals < 0) return OCRF_EINVAL;
  if (n_intervals == 0) return 0;
  if (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !interval_starts || !interval_lengths || !out)
    return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int cq = vec ? c / 4 : c;
  const long long items = (long long)n_intervals * cq;
  const unsigned grid = (unsigned)((items + 255) / 256);
  if (vec)
    bev_pool_forward_kernel<true><<<grid, 256, 0, st>>>(c, cq, n_intervals, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                                        interval_starts, interval_lengths, out);
  else
    bev_pool_forward_kernel<false><<<grid, 256, 0, st>>>(c, cq, n_intervals, depth, feat, ranks_depth, ranks_feat,
                                                         ranks_bev, interval_starts, interval_lengths, out);
  OCRF_CHECK_LAST();
  return 0;
}

// workspace: keys/vals ping-pong of the (ranks_feat, point) sort + its onesweep state + the device-side count
static size_t bp_ws_layout(uint64_t n, size_t* keys_a, size_t* keys_b, size_t* vals_a, size_t* vals_b, size_t* sort_ws,
                           size_t* n_dev) {
  size_t off = 0;
  const uint64_t m = n ? n : 1;
  *keys_a = off; off = align128(off + m * 8);
  *keys_b = off; off = align128(off + m * 8);
  *vals_a = off; off = align128(off + m * 4);
  *vals_b = off; off = align128(off + m * 4);
  *sort_ws = off; off = align128(off + sort_ws_layout(m).total + 128);
  *n_dev = off; off = align128(off + 16);
  return off;
}

extern "C" size_t ocrf_bev_pool_backward_workspace_bytes(uint64_t n_points) {
  size_t a, b, c, d, e, f;
  return bp_ws_layout(n_points, &a, &b, &c, &d, &e, &f) + 128;
}

extern "C" int ocrf_bev_pool_backward(void* stream, int32_t c, uint64_t n_points, int32_t n_feat, const float* out_grad,
                                      const float* depth, const float* feat, const int32_t* ranks_depth,
                                      const int32_t* ranks_feat, const int32_t* ranks_bev, float* depth_grad,
                                      float* feat_grad, void* ws) {
  if (c <= 0 || c > 32 * BP_MAX_CPL || n_feat < 0) return OCRF_EINVAL;
  if (n_points >= (1ull << 30)) return OCRF_ECAPACITY;
  if (n_feat == 0) return 0;
  if (!out_grad || !depth || !feat || !depth_grad || !feat_grad || !ws) return OCRF_EINVAL;
  if (n_points > 0 && (!ranks_depth || !ranks_feat || !ranks_bev)) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t oka, okb, ova, ovb, osw, ond;
  bp_ws_layout(n_points, &oka, &okb, &ova, &ovb, &osw, &ond);
  const uint64_t* keys_sorted = at<uint64_t>(ws, oka);
  const uint32_t* vals_sorted = at<uint32_t>(ws, ova);
  if (n_points > 0) {
    const int bits = bp_key_bits(n_feat);
    const int passes = (bits + 7) / 8;
    // the sort ends in the "b" pair when the pass count is odd, in "a" when it is even: start so that it ends in a
    uint64_t* ka = at<uint64_t>(ws, (passes & 1) ? okb : oka);
    uint32_t* va = at<uint32_t>(ws, (passes & 1) ? ovb : ova);
    uint64_t* kb = at<uint64_t>(ws, (passes & 1) ? oka : okb);
    uint32_t* vb = at<uint32_t>(ws, (passes & 1) ? ova : ovb);
    uint32_t* n_dev = at<uint32_t>(ws, ond);
    bev_pool_keys_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, st>>>((uint32_t)n_points, ranks_feat, ka, va, n_dev);
    const int rc = sort_pairs_device(st, n_dev, n_points, 0, bits, ka, va, kb, vb, at<char>(ws, osw));
    if (rc) return rc;
  }
  const int warps = 8;
  const unsigned grid = (unsigned)((n_feat + warps - 1) / warps);
#define OCRF_BP_LAUNCH(CPL)                                                                                            \
  bev_pool_backward_kernel<CPL><<<grid, warps * 32, 0, st>>>(c, (uint32_t)n_points, n_feat, out_grad, depth, feat,        \
                                                             ranks_depth, ranks_bev, keys_sorted, vals_sorted, depth_grad, \
                                                             feat_grad)
  switch ((c + 31) / 32) {
    case 1: OCRF_BP_LAUNCH(1); break;
    case 2: OCRF_BP_LAUNCH(2); break;
    case 3: OCRF_BP_LAUNCH(3); break;
    default: OCRF_BP_LAUNCH(4); break;
  }
#undef OCRF_BP_LAUNCH
  OCRF_CHECK_LAST();
  return 0;
}
