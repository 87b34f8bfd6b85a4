// Stage 1: voxel-Gaussian preprocess for a batch of views, fused with the tile-count scan.
//
// Replaces preprocessCUDA (cuda_rasterizer/forward.cu:155-256), computeCov3D (:118-152),
// computeCov2D (:74-113), in_frustum / getRect / ndc2Pix (auxiliary.h:41-56,139-164) and the
// cub::DeviceScan::InclusiveSum that follows it (rasterizer_impl.cu:277), plus checkFrustum
// (rasterizer_impl.cu:54-66) for markVisible.
//
// Design: one CTA of 256 threads per 256 consecutive Gaussians of one view.  Positions and scales
// ([P,3] AoS) are fetched as 192 float4 per array into shared memory, quaternions as one float4 per
// thread, so every HBM request is a full 16-byte vector.  Each thread then projects one Gaussian.
// The per-CTA sum of tiles_touched is chained to the previous CTAs with a decoupled look-back
// (CTAs take a ticket so the chain order is launch order), which turns the reference's separate
// scan kernel + temp storage into ~100 extra instructions here.
//
// Arithmetic: everything that feeds the sort keys (depth bits, pixel centre, radius, tile
// rectangle) is written with explicit round-to-nearest intrinsics in the exact contraction pattern
// nvcc produces for the reference's expressions (a*b + c*d + e*f == fma(e,f, fma(a,b, c*d)); IEEE
// division / reciprocal / sqrt; ndc2Pix in double with one fma), so keys are bit-identical to the
// reference build.  oracle/ocrf_oracle.c states the same pattern in C.
#include "common.cuh"
#include "gaussian_math.cuh"

namespace ocrf {

#ifndef OCRF_PRE_MINB
#define OCRF_PRE_MINB 5  // 48 registers: five CTAs per SM (the kernel is latency bound: ticket, staging, look-back)
#endif

__global__ void __launch_bounds__(PRE_THREADS, OCRF_PRE_MINB) preprocess_forward_kernel(
    OcrfShape sh, int blocks_per_view, const float* __restrict__ means3D, const float* __restrict__ scales,
    const float* __restrict__ rotations, const float* __restrict__ cov3D_precomp, const float* __restrict__ opacities,
    const float* __restrict__ shs, const Camera* __restrict__ cams, float scale_modifier, int prefiltered,
    float min_opacity, int32_t* __restrict__ radii, uint32_t* __restrict__ header, float* __restrict__ depths, float2* __restrict__ xy,
    float4* __restrict__ conic_opacity, uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ offsets,
    float* __restrict__ rgb, uint8_t* __restrict__ clamped, unsigned long long* __restrict__ scan_status,
    uint64_t* __restrict__ vis_keys, uint32_t* __restrict__ vis_vals, uint32_t* __restrict__ view_start) {
  pdl_enter();
  __shared__ __align__(16) float s_pos[PRE_THREADS * 3];
  __shared__ __align__(16) float s_scl[PRE_THREADS * 3];
  __shared__ Camera s_cam;
  __shared__ unsigned long long s_warp_tot[PRE_THREADS / 32];
  __shared__ uint32_t s_ticket;
  __shared__ unsigned long long s_prefix;

  const int tid = threadIdx.x;
  if (tid == 0) s_ticket = atomicAdd(&header[HDR_TICKET], 1u);
  __syncthreads();
  const int bid = (int)s_ticket;  // logical CTA index, in arrival order
  const int v = bid / blocks_per_view;
  const int i0 = (bid - v * blocks_per_view) * PRE_THREADS;
  const int s = v / sh.views_per_sample;
  const int n = min(PRE_THREADS, sh.P - i0);
  const size_t gbase = (size_t)s * sh.P + i0;  // first Gaussian of this CTA inside the sample arrays

  if (tid < OCRF_CAM_STRIDE) reinterpret_cast<float*>(&s_cam)[tid] = reinterpret_cast<const float*>(cams + v)[tid];
  // vectorised staging of the [n,3] position / scale slabs
  {
    const float* src = means3D + gbase * 3;
    const int nf = n * 3;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const int nv = nf >> 2;
      if (tid < nv) reinterpret_cast<float4*>(s_pos)[tid] = __ldg(reinterpret_cast<const float4*>(src) + tid);
      if (tid < (nf & 3)) s_pos[(nv << 2) + tid] = __ldg(src + (nv << 2) + tid);
    } else {
      for (int k = tid; k < nf; k += PRE_THREADS) s_pos[k] = __ldg(src + k);
    }
    if (scales != nullptr) {
      const float* ssrc = scales + gbase * 3;
      if ((reinterpret_cast<uintptr_t>(ssrc) & 15) == 0) {
        const int nv = nf >> 2;
        if (tid < nv) reinterpret_cast<float4*>(s_scl)[tid] = __ldg(reinterpret_cast<const float4*>(ssrc) + tid);
        if (tid < (nf & 3)) s_scl[(nv << 2) + tid] = __ldg(ssrc + (nv << 2) + tid);
      } else {
        for (int k = tid; k < nf; k += PRE_THREADS) s_scl[k] = __ldg(ssrc + k);
      }
    }
  }
  __syncthreads();

  uint32_t my_tiles = 0;
  float my_depth = 0.f;
  const size_t o = (size_t)v * sh.P + i0 + tid;  // index into the per-(view, Gaussian) arrays
  if (tid < n) {
    int rad_out = 0;
    const float x = s_pos[3 * tid], y = s_pos[3 * tid + 1], z = s_pos[3 * tid + 2];
    const float* vm = s_cam.view;
    const float* pm = s_cam.proj;
    const float tz = xform_row(vm, 2, x, y, z);
    // Foreground filter (optional, min_opacity > 0): alpha = min(0.99, opacity * G) with G <= 1, so a Gaussian whose
    // opacity is below the blend threshold 1/255 can never pass `alpha < 1/255 -> continue` (forward.cu:345,
    // backward.cu:478): dropping it here changes no pixel and no gradient, only radii / the pair count.
    const bool background = min_opacity > 0.f && __ldg(opacities + gbase + tid) < min_opacity;
    if (tz > 0.2f && !background) {
      const float tx = xform_row(vm, 0, x, y, z), ty = xform_row(vm, 1, x, y, z);
      const float hx = xform_row(pm, 0, x, y, z), hy = xform_row(pm, 1, x, y, z), hw = xform_row(pm, 3, x, y, z);
      const float pw = __frcp_rn(__fadd_rn(hw, 0.0000001f));
      const float ndcx = __fmul_rn(hx, pw), ndcy = __fmul_rn(hy, pw);
      float c6[6];
      if (cov3D_precomp != nullptr) {
        const float* cp = cov3D_precomp + (gbase + tid) * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = __ldg(cp + k);
      } else {
        const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + gbase + tid);
        cov3d_from_scale_rot(s_scl[3 * tid], s_scl[3 * tid + 1], s_scl[3 * tid + 2], scale_modifier, q, c6);
      }
      const float fy = __fdiv_rn((float)sh.H, __fmul_rn(2.0f, s_cam.tanfovy));
      const float fx = __fdiv_rn((float)sh.W, __fmul_rn(2.0f, s_cam.tanfovx));
      float a, b, c;
      cov2d_ewa(tx, ty, tz, fx, fy, s_cam.tanfovx, s_cam.tanfovy, c6, vm, a, b, c);
      const float det = __fmaf_rn(a, c, -__fmul_rn(b, b));
      if (det != 0.0f) {
        const float det_inv = __frcp_rn(det);
        const float mid = __fmul_rn(0.5f, __fadd_rn(a, c));
        const float disc = __fsqrt_rn(fmaxf(0.1f, __fmaf_rn(mid, mid, -det)));
        const float lam = fmaxf(__fadd_rn(mid, disc), __fsub_rn(mid, disc));
        const int rad = (int)ceilf(__fmul_rn(3.f, __fsqrt_rn(lam)));
        const float px = ndc_to_pix(ndcx, sh.W), py = ndc_to_pix(ndcy, sh.H);
        const int gx = ceil_div(sh.W, TILE), gy = ceil_div(sh.H, TILE);
        int x0, y0, x1, y1;
        tile_rect(px, py, rad, gx, gy, x0, y0, x1, y1);
        const uint32_t area = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
        if (area != 0) {
          if (shs != nullptr) {  // forward.cu:20-71
            float dx = x - s_cam.campos[0], dy = y - s_cam.campos[1], dz = z - s_cam.campos[2];
            const float len = sqrtf(dx * dx + dy * dy + dz * dz);
            dx /= len; dy /= len; dz /= len;
            float bas[16];
            sh_basis(sh.sh_degree, dx, dy, dz, bas);
            const int nb = (sh.sh_degree + 1) * (sh.sh_degree + 1);
            const float* coef = shs + (gbase + tid) * (size_t)sh.sh_M * 3;
            float acc[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < nb; k++) {
              acc[0] += bas[k] * __ldg(coef + 3 * k);
              acc[1] += bas[k] * __ldg(coef + 3 * k + 1);
              acc[2] += bas[k] * __ldg(coef + 3 * k + 2);
            }
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
              const float val = acc[ch] + 0.5f;
              clamped[o * 3 + ch] = val < 0.f;
              rgb[o * 3 + ch] = fmaxf(val, 0.f);
            }
          }
          depths[o] = tz;
          my_depth = tz;
          xy[o] = make_float2(px, py);
          conic_opacity[o] = make_float4(__fmul_rn(c, det_inv), __fmul_rn(b, -det_inv), __fmul_rn(a, det_inv),
                                         __ldg(opacities + gbase + tid));
          rad_out = rad;
          my_tiles = area;
        }
      }
    } else if (prefiltered && !background) {
      atomicOr(&header[HDR_ERROR], ERR_PREFILTERED);  // the reference traps here (auxiliary.h:156-160)
    }
    radii[o] = rad_out;
    tiles_touched[o] = my_tiles;
  }

  // ---- CTA-wide inclusive scan, chained across CTAs by decoupled look-back.  One 64-bit value carries
  //      two counters: low word = tiles touched (-> offsets), high word = visible Gaussians (-> the
  //      compact slot of every visible Gaussian in the (view | depth) sort input). ----
  const int lane = tid & 31, warp = tid >> 5;
  unsigned long long incl = ((unsigned long long)(my_tiles != 0) << 32) | my_tiles;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp_tot[warp] = incl;
  __syncthreads();
  unsigned long long warp_off = 0, cta_total = 0;
#pragma unroll
  for (int w = 0; w < PRE_THREADS / 32; w++) {
    const unsigned long long t = s_warp_tot[w];
    if (w < warp) warp_off += t;
    cta_total += t;
  }
  if (warp == 0) {
    const unsigned long long excl = lookback_warp(scan_status, bid, cta_total);
    if (lane == 0) {
      s_prefix = excl;
      if (i0 == 0) view_start[v] = (uint32_t)(excl >> 32);
      if (bid == (int)gridDim.x - 1) {
        header[HDR_NUM_PAIRS] = (uint32_t)(excl + cta_total);
        header[HDR_NUM_VIS] = (uint32_t)((excl + cta_total) >> 32);
        view_start[sh.V] = (uint32_t)((excl + cta_total) >> 32);
      }
    }
  }
  __syncthreads();
  if (tid < n) {
    const unsigned long long mine = s_prefix + warp_off + incl;
    offsets[o] = (uint32_t)mine;
    if (my_tiles != 0) {
      const uint32_t slot = (uint32_t)(mine >> 32) - 1;
      vis_keys[slot] = ((uint64_t)(uint32_t)v << 32) | __float_as_uint(my_depth);
      vis_vals[slot] = (uint32_t)o;
    }
  }
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
  present[i] = xform_row(view, 2, x, y, z) > 0.2f;
}

}  // namespace ocrf

using namespace ocrf;

extern "C" int ocrf_geom_layout(const OcrfShape* sh, int use_sh, OcrfGeomLayout* out) {
  if (!sh || !out || sh->V <= 0 || sh->P < 0) return OCRF_EINVAL;
  const size_t n = (size_t)sh->V * sh->P;
  const size_t blocks = (size_t)sh->V * ceil_div(sh->P > 0 ? sh->P : 1, PRE_THREADS);
  size_t off = 0;
  out->header = off;        off = align128(off + 32 * 4);
  out->scan_status = off;   off = align128(off + blocks * 8);
  out->vis_sort_ws = off;   off = align128(off + sort_ws_layout(n ? n : 1).total + 128);  // zeroed with the two above
  out->depths = off;        off = align128(off + n * 4);
  out->xy = off;            off = align128(off + n * 8);
  out->conic_opacity = off; off = align128(off + n * 16);
  out->tiles_touched = off; off = align128(off + n * 4);
  out->offsets = off;       off = align128(off + n * 4);
  out->rgb = off;           off = align128(off + (use_sh ? n * 12 : 0));
  out->clamped = off;       off = align128(off + (use_sh ? n * 3 : 0));
  out->vis_keys = off;      off = align128(off + n * 8);
  out->vis_keys_tmp = off;  off = align128(off + n * 8);
  out->vis_vals = off;      off = align128(off + n * 4);
  out->vis_vals_tmp = off;  off = align128(off + n * 4);
  out->view_start = off;    off = align128(off + ((size_t)sh->V + 1) * 4);
  out->total = off + 128;
  return 0;
}

extern "C" int ocrf_preprocess_forward(void* stream, const OcrfShape* sh, const float* means3D, const float* scales,
                                       const float* rotations, const float* cov3D_precomp, const float* opacities,
                                       const float* shs, const float* cams, float scale_modifier, int prefiltered,
                                       int32_t* radii, void* geom_ws) {
  return ocrf_preprocess_forward_filtered(stream, sh, means3D, scales, rotations, cov3D_precomp, opacities, shs, cams,
                                          scale_modifier, prefiltered, 0.f, radii, geom_ws);
}

extern "C" int ocrf_preprocess_forward_filtered(void* stream, const OcrfShape* sh, const float* means3D,
                                                const float* scales, const float* rotations, const float* cov3D_precomp,
                                                const float* opacities, const float* shs, const float* cams,
                                                float scale_modifier, int prefiltered, float min_opacity, int32_t* radii,
                                                void* geom_ws) {
  if (!sh || !means3D || !opacities || !cams || !radii || !geom_ws) return OCRF_EINVAL;
  if (sh->V <= 0 || sh->P <= 0 || sh->W <= 0 || sh->H <= 0 || sh->views_per_sample <= 0 ||
      sh->V % sh->views_per_sample != 0 || sh->S * sh->views_per_sample != sh->V)
    return OCRF_EINVAL;
  if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr)) return OCRF_EINVAL;
  if (shs != nullptr && (sh->sh_degree < 0 || sh->sh_degree > 3 || sh->sh_M < (sh->sh_degree + 1) * (sh->sh_degree + 1)))
    return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OcrfGeomLayout L;
  ocrf_geom_layout(sh, shs != nullptr, &L);
  const int bpv = ceil_div(sh->P, PRE_THREADS);
  const int blocks = bpv * sh->V;
  // header and look-back state are adjacent: one memset resets both
  cudaMemsetAsync(at<char>(geom_ws, L.header), 0, L.depths - L.header, st);
  OCRF_LAUNCH(preprocess_forward_kernel, dim3(blocks), dim3(PRE_THREADS), 0, st, 
      *sh, bpv, means3D, scales, rotations, cov3D_precomp, opacities, shs, reinterpret_cast<const Camera*>(cams),
      scale_modifier, prefiltered, min_opacity > 1.0f / 255.0f ? 1.0f / 255.0f : min_opacity, radii,
      at<uint32_t>(geom_ws, L.header), at<float>(geom_ws, L.depths),
      at<float2>(geom_ws, L.xy), at<float4>(geom_ws, L.conic_opacity), at<uint32_t>(geom_ws, L.tiles_touched),
      at<uint32_t>(geom_ws, L.offsets), at<float>(geom_ws, L.rgb), at<uint8_t>(geom_ws, L.clamped),
      at<unsigned long long>(geom_ws, L.scan_status),
      // the sort of the visible Gaussians ends in (vis_keys, vis_vals): start in the tmp half when its pass count is odd
      at<uint64_t>(geom_ws, vis_sort_starts_in_tmp(sh->V) ? L.vis_keys_tmp : L.vis_keys),
      at<uint32_t>(geom_ws, vis_sort_starts_in_tmp(sh->V) ? L.vis_vals_tmp : L.vis_vals),
      at<uint32_t>(geom_ws, L.view_start));
  OCRF_CHECK_LAST();
  return 0;
}

extern "C" int ocrf_mark_visible(void* stream, int32_t P, const float* means3D, const float* viewmatrix,
                                 const float* projmatrix, uint8_t* present) {
  (void)projmatrix;
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return OCRF_EINVAL;
  if (P == 0) return 0;
  mark_visible_kernel<<<ceil_div(P, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P, means3D, viewmatrix, present);
  OCRF_CHECK_LAST();
  return 0;
}
