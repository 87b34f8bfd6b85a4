// Device helpers shared by the blend backward kernels (render_bwd.cu, render_tc_bwd.cu).
#pragma once
#include "common.cuh"

namespace ocrf {

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 1 / x for x in [0.01, 1] (x = 1 - alpha): one MUFU.RCP, no range fix-up code
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Reduce-scatter of 8 values over the warp: afterwards lane L (any L) holds the warp total of value
// index ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1).
__device__ __forceinline__ float butterfly8(float (&v)[8], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  float w[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float keep = h16 ? v[i + 4] : v[i];
    const float send = h16 ? v[i] : v[i + 4];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float u[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float keep = h8 ? w[i + 2] : w[i];
    const float send = h8 ? w[i] : w[i + 2];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float keep = h4 ? u[1] : u[0];
  const float send = h4 ? u[0] : u[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// ---- tensor-core helpers for the feature gradients of the generic-channel backward -------------------------------
// dF[record][channel] = sum over the pixels of a tile of w[pixel][record] * g[pixel][channel] is a dense product.
// Every warp multiplies ITS 32 pixels: A = w^T [16 records x 32 pixels] (staged through shared memory as the records
// are walked), B = g [32 pixels x CP channels] (staged once per tile), C = a [16 x CP] partial that the eight warps
// then add up.  mma.sync.m16n8k8 TF32 with the 3-term split (hi*hi + lo*hi + hi*lo) keeps fp32 accuracy: plain TF32
// (10-bit mantissa) misses the 1e-4 gradient bar on sums of 256 mixed-sign terms.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// the same split for a consumer that TRUNCATES fp32 to tf32 (both tensor-core paths of sm_100 do, probed with
// tools/probe/umma_probe.cu): hi is the value itself, lo the exact remainder
__device__ __forceinline__ void split_trunc(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
}  // namespace ocrf
