// Hardware probe for the tcgen05 building blocks of the 80-channel blend kernels (render_tc.cu):
//   * tcgen05.mma.cta_group::1.kind::tf32 with A in tensor memory (written by the threads with tcgen05.st, lane = row)
//     and B in shared memory, no-swizzle ("interleave") canonical layout, through BOTH major-nesses of the same buffer
//       forward  view: D[128 x 80] = A[128 x 32 records] . F[32 records x 80 channels]   (B is MN-major: N = channel)
//       backward view: D[128 x 32] = G[128 x 80 channels] . F^T                         (B is K-major:  N = record)
//   * how the tensor core reads an fp32 bit pattern as tf32 (truncation or rounding), which decides how the hi/lo
//     split of the 3xTF32 product has to be formed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/probe/umma_probe.cu ; prints max errors.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int R = 32, C = 80;
constexpr uint32_t S_C = 128;         // bytes between core matrices adjacent along the channels
constexpr uint32_t S_R = 20 * 128;    // bytes between core matrices adjacent along the records

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, int accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ F,
                                                    const float* __restrict__ G, float* __restrict__ Df,
                                                    float* __restrict__ Db) {
  __shared__ __align__(128) unsigned char s_f[R * C * 4];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // F -> canonical core-matrix layout: core (r / 8, c / 4) = 8 records x 16 bytes
  for (int e = tid; e < R * (C / 4); e += 128) {
    const int r = e % R, cc = e / R;
    const float4 v = *reinterpret_cast<const float4*>(F + r * C + cc * 4);
    *reinterpret_cast<float4*>(s_f + (r >> 3) * S_R + cc * S_C + (r & 7) * 16) = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t D_COL = 0, A_COL = 128;

  // ---------------- forward view ----------------
  {
    uint32_t a[R];
#pragma unroll
    for (int k = 0; k < R; k++) a[k] = __float_as_uint(A[tid * R + k]);
#pragma unroll
    for (int k = 0; k < R; k += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + lane_base + A_COL + k),
                   "r"(a[k]), "r"(a[k + 1]), "r"(a[k + 2]), "r"(a[k + 3]), "r"(a[k + 4]), "r"(a[k + 5]), "r"(a[k + 6]),
                   "r"(a[k + 7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = make_idesc(128, C, 1);
      for (int ks = 0; ks < R / 8; ks++)
        mma_ts(tm + D_COL, tm + A_COL + ks * 8, make_desc(smem_u32(s_f) + ks * S_R, S_R, S_C), idesc, ks > 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW0:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D0;\nbra W0;\nD0:\n}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c0 = 0; c0 < C; c0 += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tm + lane_base + D_COL + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; i++) Df[tid * C + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---------------- backward view ----------------
  {
#pragma unroll
    for (int k = 0; k < C; k += 8) {
      uint32_t g[8];
#pragma unroll
      for (int i = 0; i < 8; i++) g[i] = __float_as_uint(G[tid * C + k + i]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + lane_base + A_COL + k),
                   "r"(g[0]), "r"(g[1]), "r"(g[2]), "r"(g[3]), "r"(g[4]), "r"(g[5]), "r"(g[6]), "r"(g[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = make_idesc(128, R, 0);
      for (int ks = 0; ks < C / 8; ks++)
        mma_ts(tm + D_COL, tm + A_COL + ks * 8, make_desc(smem_u32(s_f) + ks * 2 * S_C, S_C, S_R), idesc, ks > 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 1;\n@p bra D1;\nbra W1;\nD1:\n}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c0 = 0; c0 < R; c0 += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tm + lane_base + D_COL + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; i++) Db[tid * R + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A(128 * R), F(R * C), G(128 * C), Df(128 * C), Db(128 * R);
  srand(7);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& v : A) v = rnd();
  for (auto& v : F) v = rnd();
  for (auto& v : G) v = rnd();
  float *dA, *dF, *dG, *dDf, *dDb;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dF, F.size() * 4); cudaMalloc(&dG, G.size() * 4);
  cudaMalloc(&dDf, Df.size() * 4); cudaMalloc(&dDb, Db.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dDf, 0xff, Df.size() * 4); cudaMemset(dDb, 0xff, Db.size() * 4);
  probe_kernel<<<1, 128>>>(dA, dF, dG, dDf, dDb);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(Df.data(), dDf, Df.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(Db.data(), dDb, Db.size() * 4, cudaMemcpyDeviceToHost);
  for (int mode = 0; mode < 3; mode++) {  // 0 exact, 1 truncated inputs, 2 rounded inputs
    auto q = [&](float x) { return mode == 0 ? x : mode == 1 ? trunc_tf32(x) : round_tf32(x); };
    double ef = 0, eb = 0;
    for (int m = 0; m < 128; m++) {
      for (int c = 0; c < C; c++) {
        double s = 0;
        for (int r = 0; r < R; r++) s += (double)q(A[m * R + r]) * q(F[r * C + c]);
        ef = fmax(ef, fabs(s - Df[m * C + c]));
      }
      for (int r = 0; r < R; r++) {
        double s = 0;
        for (int c = 0; c < C; c++) s += (double)q(G[m * C + c]) * q(F[r * C + c]);
        eb = fmax(eb, fabs(s - Db[m * R + r]));
      }
    }
    printf("reference %s: forward-view max err %.3e   backward-view max err %.3e\n",
           mode == 0 ? "exact fp32 inputs " : mode == 1 ? "truncated to tf32 " : "rounded to tf32   ", ef, eb);
  }
  printf("sample Df[0][0..3] = %g %g %g %g\n", Df[0], Df[1], Df[2], Df[3]);
  return 0;
}
