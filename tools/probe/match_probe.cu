// What does __match_any_sync cost on sm_100a?  One CTA on one SM, 1 or 32 warps, 512 independent matches per warp on
// values with a controlled number of distinct groups per warp; the 8-ballot emulation of an 8-bit match beside it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o match_probe match_probe.cu && ./match_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int MODE>  // 0: match.any, 1: eight ballots
__global__ void probe(uint32_t mask, int iters, uint32_t* out, long long* cycles) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < iters; it++) {
    const uint32_t d = hash32(threadIdx.x * 977u + it * 131u + acc * 0u) & mask;
    uint32_t peers;
    if (MODE == 0) {
      peers = __match_any_sync(0xffffffffu, d);
    } else {
      peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const uint32_t vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? vote : ~vote;
      }
    }
    acc += __popc(peers & ((1u << lane) - 1));
  }
  __syncthreads();
  const long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 4096);
  cudaMallocManaged(&cyc, 8);
  const int iters = 512;
  const uint32_t masks[] = {0u, 1u, 3u, 7u, 15u, 31u, 255u, 0x1fffu};
  for (int threads : {32, 256, 1024}) {
    for (uint32_t m : masks) {
      long long c[2];
      for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
          if (mode == 0) probe<0><<<1, threads>>>(m, iters, out, cyc); else probe<1><<<1, threads>>>(m, iters, out, cyc);
          cudaDeviceSynchronize();
        }
        c[mode] = *cyc;
      }
      const double per = 1.0 / (double)iters / (threads / 32);
      printf("threads %4d  mask 0x%04x  match.any %7.1f cyc per warp-op per SM   8 ballots %7.1f   (total %lld / %lld)\n", threads, m,
             c[0] * per, c[1] * per, c[0], c[1]);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
