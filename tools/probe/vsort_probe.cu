// Where does visible_sort_reg_kernel spend its time?  Compiles the production kernel with -DOCRF_VS_STAMPS (clock64 of
// thread 0 of every CTA at each phase), runs it on 6 views x ~19 k random depths and prints the phase durations of the
// slowest CTA plus the event-timed kernel duration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../ocrfdet_b200/csrc \
//        -DOCRF_VS_STAMPS -o vsort_probe vsort_probe.cu && ./vsort_probe
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../ocrfdet_b200/csrc/visible_sort.cu"

int main(int argc, char** argv) {
  const int V = 6;
  const int n_per = argc > 1 ? atoi(argv[1]) : 19000;
  const int P = std::max(100000, n_per);  // a view never has more visible Gaussians than its sample has Gaussians:
                                          // the host sizes the kernel's shared memory from P
  std::mt19937 rng(1);
  std::vector<uint32_t> view_start(V + 1);
  for (int v = 0; v <= V; v++) view_start[v] = (uint32_t)v * n_per;
  const size_t N = (size_t)V * n_per;
  std::vector<uint64_t> keys(N);
  std::vector<uint32_t> vals(N), tiles((size_t)V * P), offs((size_t)V * P, 0u);
  std::uniform_real_distribution<float> depth(0.2f, 60.f);
  for (int v = 0; v < V; v++)
    for (int i = 0; i < n_per; i++) {
      float d = depth(rng);
      uint32_t b;
      memcpy(&b, &d, 4);
      keys[(size_t)v * n_per + i] = ((uint64_t)v << 32) | b;
      vals[(size_t)v * n_per + i] = (uint32_t)(v * P + (rng() % P));
    }
  for (auto& t : tiles) t = 1 + rng() % 9;
  uint32_t *d_vs, *d_vals, *d_vals1, *d_tiles, *d_offs, *d_sorted;
  uint64_t *d_keys, *d_keys1;
  long long* d_stamps;
  cudaMalloc(&d_vs, (V + 1) * 4);
  cudaMalloc(&d_keys, N * 8); cudaMalloc(&d_keys1, N * 8);
  cudaMalloc(&d_vals, N * 4); cudaMalloc(&d_vals1, N * 4);
  cudaMalloc(&d_tiles, tiles.size() * 4); cudaMalloc(&d_offs, offs.size() * 4); cudaMalloc(&d_sorted, N * 4);
  cudaMalloc(&d_stamps, V * 8 * 64 * 8);
  cudaMemset(d_stamps, 0, V * 8 * 64 * 8);
  cudaMemcpy(d_vs, view_start.data(), (V + 1) * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_tiles, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_offs, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(ocrf::g_vs_stamps, &d_stamps, sizeof(d_stamps));
  OcrfShape sh = {};
  sh.S = 1; sh.P = P; sh.V = V; sh.views_per_sample = V; sh.W = 704; sh.H = 256; sh.C = 3;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 6; rep++) {
    cudaMemcpy(d_keys, keys.data(), N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_vals, vals.data(), N * 4, cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    int rc = ocrf::visible_sort(0, &sh, d_vs, d_keys, d_vals, d_keys1, d_vals1, d_tiles, d_offs, d_sorted);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rc) { printf("rc %d\n", rc); return 1; }
    best = std::min(best, ms);
  }
  std::vector<uint64_t> out(N);
  cudaMemcpy(out.data(), d_keys, N * 8, cudaMemcpyDeviceToHost);
  bool sorted = true;
  for (int v = 0; v < V; v++)
    for (int i = 1; i < n_per; i++) sorted &= out[(size_t)v * n_per + i - 1] <= out[(size_t)v * n_per + i];
  std::vector<long long> st(V * 8 * 64);
  cudaMemcpy(st.data(), d_stamps, st.size() * 8, cudaMemcpyDeviceToHost);
  printf("n per view %d  sorted %d  kernel (events, best of 6) %.1f us  %s\n", n_per, (int)sorted, best * 1e3f,
         cudaGetErrorString(cudaGetLastError()));
  for (int c = 0; c < 2; c++) {  // view 0 rank 0 and view 5 rank 7
    const long long* s = &st[(c ? (5 * 8 + 7) : 0) * 64];
    printf("CTA %s: total %lld cycles\n", c ? "view 5 rank 7" : "view 0 rank 0", s[37] - s[0]);
    printf("  load %lld\n", s[1] - s[0]);
    for (int p = 0; p < 4; p++) {
      const long long* q = s + 8 * p;
      printf("  pass %d: rank %lld  warp-prefix %lld  cluster-barrier %lld  totals+scan %lld  scatter %lld  cluster-barrier %lld  read-back %lld\n",
             p, q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5], q[7] - q[6], (p < 3 ? q[9] : s[33]) - q[7]);
    }
    printf("  (read-back of pass 3 includes the gathers of the Gaussian indices and tile counts)\n");
    printf("  epilogue: scan inside the CTA %lld  cluster-barrier %lld  write %lld  final barrier %lld\n", s[34] - s[33], s[35] - s[34],
           s[36] - s[35], s[37] - s[36]);
  }
  return 0;
}
