// Stage 5b: OpacityVoxelToBEVConverter + HeightAttention -- lifted per-voxel opacity [B,13,S,S] -> BEV opacity logit
// [B,1,S,S] (view_transformer_ocrf.py:463-518, gates :421-461; called at :1196).
//
// The module is a two-level U-Net of five conv blocks (depthwise 3x3 -> 1x1 -> batch norm -> ReLU), each followed by
// a height-attention gate (per channel slice: global max -> 1x1 -> ReLU -> 1x1 -> sigmoid), with 2x2 max pools, two
// stride-2 transposed convolutions, skip concatenations and a 1x1 output conv; 1847 parameters in total.  In torch it
// is ~60 launches forward (cuDNN algorithm selection included) on tensors of at most 4 x 128 x 128 per sample; with
// training-mode batch norm every block needs a reduction over the whole batch before its activation, and every gate a
// reduction over the whole plane before the next block, which is what really shapes the schedule.
//
// Here a block is TWO kernels -- `conv` (the block's input is never materialised: the pooling, the transposed
// convolution, the gating and the concatenation are evaluated on the fly from the previous activations; depthwise +
// pointwise in registers; batch statistics by block reduction + one atomic per channel per block) and `act` (batch
// norm, ReLU, position add, per-(sample, channel) arg-max as one packed 64-bit atomicMax) -- so the forward is 11
// launches, chained with programmatic dependent launch.  The backward mirrors it with three kernels per block:
// `gout` (gradient of the block's gated output from its consumers: adjoint of pool / transposed conv / concatenation,
// plus the sums the gate needs), `gact` (gate + ReLU backward, batch-norm sums) and `gconv` (batch-norm, pointwise and
// depthwise backward; the depthwise adjoint is a gather over the 3x3 neighbours whose pointwise gradients are
// recomputed, so no gradient tensor of the block's input channels is staged).
// All reductions that feed parameters go through block-level sums and one atomic per value per block.
#include "common.cuh"

namespace ocrf {
namespace hoaconv {

constexpr int NLEV = 5;
constexpr int MAXC = 16;
constexpr float BN_EPS = 1e-5f;
#ifndef OCRF_HOAC_THREADS
#define OCRF_HOAC_THREADS 256
#endif
constexpr int THREADS = OCRF_HOAC_THREADS;

// parameter offsets inside the packed vector (= named_parameters() order of the reference module)
struct LevelDef {
  int cin, cout, shift;           // resolution = S >> shift
  int p_dw, p_bdw, p_pw, p_bpw, p_gam, p_bet, p_gate;
  int kind;                       // 0: the module input; 1: pool(prev); 2: cat(upconv(prev), skip)
  int prev, skip, p_up, p_bup, cup;
};
__constant__ LevelDef LEVELS[NLEV] = {
    {13, 4, 0, 0, 117, 130, 182, 186, 190, 194, 0, -1, -1, 0, 0, 0},            // encoder1 / ca1
    {4, 8, 1, 202, 238, 242, 274, 282, 290, 298, 1, 0, -1, 0, 0, 0},            // encoder2 / ca2
    {8, 16, 2, 330, 402, 410, 538, 554, 570, 586, 1, 1, -1, 0, 0, 0},           // bottleneck / ca_bottleneck
    {16, 8, 1, 1234, 1378, 1394, 1522, 1530, 1538, 1546, 2, 2, 1, 714, 1226, 8},   // decoder2 / ca_dec2 (upconv2)
    {8, 4, 0, 1710, 1782, 1790, 1822, 1826, 1830, 1834, 2, 3, 0, 1578, 1706, 4}};  // decoder1 / ca_dec1 (upconv1)
constexpr int P_WOUT = 1842, P_BOUT = 1846, P_TOTAL = 1847;
static const int H_COUT[NLEV] = {4, 8, 16, 8, 4};
static const int H_CIN[NLEV] = {13, 4, 8, 16, 8};
static const int H_SHIFT[NLEV] = {0, 1, 2, 1, 0};

// Everything a kernel needs, by value.
struct Net {
  int B, S, train, pos_batched;
  const float* x;         // [B,13,S,S]
  const float* pos;       // [1 or B,4,S,S]
  const float* p;         // [P_TOTAL]
  const float* run_mean;  // [NLEV][MAXC] (eval mode)
  const float* run_var;
  float* t[NLEV];         // pre-norm block outputs [B,cout,h,w]
  float* e[NLEV];         // activations after ReLU (+ position at level 0), before the gate
  double* bnsum;          // [NLEV][MAXC][2] sum, sum of squares (training).  fp64: the sums of the blocks arrive in any
                          // order, and in fp32 their last bits -- hence the batch statistics, the activations and,
                          // through near-equal maxima, the ARG-MAX of a gate -- differed from run to run
  float* stats_out;       // optional [NLEV][MAXC][2] copy for the caller (running statistics)
  unsigned long long* maxkey;  // [NLEV][B][MAXC] (ordered value << 32 | ~index)
};

__device__ __forceinline__ uint32_t ordered_bits(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// mean and 1/sqrt(var + eps) of channel c of level l
__device__ __forceinline__ void bn_stats(const Net& n, int l, int c, float& mean, float& inv) {
  if (n.train) {
    const int hw = (n.S >> LEVELS[l].shift) * (n.S >> LEVELS[l].shift);
    const float cnt = (float)n.B * (float)hw;
    const double s = n.bnsum[(l * MAXC + c) * 2], s2 = n.bnsum[(l * MAXC + c) * 2 + 1];
    const double m = s / (double)cnt;
    mean = (float)m;
    const float var = (float)fmax(s2 / (double)cnt - m * m, 0.0);
    inv = rsqrtf(var + BN_EPS);
  } else {
    mean = n.run_mean[l * MAXC + c];
    inv = rsqrtf(n.run_var[l * MAXC + c] + BN_EPS);
  }
}

// gates of level l for sample b into g[cout] (HeightAttention: four slices of cs = cout / 4 channels), optionally the
// hidden state for the backward.  Called by ONE thread per (b) or redundantly -- it is a few dozen flops.
__device__ void gate_values(const Net& n, int l, int b, float* g, float* mx_out, float* pre_out) {
  const LevelDef& L = LEVELS[l];
  const int cs = L.cout / 4;
  float mx[MAXC];
  for (int c = 0; c < L.cout; c++) {
    mx[c] = from_ordered((uint32_t)(n.maxkey[((size_t)l * n.B + b) * MAXC + c] >> 32));
    if (mx_out) mx_out[c] = mx[c];
  }
  for (int s = 0; s < 4; s++) {
    const float* w1 = n.p + L.p_gate + s * 2 * cs * cs;
    const float* w2 = w1 + cs * cs;
    float hid[4];
    for (int o = 0; o < cs; o++) {
      float a = 0.f;
      for (int c = 0; c < cs; c++) a = fmaf(w1[o * cs + c], mx[s * cs + c], a);
      if (pre_out) pre_out[s * cs + o] = a;
      hid[o] = fmaxf(a, 0.f);
    }
    for (int o = 0; o < cs; o++) {
      float a = 0.f;
      for (int c = 0; c < cs; c++) a = fmaf(w2[o * cs + c], hid[c], a);
      g[s * cs + o] = 1.f / (1.f + __expf(-a));
    }
  }
}

// gated activation enc_l[b][c][y][x]
__device__ __forceinline__ float enc(const Net& n, int l, int b, int c, int y, int x, const float* gate) {
  const int hw = n.S >> LEVELS[l].shift;
  return gate[c] * n.e[l][(((size_t)b * LEVELS[l].cout + c) * hw + y) * hw + x];
}

// The (virtual) input of block l, channel c, at (y, x) of its own resolution; gates of the producing levels in smem.
__device__ __forceinline__ float block_input(const Net& n, int l, int b, int c, int y, int x, const float* gate_prev,
                                             const float* gate_skip) {
  const LevelDef& L = LEVELS[l];
  if (L.kind == 0) return n.x[(((size_t)b * L.cin + c) * n.S + y) * n.S + x];
  if (L.kind == 1) {  // 2x2 max pool of the gated previous activations
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) m = fmaxf(m, enc(n, L.prev, b, c, 2 * y + i, 2 * x + j, gate_prev));
    return m;
  }
  if (c < L.cup) {  // ConvTranspose2d(k = 2, s = 2) of the gated previous activations
    const int cprev = LEVELS[L.prev].cout;
    float a = n.p[L.p_bup + c];
    const float* w = n.p + L.p_up + (c * 2 + (y & 1)) * 2 + (x & 1);  // [k][cup][2][2]
    for (int k = 0; k < cprev; k++) a = fmaf(enc(n, L.prev, b, k, y >> 1, x >> 1, gate_prev), w[k * L.cup * 4], a);
    return a;
  }
  return enc(n, L.skip, b, c - L.cup, y, x, gate_skip);
}

__device__ __forceinline__ float warp_sum_c(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sums of `count` per-thread values -> one atomicAdd per value (all threads call; scratch [8][count])
__device__ void block_sums_to(const float* vals, int count, float* scratch, float* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  for (int k = 0; k < count; k++) {
    const float v = warp_sum_c(vals[k]);
    if (lane == 0) scratch[warp * count + k] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < count; k += THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) t += scratch[w * count + k];
    if (t != 0.f) atomicAdd(dst + k, t);
  }
}

// the same with an fp64 destination (a block's own sum is computed in a fixed order; only the order of the blocks varies)
__device__ void block_sums_to_f64(const float* vals, int count, float* scratch, double* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  for (int k = 0; k < count; k++) {
    const float v = warp_sum_c(vals[k]);
    if (lane == 0) scratch[warp * count + k] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < count; k += THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) t += scratch[w * count + k];
    if (t != 0.f) atomicAdd(dst + k, (double)t);
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
// conv: thread = pixel of block l; grid (ceil(h*w / THREADS), B)
__global__ void __launch_bounds__(THREADS) conv_kernel(Net n, int l) {
  pdl_enter();
  __shared__ float s_gp[MAXC], s_gs[MAXC];
  __shared__ float s_red[8 * 2 * MAXC];
  const LevelDef& L = LEVELS[l];
  const int b = blockIdx.y, hw = n.S >> L.shift;
  if (threadIdx.x == 0) {
    if (L.kind >= 1) gate_values(n, L.prev, b, s_gp, nullptr, nullptr);
    if (L.kind == 2) gate_values(n, L.skip, b, s_gs, nullptr, nullptr);
  }
  __syncthreads();
  const int pix = blockIdx.x * THREADS + threadIdx.x;
  const bool live = pix < hw * hw;
  const int y = live ? pix / hw : 0, x = live ? pix - y * hw : 0;
  float t[MAXC];
#pragma unroll
  for (int o = 0; o < MAXC; o++) t[o] = o < L.cout ? n.p[L.p_bpw + o] : 0.f;
  if (live) {
    for (int c = 0; c < L.cin; c++) {
      float d = n.p[L.p_bdw + c];
      const float* wd = n.p + L.p_dw + c * 9;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          const int yy = y + dy, xx = x + dx;
          if (yy < 0 || yy >= hw || xx < 0 || xx >= hw) continue;
          d = fmaf(wd[(dy + 1) * 3 + dx + 1], block_input(n, l, b, c, yy, xx, s_gp, s_gs), d);
        }
#pragma unroll
      for (int o = 0; o < MAXC; o++)
        if (o < L.cout) t[o] = fmaf(n.p[L.p_pw + o * L.cin + c], d, t[o]);
    }
#pragma unroll
    for (int o = 0; o < MAXC; o++)
      if (o < L.cout) n.t[l][(((size_t)b * L.cout + o) * hw + y) * hw + x] = t[o];
  }
  if (n.train) {
    float v[2 * MAXC];
#pragma unroll
    for (int o = 0; o < MAXC; o++) {
      const float tv = (live && o < L.cout) ? t[o] : 0.f;
      v[2 * o] = tv;
      v[2 * o + 1] = tv * tv;
    }
    block_sums_to_f64(v, 2 * L.cout, s_red, n.bnsum + (size_t)l * MAXC * 2);
  }
}

// act: e = relu(bn(t)) (+ position at level 0); per (sample, channel) arg-max.  grid (ceil(h*w / THREADS), cout, B)
__global__ void __launch_bounds__(THREADS) act_kernel(Net n, int l) {
  pdl_enter();
  __shared__ unsigned long long s_key[THREADS / 32];
  const LevelDef& L = LEVELS[l];
  const int b = blockIdx.z, c = blockIdx.y, hw = n.S >> L.shift;
  const int pix = blockIdx.x * THREADS + threadIdx.x;
  float mean, inv;
  bn_stats(n, l, c, mean, inv);
  unsigned long long key = 0ull;
  if (pix < hw * hw) {
    const size_t idx = ((size_t)b * L.cout + c) * hw * hw + pix;
    float v = fmaxf((n.t[l][idx] - mean) * inv * n.p[L.p_gam + c] + n.p[L.p_bet + c], 0.f);
    if (l == 0) v += n.pos[((size_t)(n.pos_batched ? b : 0) * L.cout + c) * hw * hw + pix];
    n.e[l][idx] = v;
    key = ((unsigned long long)ordered_bits(v) << 32) | (uint32_t)(~(uint32_t)pix);  // ties: the FIRST pixel wins
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < THREADS / 32; w++) key = s_key[w] > key ? s_key[w] : key;
    atomicMax(&n.maxkey[((size_t)l * n.B + b) * MAXC + c], key);
  }
}

// out = output_conv(gated decoder1): grid (ceil(S*S / THREADS), B)
__global__ void __launch_bounds__(THREADS) out_kernel(Net n, float* __restrict__ out) {
  pdl_enter();
  __shared__ float s_g[MAXC];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) gate_values(n, NLEV - 1, b, s_g, nullptr, nullptr);
  if (n.stats_out != nullptr && blockIdx.x == 0 && b == 0)  // the batch statistics for the caller (every block is done)
    for (int i = threadIdx.x; i < NLEV * MAXC * 2; i += THREADS) n.stats_out[i] = (float)n.bnsum[i];
  __syncthreads();
  const int pix = blockIdx.x * THREADS + threadIdx.x, hw = n.S;
  if (pix >= hw * hw) return;
  float a = n.p[P_BOUT];
#pragma unroll
  for (int c = 0; c < 4; c++) a = fmaf(n.p[P_WOUT + c], s_g[c] * n.e[NLEV - 1][((size_t)b * 4 + c) * hw * hw + pix], a);
  out[(size_t)b * hw * hw + pix] = a;
}

// ---- backward -------------------------------------------------------------------------------------------------------
struct Grad {
  const float* g_out;    // [B,1,S,S]
  float* genc[NLEV];     // gradient of the gated block output, later (in place) of the pre-ReLU value z
  float* gin[NLEV];      // gradient of the block's (virtual) input channels [B,cin,h,w]
  float* gsum;           // [NLEV][B][MAXC] sum over the plane of genc * e  (-> d/d gate)
  float* bsum;           // [NLEV][MAXC][2]  sum g_z, sum g_z * xhat
  float* g_p;            // [P_TOTAL] accumulated
  float* g_pos;          // [1 or B,4,S,S] accumulated when shared over the batch
  float* g_x;            // [B,13,S,S]
};

// gradient of enc_l (the gated output of block l) from its consumers; also sum(genc * e) per (b, c).
// grid (ceil(h*w / THREADS), cout, B)
__global__ void __launch_bounds__(THREADS) gout_kernel(Net n, Grad g, int l) {
  pdl_enter();
  __shared__ float s_g[MAXC];
  __shared__ float s_red[8 * 8];
  const LevelDef& L = LEVELS[l];
  const int b = blockIdx.z, c = blockIdx.y, hw = n.S >> L.shift;
  if (threadIdx.x == 0) gate_values(n, l, b, s_g, nullptr, nullptr);
  __syncthreads();
  const int pix = blockIdx.x * THREADS + threadIdx.x;
  const bool live = pix < hw * hw;
  const int y = live ? pix / hw : 0, x = live ? pix - y * hw : 0;
  float ge = 0.f, par[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int npar = 0;
  float* par_dst = nullptr;
  const size_t idx = ((size_t)b * L.cout + c) * hw * hw + pix;
  if (l == NLEV - 1) {  // output conv
    const float go = live ? g.g_out[(size_t)b * hw * hw + pix] : 0.f;
    ge = n.p[P_WOUT + c] * go;
    par[0] = live ? go * s_g[c] * n.e[l][idx] : 0.f;  // d/d w_out[c]
    par[1] = c == 0 ? go : 0.f;                        // d/d b_out (counted once)
    npar = 2;
  } else {
    // consumers of level l: a transposed conv (levels 2 -> 3, 3 -> 4), and/or a pool (0 -> 1, 1 -> 2) + a skip (1 -> 3, 0 -> 4)
    const int up_l = (l == 2) ? 3 : (l == 3 ? 4 : -1);
    const int pool_l = (l == 0) ? 1 : (l == 1 ? 2 : -1);
    const int skip_l = (l == 1) ? 3 : (l == 0 ? 4 : -1);
    if (up_l >= 0 && live) {  // adjoint of ConvTranspose2d: channel c here is input channel k = c of the upconv
      const LevelDef& U = LEVELS[up_l];
      const int hw2 = hw * 2;
      for (int oc = 0; oc < U.cup; oc++)
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < 2; j++)
            ge = fmaf(n.p[U.p_up + ((c * U.cup + oc) * 2 + i) * 2 + j],
                      g.gin[up_l][(((size_t)b * U.cin + oc) * hw2 + 2 * y + i) * hw2 + 2 * x + j], ge);
    }
    if (pool_l >= 0 && live) {  // adjoint of the 2x2 max pool: the first maximum of the window receives the gradient
      const int hwp = hw / 2, wy = y >> 1, wx = x >> 1;
      float best = -INFINITY;
      int bi = 0;
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const float v = s_g[c] * n.e[l][(((size_t)b * L.cout + c) * hw + 2 * wy + i) * hw + 2 * wx + j];
          if (v > best) { best = v; bi = i * 2 + j; }
        }
      if (bi == (y & 1) * 2 + (x & 1)) ge += g.gin[pool_l][(((size_t)b * LEVELS[pool_l].cin + c) * hwp + wy) * hwp + wx];
    }
    if (skip_l >= 0 && live) {
      const LevelDef& K = LEVELS[skip_l];
      ge += g.gin[skip_l][(((size_t)b * K.cin + K.cup + c) * hw + y) * hw + x];
    }
  }
  if (live) g.genc[l][idx] = ge;
  par[npar] = live ? ge * n.e[l][idx] : 0.f;  // -> d/d gate[b][c]
  // the two (or one) parameter partials go to g_p, the gate sum to gsum: reduce all, route by index
  float v[3] = {par[0], par[1], par[npar]};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 3; k++) {
    const float r = warp_sum_c(v[k]);
    if (lane == 0) s_red[warp * 3 + k] = r;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float tsum = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) tsum += s_red[w * 3 + threadIdx.x];
    if (threadIdx.x == 2) atomicAdd(&g.gsum[((size_t)l * n.B + b) * MAXC + c], tsum);
    else if (l == NLEV - 1 && tsum != 0.f) atomicAdd(&g.g_p[threadIdx.x == 0 ? P_WOUT + c : P_BOUT], tsum);
  }
  (void)par_dst;
}

// parameter gradients of the transposed convolutions: d/dW[k][oc][i][j] = sum enc_prev[k][y][x] gin[oc][2y+i][2x+j],
// d/db[oc] = sum gin[oc].  One block per (k, oc) pair (+ cup blocks for the bias); level = the CONSUMER (3 or 4).
__global__ void __launch_bounds__(THREADS) upconv_grad_kernel(Net n, Grad g, int l) {
  pdl_enter();
  __shared__ float s_red[8 * 4];
  __shared__ float s_gate[MAXC];
  const LevelDef& U = LEVELS[l];
  const int cprev = LEVELS[U.prev].cout, hwp = n.S >> LEVELS[U.prev].shift, hw = hwp * 2;
  const int pair = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (pair < cprev * U.cup) {
    const int k = pair / U.cup, oc = pair - k * U.cup;
    for (int b = 0; b < n.B; b++) {
      __syncthreads();
      if (threadIdx.x == 0) gate_values(n, U.prev, b, s_gate, nullptr, nullptr);
      __syncthreads();
      for (int pix = threadIdx.x; pix < hwp * hwp; pix += THREADS) {
        const int y = pix / hwp, x = pix - y * hwp;
        const float ev = enc(n, U.prev, b, k, y, x, s_gate);
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < 2; j++)
            acc[i * 2 + j] = fmaf(ev, g.gin[l][(((size_t)b * U.cin + oc) * hw + 2 * y + i) * hw + 2 * x + j], acc[i * 2 + j]);
      }
    }
    block_sums_to(acc, 4, s_red, g.g_p + U.p_up + (size_t)pair * 4);
  } else {
    const int oc = pair - cprev * U.cup;
    for (int b = 0; b < n.B; b++)
      for (int pix = threadIdx.x; pix < hw * hw; pix += THREADS) acc[0] += g.gin[l][((size_t)b * U.cin + oc) * hw * hw + pix];
    block_sums_to(acc, 1, s_red, g.g_p + U.p_bup + oc);
  }
}

// gate + ReLU backward: genc (in place) becomes g_z; accumulates the batch-norm sums and the gate parameter gradients.
// grid (ceil(h*w / THREADS), cout, B)
__global__ void __launch_bounds__(THREADS) gact_kernel(Net n, Grad g, int l) {
  pdl_enter();
  __shared__ float s_gate[MAXC], s_gmx[MAXC];
  __shared__ float s_red[8 * 2];
  const LevelDef& L = LEVELS[l];
  const int b = blockIdx.z, c = blockIdx.y, hw = n.S >> L.shift, cs = L.cout / 4;
  if (threadIdx.x == 0) {
    // HeightAttention backward for sample b (every block of the sample repeats these few flops; the parameter
    // gradients are added by the block with blockIdx.x == 0 and c == 0 only)
    float mx[MAXC], pre[MAXC];
    gate_values(n, l, b, s_gate, mx, pre);
    const bool owner = blockIdx.x == 0 && c == 0;
    for (int s = 0; s < 4; s++) {
      const float* w1 = n.p + L.p_gate + s * 2 * cs * cs;
      const float* w2 = w1 + cs * cs;
      float go[4], ghid[4] = {0.f, 0.f, 0.f, 0.f};
      for (int o = 0; o < cs; o++) {
        const float gt = s_gate[s * cs + o];
        go[o] = g.gsum[((size_t)l * n.B + b) * MAXC + s * cs + o] * gt * (1.f - gt);
      }
      for (int o = 0; o < cs; o++)
        for (int k = 0; k < cs; k++) {
          if (owner) atomicAdd(&g.g_p[L.p_gate + s * 2 * cs * cs + cs * cs + o * cs + k], go[o] * fmaxf(pre[s * cs + k], 0.f));
          ghid[k] = fmaf(go[o], w2[o * cs + k], ghid[k]);
        }
      float gm[4] = {0.f, 0.f, 0.f, 0.f};
      for (int o = 0; o < cs; o++) {
        const float gp = pre[s * cs + o] > 0.f ? ghid[o] : 0.f;
        for (int k = 0; k < cs; k++) {
          if (owner) atomicAdd(&g.g_p[L.p_gate + s * 2 * cs * cs + o * cs + k], gp * mx[s * cs + k]);
          gm[k] = fmaf(gp, w1[o * cs + k], gm[k]);
        }
      }
      for (int k = 0; k < cs; k++) s_gmx[s * cs + k] = gm[k];
    }
  }
  __syncthreads();
  const int pix = blockIdx.x * THREADS + threadIdx.x;
  float mean, inv;
  bn_stats(n, l, c, mean, inv);
  float v[2] = {0.f, 0.f};
  if (pix < hw * hw) {
    const size_t idx = ((size_t)b * L.cout + c) * hw * hw + pix;
    const uint32_t amax = ~(uint32_t)(n.maxkey[((size_t)l * n.B + b) * MAXC + c] & 0xffffffffull);
    float ge = s_gate[c] * g.genc[l][idx] + ((uint32_t)pix == amax ? s_gmx[c] : 0.f);  // d/d (e [+ position])
    if (l == 0) {
      float* gp = g.g_pos + ((size_t)(n.pos_batched ? b : 0) * L.cout + c) * hw * hw + pix;
      if (n.pos_batched) *gp = ge; else atomicAdd(gp, ge);
    }
    const float xhat = (n.t[l][idx] - mean) * inv;
    const float z = xhat * n.p[L.p_gam + c] + n.p[L.p_bet + c];
    const float gz = z > 0.f ? ge : 0.f;
    g.genc[l][idx] = gz;
    v[0] = gz;
    v[1] = gz * xhat;
  }
  block_sums_to(v, 2, s_red, g.bsum + ((size_t)l * MAXC + c) * 2);
}

// batch norm + pointwise + depthwise backward.  thread = pixel; grid (ceil(h*w / THREADS), B)
__global__ void __launch_bounds__(THREADS) gconv_kernel(Net n, Grad g, int l) {
  pdl_enter();
  __shared__ float s_gp[MAXC], s_gs[MAXC];
  __shared__ float s_a[MAXC], s_b[MAXC], s_c[MAXC], s_mean[MAXC], s_inv[MAXC];  // g_t = a gz + b + c xhat
  __shared__ float s_red[8 * 32];
  const LevelDef& L = LEVELS[l];
  const int b = blockIdx.y, hw = n.S >> L.shift;
  if (threadIdx.x == 0) {
    if (L.kind >= 1) gate_values(n, L.prev, b, s_gp, nullptr, nullptr);
    if (L.kind == 2) gate_values(n, L.skip, b, s_gs, nullptr, nullptr);
  }
  if (threadIdx.x < L.cout) {
    const int o = threadIdx.x;
    float mean, inv;
    bn_stats(n, l, o, mean, inv);
    const float gam = n.p[L.p_gam + o];
    s_mean[o] = mean;
    s_inv[o] = inv;
    s_a[o] = gam * inv;
    if (n.train) {
      const float cnt = (float)n.B * (float)(hw * hw);
      s_b[o] = -gam * inv * g.bsum[((size_t)l * MAXC + o) * 2] / cnt;
      s_c[o] = -gam * inv * g.bsum[((size_t)l * MAXC + o) * 2 + 1] / cnt;
    } else {
      s_b[o] = 0.f;
      s_c[o] = 0.f;
    }
    if (blockIdx.x == 0 && b == 0) {  // d/d gamma, d/d beta are the batch-norm sums themselves
      atomicAdd(&g.g_p[L.p_gam + o], g.bsum[((size_t)l * MAXC + o) * 2 + 1]);
      atomicAdd(&g.g_p[L.p_bet + o], g.bsum[((size_t)l * MAXC + o) * 2]);
    }
  }
  __syncthreads();
  const int pix = blockIdx.x * THREADS + threadIdx.x;
  const bool live = pix < hw * hw;
  const int y = live ? pix / hw : 0, x = live ? pix - y * hw : 0;
  // g_t at a pixel (all output channels)
  auto gt_at = [&](int yy, int xx, float* gt) {
#pragma unroll
    for (int o = 0; o < MAXC; o++) {
      if (o >= L.cout) { gt[o] = 0.f; continue; }
      const size_t idx = (((size_t)b * L.cout + o) * hw + yy) * hw + xx;
      const float xhat = (n.t[l][idx] - s_mean[o]) * s_inv[o];
      gt[o] = s_a[o] * g.genc[l][idx] + s_b[o] + s_c[o] * xhat;
    }
  };
  float gin[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; c++) gin[c] = 0.f;
  float gtc[MAXC];  // centre pixel
  if (live) {
    // depthwise adjoint: g_in[c][p] = sum over taps of w_dw[c][tap] g_d[c][p - tap], g_d = W_pw^T g_t
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        const int yy = y - dy, xx = x - dx;  // the output pixel whose tap (dy, dx) reads me
        if (yy < 0 || yy >= hw || xx < 0 || xx >= hw) continue;
        float gt[MAXC];
        gt_at(yy, xx, gt);
        if (dy == 0 && dx == 0) {
#pragma unroll
          for (int o = 0; o < MAXC; o++) gtc[o] = gt[o];
        }
        for (int c = 0; c < L.cin; c++) {
          float gd = 0.f;
#pragma unroll
          for (int o = 0; o < MAXC; o++)
            if (o < L.cout) gd = fmaf(n.p[L.p_pw + o * L.cin + c], gt[o], gd);
          gin[c] = fmaf(n.p[L.p_dw + c * 9 + (dy + 1) * 3 + dx + 1], gd, gin[c]);
        }
      }
    for (int c = 0; c < L.cin; c++) {
      float* dst = (l == 0 ? g.g_x : g.gin[l]) + (((size_t)b * L.cin + c) * hw + y) * hw + x;
      *dst = gin[c];
    }
  } else {
#pragma unroll
    for (int o = 0; o < MAXC; o++) gtc[o] = 0.f;
  }
  // parameter gradients: pointwise bias / weights and depthwise bias / weights, channel by channel
  {
    float v[MAXC];
#pragma unroll
    for (int o = 0; o < MAXC; o++) v[o] = gtc[o];
    block_sums_to(v, L.cout, s_red, g.g_p + L.p_bpw);
  }
  for (int c = 0; c < L.cin; c++) {
    // d[c] (depthwise output) and the nine input taps at my pixel
    float tap[9], d = 0.f, gd = 0.f;
    if (live) {
      d = n.p[L.p_bdw + c];
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          const int yy = y + dy, xx = x + dx, k = (dy + 1) * 3 + dx + 1;
          tap[k] = (yy < 0 || yy >= hw || xx < 0 || xx >= hw) ? 0.f : block_input(n, l, b, c, yy, xx, s_gp, s_gs);
          d = fmaf(n.p[L.p_dw + c * 9 + k], tap[k], d);
        }
#pragma unroll
      for (int o = 0; o < MAXC; o++)
        if (o < L.cout) gd = fmaf(n.p[L.p_pw + o * L.cin + c], gtc[o], gd);
    } else {
#pragma unroll
      for (int k = 0; k < 9; k++) tap[k] = 0.f;
    }
    float v[MAXC + 10];
#pragma unroll
    for (int o = 0; o < MAXC; o++) v[o] = gtc[o] * d;  // d/d W_pw[o][c]
#pragma unroll
    for (int k = 0; k < 9; k++) v[MAXC + k] = gd * tap[k];  // d/d w_dw[c][k]
    v[MAXC + 9] = gd;                                        // d/d b_dw[c]
    // route: the three destinations are not contiguous -> three reductions sharing one scratch
    {
      float pw[MAXC];
#pragma unroll
      for (int o = 0; o < MAXC; o++) pw[o] = v[o];
      // W_pw is [cout][cin]: column c has stride cin -> reduce into a small staging row, then scatter
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      __syncthreads();
      for (int o = 0; o < L.cout; o++) {
        const float r = warp_sum_c(pw[o]);
        if (lane == 0) s_red[warp * 32 + o] = r;
      }
      for (int k = 0; k < 10; k++) {
        const float r = warp_sum_c(v[MAXC + k]);
        if (lane == 0) s_red[warp * 32 + MAXC + k] = r;
      }
      __syncthreads();
      if (threadIdx.x < MAXC + 10) {
        const int k = threadIdx.x;
        if (k < L.cout || k >= MAXC) {
          float tsum = 0.f;
#pragma unroll
          for (int w = 0; w < THREADS / 32; w++) tsum += s_red[w * 32 + k];
          float* dst = k < MAXC ? g.g_p + L.p_pw + k * L.cin + c
                                : (k < MAXC + 9 ? g.g_p + L.p_dw + c * 9 + (k - MAXC) : g.g_p + L.p_bdw + c);
          if (tsum != 0.f) atomicAdd(dst, tsum);
        }
      }
    }
  }
}

}  // namespace hoaconv
}  // namespace ocrf

using namespace ocrf;
using namespace ocrf::hoaconv;

// workspace (floats): per level t and e, then the gradient buffers genc and gin, then the small state
struct WsLayout {
  size_t t[NLEV], e[NLEV], genc[NLEV], gin[NLEV], small, total;
};
static WsLayout ws_layout(int B, int S) {
  WsLayout w;
  size_t off = 0;
  for (int l = 0; l < NLEV; l++) {
    const size_t hw = (size_t)(S >> H_SHIFT[l]) * (S >> H_SHIFT[l]);
    w.t[l] = off; off += (size_t)B * H_COUT[l] * hw;
    w.e[l] = off; off += (size_t)B * H_COUT[l] * hw;
  }
  for (int l = 0; l < NLEV; l++) {
    const size_t hw = (size_t)(S >> H_SHIFT[l]) * (S >> H_SHIFT[l]);
    w.genc[l] = off; off += (size_t)B * H_COUT[l] * hw;
    w.gin[l] = off; off += (size_t)B * H_CIN[l] * hw;
  }
  off = (off + 31) & ~size_t(31);
  w.small = off;  // bnsum f64 [5][16][2] | bsum [5][16][2] | gsum [5][B][16] | maxkey u64 [5][B][16]
  off += 3 * NLEV * MAXC * 2 + (size_t)NLEV * B * MAXC + 2 * (size_t)NLEV * B * MAXC;
  w.total = off + 32;
  return w;
}

static void bind(Net& n, Grad* g, const WsLayout& w, float* ws, int B) {
  for (int l = 0; l < NLEV; l++) {
    n.t[l] = ws + w.t[l];
    n.e[l] = ws + w.e[l];
    if (g) { g->genc[l] = ws + w.genc[l]; g->gin[l] = ws + w.gin[l]; }
  }
  float* sm = ws + w.small;
  n.bnsum = reinterpret_cast<double*>(sm);
  float* bsum = sm + 2 * NLEV * MAXC * 2;
  float* gsum = bsum + NLEV * MAXC * 2;
  n.maxkey = reinterpret_cast<unsigned long long*>(gsum + (size_t)NLEV * B * MAXC);
  if (g) { g->bsum = bsum; g->gsum = gsum; }
}

extern "C" size_t ocrf_hoa_converter_workspace_floats(int32_t B, int32_t S) {
  if (B <= 0 || S < 8 || (S & 3)) return 0;
  return ws_layout(B, S).total;
}

extern "C" int ocrf_hoa_converter_forward(void* stream, int32_t B, int32_t S, int32_t train, const float* x,
                                          const float* position, int32_t position_batched, const float* params,
                                          const float* running_mean, const float* running_var, float* out,
                                          float* batch_stats, float* workspace) {
  if (!x || !position || !params || !out || !workspace || B <= 0 || S < 8 || (S & 3)) return OCRF_EINVAL;
  if (!train && (!running_mean || !running_var)) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WsLayout w = ws_layout(B, S);
  Net n = {};
  n.B = B; n.S = S; n.train = train ? 1 : 0; n.pos_batched = position_batched ? 1 : 0;
  n.x = x; n.pos = position; n.p = params; n.run_mean = running_mean; n.run_var = running_var;
  bind(n, nullptr, w, workspace, B);
  n.stats_out = (train && batch_stats) ? batch_stats : nullptr;  // sum / sum of squares per level and channel: the host
                                                                 // derives mean, var (running statistics); out_kernel writes it
  cudaError_t e = cudaMemsetAsync(workspace + w.small, 0, (w.total - 32 - w.small) * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  for (int l = 0; l < NLEV; l++) {
    const int hw = S >> H_SHIFT[l];
    const unsigned gx = (unsigned)((hw * hw + THREADS - 1) / THREADS);
    OCRF_LAUNCH(conv_kernel, dim3(gx, B), dim3(THREADS), 0, st, n, l);
    OCRF_LAUNCH(act_kernel, dim3(gx, H_COUT[l], B), dim3(THREADS), 0, st, n, l);
  }
  OCRF_LAUNCH(out_kernel, dim3((unsigned)((S * S + THREADS - 1) / THREADS), B), dim3(THREADS), 0, st, n, out);
  OCRF_CHECK_LAST();
  return 0;
}

extern "C" int ocrf_hoa_converter_backward(void* stream, int32_t B, int32_t S, int32_t train, const float* x,
                                           const float* position, int32_t position_batched, const float* params,
                                           const float* running_mean, const float* running_var, const float* g_out,
                                           float* g_x, float* g_position, float* g_params, float* workspace) {
  if (!x || !position || !params || !g_out || !g_x || !g_position || !g_params || !workspace) return OCRF_EINVAL;
  if (B <= 0 || S < 8 || (S & 3)) return OCRF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WsLayout w = ws_layout(B, S);
  Net n = {};
  Grad g = {};
  n.B = B; n.S = S; n.train = train ? 1 : 0; n.pos_batched = position_batched ? 1 : 0;
  n.x = x; n.pos = position; n.p = params; n.run_mean = running_mean; n.run_var = running_var;
  bind(n, &g, w, workspace, B);
  g.g_out = g_out; g.g_p = g_params; g.g_pos = g_position; g.g_x = g_x;
  // the forward's t, e, bnsum and maxkey are still in the workspace; clear only the backward accumulators
  cudaError_t e = cudaMemsetAsync(g.bsum, 0, (NLEV * MAXC * 2 + (size_t)NLEV * B * MAXC) * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (!position_batched) {
    e = cudaMemsetAsync(g_position, 0, (size_t)4 * S * S * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
  }
  for (int l = NLEV - 1; l >= 0; l--) {
    const int hw = S >> H_SHIFT[l];
    const unsigned gx = (unsigned)((hw * hw + THREADS - 1) / THREADS);
    OCRF_LAUNCH(gout_kernel, dim3(gx, H_COUT[l], B), dim3(THREADS), 0, st, n, g, l);
    OCRF_LAUNCH(gact_kernel, dim3(gx, H_COUT[l], B), dim3(THREADS), 0, st, n, g, l);
    OCRF_LAUNCH(gconv_kernel, dim3(gx, B), dim3(THREADS), 0, st, n, g, l);
    if (l >= 3) {  // its input contains a transposed convolution: that layer's parameters
      const int pairs = H_COUT[l - 1] * (l == 3 ? 8 : 4) + (l == 3 ? 8 : 4);
      OCRF_LAUNCH(upconv_grad_kernel, dim3((unsigned)pairs), dim3(THREADS), 0, st, n, g, l);
    }
  }
  OCRF_CHECK_LAST();
  return 0;
}
