// Stage 4 for many feature channels (32 < C <= 80, BASELINE config 4): the blend backward with its per-(pixel, record)
// C-term dot product on the 5th-generation tensor cores.
//
// The reverse traversal of cuda_rasterizer/backward.cu:399-557 needs, per (pixel p, record j),
//     dot[p][j] = sum_c dL/dpixel[p][c] * feature[j][c]
// (render_backward_generic_kernel: C/2 FFMA2 + C/4 LDS.128 per pair, about half of its instructions).  For a tile that
// is  D1[256 px][R records] = G[256][C] . F^T[C][R],  issued here as tcgen05.mma (kind::tf32, M = 128, N = 32, K = 8):
//   A operand  G = dL/dpixel of the tile, RESIDENT in tensor memory for the whole tile: every thread stores its pixel's
//              row once (g and g_lo = g - trunc_tf32(g): 2 x 2 x C columns)
//   B operand  feature rows of 32 records, K-major in shared memory: a row IS a run of 16-byte chunks of the canonical
//              core-matrix layout, so the loader warp lets cp.async put them in place and only derives the f_lo image
//   D1         2 x (2 x 32) columns, double-buffered: the MMAs of block b + 1 run under the scalar work of block b;
//              a worker reads its pixel's 32 dot products with two tcgen05.ld
//   precision  3xTF32 (g.f + g_lo.f + g.f_lo), ~2^-21 relative per term: inside the 1e-4 gradient bar.
// Everything after the dot product is render_backward_generic_kernel's scheme: scalar recurrence for the colour
// accumulated behind, 8-value butterfly for the geometry gradients, per-warp 3xTF32 mma.sync product for the feature
// gradients, one set of global reductions per record per tile.
// Tensor memory: 4 C + 128 <= 448 columns -> one CTA per SM (8 worker warps + MMA warp + loader warp).
#include "blend_math.cuh"
#include "tc_common.cuh"

namespace ocrf {
namespace tc {

constexpr int FB = 16;  // records per cross-warp reduction of the feature gradients
constexpr int WP = 36;  // pitch of a record's 32 pixel weights (== 4 mod 32: conflict-free A fragments)

// A CTA owns MB M-blocks of 128 pixels: MB = 2 a whole 16x16 tile (448 tensor-memory columns: one CTA per SM), MB = 1
// the upper or lower half of one (192 columns: TWO CTAs per SM, so the start-up of one half tile -- 40 KB of upstream
// gradient, the first feature block, the first products -- runs under the traversal of the other).  Blocks are
// BR = 16 MB records; the feature staging and the global reductions double with half tiles, the products do not.
template <int CP, int MB>
struct BwdSmem {
  static constexpr int BR = 16 * MB;   // records per block == N of the tcgen05.mma
  static constexpr int NWW = 4 * MB;   // worker warps
  static constexpr int KC = CP / 4;  // 16-byte chunks per feature row
  static constexpr uint32_t LBO = 128;          // next 4 channels of the same 8 records
  static constexpr uint32_t SBO = KC * 128;     // next 8 records
  static constexpr uint32_t PART = (BR / 8) * SBO;
  static constexpr int GP = CP + 8;             // g row pitch (== 8 mod 16: conflict-free B fragments)
  alignas(128) unsigned char f[2][2][PART];     // [stage][raw | lo]
  alignas(128) Record rec[2][BR];
  alignas(16) float macc[NWW][FB][8];   // per-warp partial moments of a group of 16 records
  alignas(16) float mom[BR][8];         // moments of the block's records, summed over the warps
  alignas(16) float facc[NWW][FB][CP];
  alignas(16) float g[NWW][32][GP];
  alignas(16) float w[NWW][FB][WP];
  alignas(16) float t[NWW][FB][WP];     // t = G * dL/dalpha per (record, pixel): the operand of the moment product
  alignas(8) uint64_t ready_f[2], free_f[2], ready_d[2], free_d[2], g_ready;
  uint32_t tmem;
};


template <int CP, int MB>
__global__ void __launch_bounds__(MB * 128 + 64, 2 / MB) render_backward_tc_kernel(
    int W, int H, int C, int P, int views_per_sample, const uint2* __restrict__ ranges /* culled lists */,
    const Record* __restrict__ records, const float* __restrict__ feats, const float* __restrict__ bg,
    const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
    const uint32_t* __restrict__ max_contrib, const float* __restrict__ dL_dpix, const float* __restrict__ dL_dopa,
    double* __restrict__ ggrad, float* __restrict__ dL_dfeats) {
  pdl_enter();
  using SM = BwdSmem<CP, MB>;
  extern __shared__ __align__(128) unsigned char smem_raw_b[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw_b);
  constexpr int GP = SM::GP;
  constexpr int BR = SM::BR, NWW = SM::NWW;
  constexpr int WORKERS = NWW * 32;
  constexpr uint32_t D_COL0 = MB * 2 * CP;  // G occupies [0, 2 MB CP): per M-block g then g_lo
  constexpr uint32_t BWD_TMEM_COLS = MB == 2 ? 512 : 256;
  constexpr int HALVES = 2 / MB;
  auto worker_bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x / HALVES, tiles_per_view = tiles_x * gridDim.y;
  const int view = blockIdx.z;
  const int tile_x = blockIdx.x / HALVES, half = blockIdx.x % HALVES;
  const int tile = blockIdx.y * tiles_x + tile_x;
  const size_t vt = (size_t)view * tiles_per_view + tile;
  const int mc = (int)max_contrib[vt];
  if (mc == 0) return;
  const uint2 range = ranges[vt];
  const int nb = (mc + BR - 1) / BR;
  const int s_idx = view / views_per_sample;

  // The workers put their global loads (80 KB of upstream gradient per tile) in flight BEFORE the CTA's set-up barrier:
  // with one CTA per SM nothing else overlaps a tile's start-up.
  const size_t HW = (size_t)H * W;
  const int wt = half * 4 + (warp < NWW ? warp : 0);  // warp of the tile: an 8 x 4 pixel block
  const int px = tile_x * TILE + (wt & 1) * 8 + (lane & 7);
  const int py = blockIdx.y * TILE + (wt >> 1) * 4 + (lane >> 3);
  const bool inside = warp < NWW && px < W && py < H;
  const size_t pix = (size_t)py * W + px;
  float gv[CP];
  float bgdot = 0.f, Tf = 0.f, gop = 0.f;
  int nc = 0;
  if (warp < NWW) {
    const float* gp = dL_dpix + (size_t)view * C * HW + (inside ? pix : 0);  // one pointer, stepped by a channel plane
#pragma unroll
    for (int k = 0; k < CP; k++) {
      gv[k] = (inside && (C == CP || k < C)) ? __ldg(gp) : 0.f;
      if (C == CP || k + 1 < C) gp += HW;
    }
    Tf = inside ? final_T[view * HW + pix] : 0.f;
    nc = inside ? (int)n_contrib[view * HW + pix] : 0;
    gop = (inside && dL_dopa) ? dL_dopa[view * HW + pix] : 0.f;
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
      mbar_init(&sm.ready_f[s], 1);
      mbar_init(&sm.free_f[s], NWW);
      mbar_init(&sm.ready_d[s], 1);
      mbar_init(&sm.free_d[s], NWW);
    }
    mbar_init(&sm.g_ready, NWW);
    mbar_fence_init();
  }
  if (warp == NWW) tmem_alloc<BWD_TMEM_COLS>(&sm.tmem);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = sm.tmem;

  if (warp < NWW) {
    // ------------------------------------------------ workers ------------------------------------------------
    const float* fbase = feats + (size_t)s_idx * P * C;
    (void)fbase;
    float* gfbase = dL_dfeats + (size_t)s_idx * P * C;
    const float fx = (float)px, fy = (float)py;
    const int mb = warp >> 2;
    const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16);
    // this pixel's upstream gradient: to tensor memory (g, g_lo) and to the shared tile the mma.sync product reads
    {
      float* grow = &sm.g[warp][lane][0];
#pragma unroll
      for (int c0 = 0; c0 < CP; c0 += 16) {
        uint32_t raw[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int k = c0 + i;
          bgdot += (k < C ? bg[k] : 0.f) * gv[k];
          raw[i] = __float_as_uint(gv[k]);
          lo[i] = __float_as_uint(tf32_lo(gv[k]));
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(grow + c0 + i) = make_float4(gv[c0 + i], gv[c0 + i + 1], gv[c0 + i + 2], gv[c0 + i + 3]);
        tmem_st16(lane_addr + mb * 2 * CP + c0, raw);
        tmem_st16(lane_addr + mb * 2 * CP + CP + c0, lo);
      }
      tmem_wait_st();
    }
    fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.g_ready);  // this warp's rows of G are in tensor memory (the MMA warp waits for all)
    float S = 0.f;    // g . (colour accumulated behind this pixel), see render_backward_generic_kernel
    float* my_w = &sm.w[warp][0][lane];
    float* my_t = &sm.t[warp][0][lane];
    // B fragments of the moment product: monomial (lane >> 2) of the local coordinates of pixels 8 ks + (lane & 3) (+ 4)
    float bmono[8];
    {
      const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int q = 8 * (e >> 1) + tig + 4 * (e & 1);  // warp-local pixel = the lane that owns it
        const float lx = (float)((wt & 1) * 8 + (q & 7)) - 7.5f, ly = (float)((wt >> 1) * 4 + (q >> 3)) - 7.5f;
        const float xm = (gid == 1 || gid == 4) ? lx : gid == 3 ? lx * lx : 1.f;  // (selects, no branches)
        const float ym = (gid == 2 || gid == 4) ? ly : gid == 5 ? ly * ly : 1.f;
        bmono[e] = gid < 6 ? xm * ym : 0.f;
      }
    }
    const float gob = Tf * (gop - bgdot);
    float T = Tf;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    for (int b = 0; b < nb; b++) {
      const int s = b & 1;
      const uint32_t use = (uint32_t)(b >> 1);
      const int hi = mc - b * BR;
      const int lo = max(0, hi - BR);
      const int cnt = hi - lo;
      mbar_wait_wd(&sm.ready_f[s], use & 1);
      mbar_wait_wd(&sm.ready_d[s], use & 1);
      fence_after_sync();
      // Two groups of 16 records, back to front.  Inside a group the work is laid out for instruction-level
      // parallelism (one CTA per SM: there are only two warps per scheduler to hide latency with):
      //   phase A  alpha, G, 1 / (1 - alpha) of eight records: independent of each other
      //   phase B  the recurrences (T, the colour dot product accumulated behind): a short dependent chain
      //   phase C  t = G dL/dalpha per record; the geometry gradients are six MOMENTS of t over the tile's pixels
      //            (sum t, t x, t y, t x^2, t x y, t y^2 about the tile centre), i.e. one more small product t^T . M
#pragma unroll 1
      for (int grp = BR / FB - 1; grp >= 0; grp--) {
        const int j0 = grp * FB;
        if (j0 >= cnt) continue;  // (CTA-uniform)
        uint32_t d[FB];
        tmem_ld16(lane_addr + D_COL0 + s * (MB * BR) + mb * BR + j0, d);
        tmem_wait_ld();
        if (grp == 0) {  // the last read of this D1 buffer: the MMA warp may overwrite it
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.free_d[s]);
        }
#pragma unroll
        for (int h = 1; h >= 0; h--) {
          float Gv[8], al[8], rc[8], dLa[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int j = j0 + h * 8 + i;
            const float4 a = reinterpret_cast<const float4*>(&sm.rec[s][j])[0];
            const float4 bq = reinterpret_cast<const float4*>(&sm.rec[s][j])[1];
            const float dx = a.x - fx, dy = a.y - fy;
            const float power = a.z * dx * dx + bq.x * dy * dy + a.w * dx * dy;  // log2 domain (scaled conic)
            const float G = ex2_approx_b(power);
            const float alpha = fminf(0.99f, bq.y * G);
            const bool ok = j < cnt && (int)__float_as_uint(bq.z) <= nc && power <= 0.0f && alpha >= 1.0f / 255.0f;
            const float r = rcp_approx(1.f - alpha);  // 1 - alpha in [0.01, 1]: one MUFU.RCP
            Gv[i] = ok ? G : 0.f;  // (a garbage row past the end of the list must not put a NaN into the sums)
            al[i] = ok ? alpha : 0.f;
            rc[i] = ok ? r : 1.f;
          }
#pragma unroll
          for (int i = 7; i >= 0; i--) {
            T *= rc[i];
            const float w = al[i] * T;
            my_w[(h * 8 + i) * WP] = w;  // column `lane` of the warp's w^T tile (0 for a pixel that did not blend)
            // (columns past the end of the list hold whatever the tensor core made of stale rows: never let them in)
            const float dot = al[i] > 0.f ? __uint_as_float(d[h * 8 + i]) - S : 0.f;
            S = fmaf(al[i], dot, S);
            dLa[i] = dot * T + gob * rc[i];
          }
          // phase C: t = G dL/dalpha, the one per-(pixel, record) scalar every geometry gradient is linear in
#pragma unroll
          for (int i = 0; i < 8; i++) my_t[(h * 8 + i) * WP] = al[i] > 0.f ? Gv[i] * dLa[i] : 0.f;
        }
        // ---- records [j0, j0 + FB) are complete in every warp: one reduction over the warps per channel ----
        __syncwarp();
        {  // this warp's partial dF[16 records][CP] = w^T[16 x 32 px] . g[32 px x CP] (3xTF32 mma.sync)
          const int gid = lane >> 2, tig = lane & 3;
          const float* wt = &sm.w[warp][0][0];
          const float* gt = &sm.g[warp][0][0];
          float cacc[CP / 8][4];
#pragma unroll
          for (int nt = 0; nt < CP / 8; nt++) cacc[nt][0] = cacc[nt][1] = cacc[nt][2] = cacc[nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            uint32_t ahi[4], alo[4];
            split_trunc(wt[gid * WP + 8 * ks + tig], ahi[0], alo[0]);
            split_trunc(wt[(gid + 8) * WP + 8 * ks + tig], ahi[1], alo[1]);
            split_trunc(wt[gid * WP + 8 * ks + tig + 4], ahi[2], alo[2]);
            split_trunc(wt[(gid + 8) * WP + 8 * ks + tig + 4], ahi[3], alo[3]);
#pragma unroll
            for (int nt = 0; nt < CP / 8; nt++) {
              uint32_t bhi[2], blo[2];
              split_trunc(gt[(8 * ks + tig) * GP + 8 * nt + gid], bhi[0], blo[0]);
              split_trunc(gt[(8 * ks + tig + 4) * GP + 8 * nt + gid], bhi[1], blo[1]);
              mma_tf32_16x8x8(cacc[nt], alo, bhi);
              mma_tf32_16x8x8(cacc[nt], ahi, blo);
              mma_tf32_16x8x8(cacc[nt], ahi, bhi);
            }
          }
          float* fo = &sm.facc[warp][0][0];
#pragma unroll
          for (int nt = 0; nt < CP / 8; nt++) {
            *reinterpret_cast<float2*>(fo + gid * CP + 8 * nt + 2 * tig) = make_float2(cacc[nt][0], cacc[nt][1]);
            *reinterpret_cast<float2*>(fo + (gid + 8) * CP + 8 * nt + 2 * tig) = make_float2(cacc[nt][2], cacc[nt][3]);
          }
          // this warp's partial moments [16 records][8] = t^T[16 x 32 px] . M[32 px x 8]; the monomials of the local
          // pixel coordinates (multiples of 0.5 below 64) are exact in tf32: two products per k-step
          const float* tt = &sm.t[warp][0][0];
          float macc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            uint32_t ahi[4], alo[4];
            split_trunc(tt[gid * WP + 8 * ks + tig], ahi[0], alo[0]);
            split_trunc(tt[(gid + 8) * WP + 8 * ks + tig], ahi[1], alo[1]);
            split_trunc(tt[gid * WP + 8 * ks + tig + 4], ahi[2], alo[2]);
            split_trunc(tt[(gid + 8) * WP + 8 * ks + tig + 4], ahi[3], alo[3]);
            const uint32_t bm[2] = {__float_as_uint(bmono[2 * ks]), __float_as_uint(bmono[2 * ks + 1])};
            mma_tf32_16x8x8(macc, alo, bm);
            mma_tf32_16x8x8(macc, ahi, bm);
          }
          *reinterpret_cast<float2*>(&sm.macc[warp][gid][2 * tig]) = make_float2(macc[0], macc[1]);
          *reinterpret_cast<float2*>(&sm.macc[warp][gid + 8][2 * tig]) = make_float2(macc[2], macc[3]);
        }
        worker_bar();
        if (tid < FB * 8) {  // moments of the group's records, summed over the warps
          const int jj = tid >> 3, m = tid & 7;
          float tot = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < NWW; w2++) tot += sm.macc[w2][jj][m];
          sm.mom[j0 + jj][m] = tot;
        }
        for (int e = tid; e < FB * CP; e += WORKERS) {
          const int jj = e / CP, ch = e - jj * CP;
          const int rj = j0 + jj;
          if (rj >= cnt || ch >= C) continue;
          float tot = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < NWW; w2++) tot += sm.facc[w2][jj][ch];
          if (tot != 0.f) atomicAdd(gfbase + (size_t)sm.rec[s][rj].id * C + ch, tot);
        }
        worker_bar();
      }
      worker_bar();
      if (tid < cnt) {
        // one thread per record: the six geometry gradients from the moments (fp64: the combination cancels when the
        // Gaussian sits far from the tile), ONE set of global reductions per record per tile
        const int j = tid;
        const float4 m03 = *reinterpret_cast<const float4*>(&sm.mom[j][0]);
        const float2 m45 = *reinterpret_cast<const float2*>(&sm.mom[j][4]);
        if (m03.x != 0.f || m03.y != 0.f || m03.z != 0.f || m03.w != 0.f || m45.x != 0.f || m45.y != 0.f) {
          const Record rc_ = sm.rec[s][j];
          const double M0 = m03.x, Mx = m03.y, My = m03.z, Mxx = m03.w, Mxy = m45.x, Myy = m45.y;
          const double a = (double)rc_.x - ((double)(tile_x * TILE) + 7.5), bb = (double)rc_.y - ((double)(blockIdx.y * TILE) + 7.5);
          const double sdx = a * M0 - Mx, sdy = bb * M0 - My;
          const double sxx = a * a * M0 - 2.0 * a * Mx + Mxx;
          const double sxy = a * bb * M0 - a * My - bb * Mx + Mxy;
          const double syy = bb * bb * M0 - 2.0 * bb * My + Myy;
          const double op = rc_.op, qa = rc_.qa, qb = rc_.qb, qc = rc_.qc;
          double q[6];
          q[0] = op * (double)LN2 * (double)ddelx_dx * (2.0 * qa * sdx + qb * sdy);
          q[1] = op * (double)LN2 * (double)ddely_dy * (2.0 * qc * sdy + qb * sdx);
          q[2] = -0.5 * op * sxx;
          q[3] = -0.5 * op * sxy;
          q[4] = -0.5 * op * syy;
          q[5] = M0;
          double* gg = ggrad + ((size_t)view * P + rc_.id) * OCRF_GGRAD_STRIDE;
#pragma unroll
          for (int e = 0; e < 6; e++) atomicAdd(gg + e, q[e]);
        }
      }
      worker_bar();  // mom / rec[s] are free again
      if (lane == 0) mbar_arrive(&sm.free_f[s]);
    }
  } else if (warp == NWW) {
    // ------------------------------------------------ MMA issuer ------------------------------------------------
    constexpr uint32_t IDESC = idesc_tf32(128, BR);
    mbar_wait_wd(&sm.g_ready, 0);
    fence_after_sync();
    for (int b = 0; b < nb; b++) {
      const int s = b & 1;
      const uint32_t use = (uint32_t)(b >> 1);
      mbar_wait_wd(&sm.ready_f[s], use & 1);
      mbar_wait_wd(&sm.free_d[s], (use & 1) ^ 1);
      fence_after_sync();
      if (lane == 0) {
        const uint32_t f_raw = smem_u32(&sm.f[s][0][0]), f_lo = smem_u32(&sm.f[s][1][0]);
#pragma unroll
        for (int mb = 0; mb < MB; mb++) {
          const uint32_t dcol = tm + D_COL0 + s * (MB * BR) + mb * BR;
          const uint32_t g_raw = tm + mb * 2 * CP, g_lo = g_raw + CP;
#pragma unroll
          for (int ks = 0; ks < CP / 8; ks++) {
            const uint64_t b_raw = smem_desc(f_raw + ks * 2 * SM::LBO, SM::LBO, SM::SBO);
            const uint64_t b_lo = smem_desc(f_lo + ks * 2 * SM::LBO, SM::LBO, SM::SBO);
            mma_ts_tf32(dcol, g_raw + ks * 8, b_raw, IDESC, ks > 0);
            mma_ts_tf32(dcol, g_lo + ks * 8, b_raw, IDESC, 1);
            mma_ts_tf32(dcol, g_raw + ks * 8, b_lo, IDESC, 1);
          }
        }
        mma_commit(&sm.ready_d[s]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ loader ------------------------------------------------
    const Record* src = records + range.x;
    const float* fbase = feats + (size_t)s_idx * P * C;
    auto issue = [&](int b) {
      const int s = b & 1;
      const uint32_t use = (uint32_t)(b >> 1);
      const int hi = mc - b * BR;
      const int lo = max(0, hi - BR);
      const int cnt = hi - lo;
      mbar_wait_wd(&sm.free_f[s], (use & 1) ^ 1);
      if (lane < cnt) {
        const uint32_t id = src[lo + lane].id;
        unsigned char* dst = &sm.f[s][0][(lane & 7) * 16 + (lane >> 3) * SM::SBO];
        const float* row = fbase + (size_t)id * C;
#pragma unroll
        for (int cc = 0; cc < SM::KC; cc++)
          if (4 * cc < C) cp_async16(dst + cc * SM::LBO, row + 4 * cc);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const int t = lane + 32 * i;
        if (t < cnt * 3)
          cp_async16(reinterpret_cast<char*>(&sm.rec[s][0]) + t * 16, reinterpret_cast<const char*>(src + lo) + t * 16);
      }
      cp_async_commit();
    };
    auto finalize = [&](int b) {  // this lane's record: the f_lo image (and the zero padding of the channels)
      const int s = b & 1;
      const int hi = mc - b * BR;
      const int cnt = hi - max(0, hi - BR);
      if (lane < cnt) {
        unsigned char* raw = &sm.f[s][0][(lane & 7) * 16 + (lane >> 3) * SM::SBO];
        unsigned char* lo = &sm.f[s][1][(lane & 7) * 16 + (lane >> 3) * SM::SBO];
#pragma unroll
        for (int cc = 0; cc < SM::KC; cc++) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * cc < C) v = *reinterpret_cast<const float4*>(raw + cc * SM::LBO);
          else *reinterpret_cast<float4*>(raw + cc * SM::LBO) = v;
          *reinterpret_cast<float4*>(lo + cc * SM::LBO) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.ready_f[s]);
    };
    issue(0);
    for (int b = 0; b < nb; b++) {
      if (b + 1 < nb) {
        issue(b + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      finalize(b);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == NWW) tmem_dealloc<BWD_TMEM_COLS>(tm);
}

template <int CP, int MB>
static int launch_backward_tc_cp(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                                 const float* colors, const float* bg, const float* fT, const uint32_t* nc,
                                 const uint32_t* mc, const float* dL_dcolor, const float* dL_dopa, double* ggrad,
                                 float* dL_dcolors) {
  const size_t dyn = sizeof(BwdSmem<CP, MB>);
  static int configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(render_backward_tc_kernel<CP, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return (int)e;
    configured[dev] = 1;
  }
  grid.x *= 2 / MB;
  OCRF_LAUNCH(render_backward_tc_kernel<CP, MB>, dim3(grid), dim3(MB * 128 + 64), dyn, st, sh->W, sh->H, sh->C, sh->P,
              sh->views_per_sample, ranges, rec, colors, bg, fT, nc, mc, dL_dcolor, dL_dopa, ggrad, dL_dcolors);
  return 0;
}

int launch_backward_tc(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                       const float* colors, const float* bg, const float* fT, const uint32_t* nc, const uint32_t* mc,
                       const float* dL_dcolor, const float* dL_dopa, double* ggrad, float* dL_dcolors) {
  // whole tiles by default: half tiles (two CTAs per SM) measured 2.13 against 1.96 ms at BASELINE config 4 -- the
  // kernel is bound by its instruction count (the per-warp mma.sync product is a third of it), not by start-up latency
  static const int mb = (getenv("OCRF_TC_BWD_MB") != nullptr && atoi(getenv("OCRF_TC_BWD_MB")) == 1) ? 1 : 2;
#define OCRF_TCB(CPV)                                                                                                     \
  return mb == 2 ? launch_backward_tc_cp<CPV, 2>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, dL_dcolor, dL_dopa, ggrad, dL_dcolors) \
                 : launch_backward_tc_cp<CPV, 1>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, dL_dcolor, dL_dopa, ggrad, dL_dcolors)
  if (sh->C <= 48) OCRF_TCB(48);
  if (sh->C <= 64) OCRF_TCB(64);
  OCRF_TCB(80);
#undef OCRF_TCB
}

}  // namespace tc
}  // namespace ocrf
