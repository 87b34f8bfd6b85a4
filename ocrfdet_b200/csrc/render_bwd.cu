// Stage 4: per-tile reverse-order backward of the blend.
//
// Replaces renderCUDA<3> backward (cuda_rasterizer/backward.cu:399-557), whose nine global
// atomicAdd per (pixel, Gaussian) pair are its dominant cost.  Here
//   * the tile's records arrive back-to-front through the same mbarrier/bulk-copy ring as forward,
//     and the traversal starts at the tile's largest n_contrib instead of the end of its list;
//   * per record, the lanes of a warp that actually blended it reduce their nine partial gradients
//     with a 14-shuffle butterfly (8-value reduce-scatter + one scalar), only when at least one
//     lane contributed;
//   * warps then combine in shared memory, and each record issues ONE set of global reductions per
//     tile: six fp64 reds for (dmean2D, dconic, dopacity) and three scalar
//     reds for the colour -- 256x fewer global atomics than the reference in the dense case.
// The arithmetic follows SURVEY.md appendix A5 (T recovered by division, suffix colour recurrence,
// background term, no clamp mask) plus the opacity-map term  +T_final/(1-alpha) * dL/dO.
// Summation order differs from the reference's unordered atomics: gradients agree to ~1e-5 rel.
#include "blend_math.cuh"

namespace ocrf {

template <int PPT> struct BwdBatch { static constexpr int value = PPT == 1 ? 64 : 128; };  // records per stage
constexpr int BWD_ACC = 12;  // floats per record accumulator row (9 used)

template <int PPT>
__global__ void __launch_bounds__(TILE_PIX / PPT) render_backward_c3_scalar_kernel(
    int W, int H, int P, int views_per_sample, int colors_per_view, const uint2* __restrict__ ranges /* culled lists */,
    const Record* __restrict__ records, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ max_contrib,
    const float* __restrict__ dL_dpix, const float* __restrict__ dL_dopa, double* __restrict__ ggrad,
    float* __restrict__ dL_dcolors) {
  pdl_enter();
  constexpr int NT = TILE_PIX / PPT;
  constexpr int NW = NT / 32;
  constexpr int BWD_BATCH = BwdBatch<PPT>::value;
  // Per-warp partial sums of the current batch: every warp visits every record exactly once per
  // batch, so it can STORE its reduced 9-vector (no shared-memory atomics -- float atomicAdd on
  // shared memory compiles to a CAS loop) and a per-warp bit mask says which rows are valid.
  __shared__ __align__(128) Record s_rec[2][BWD_BATCH];
  __shared__ __align__(16) float s_acc[NW][BWD_BATCH][BWD_ACC];
  __shared__ uint32_t s_touched[NW][BWD_BATCH / 32];
  __shared__ __align__(8) uint64_t s_bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x, tiles_per_view = gridDim.x * gridDim.y;
  const int view = blockIdx.z;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const size_t vt = (size_t)view * tiles_per_view + tile;
  const int mc = (int)max_contrib[vt];
  if (mc == 0) return;
  const uint2 range = ranges[vt];
  const Record* src = records + range.x;
  const int rounds = (mc + BWD_BATCH - 1) / BWD_BATCH;

  const size_t HW = (size_t)H * W;
  const int bx = (warp & 1) * 8 + (lane & 7);
  const int by = (warp >> 1) * (4 * PPT) + (lane >> 3);
  float fx[PPT], fy[PPT], T[PPT], Tf[PPT], g0[PPT], g1[PPT], g2[PPT], gob[PPT];
  float a0[PPT], a1[PPT], a2[PPT], lc0[PPT], lc1[PPT], lc2[PPT], last_alpha[PPT];
  int nc[PPT];
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
#pragma unroll
  for (int p = 0; p < PPT; p++) {
    const int px = blockIdx.x * TILE + bx, py = blockIdx.y * TILE + by + 4 * p;
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)py * W + px;
    fx[p] = (float)px;
    fy[p] = (float)py;
    Tf[p] = inside ? final_T[view * HW + pix] : 0.f;
    T[p] = Tf[p];
    nc[p] = inside ? (int)n_contrib[view * HW + pix] : 0;
    g0[p] = inside ? dL_dpix[((size_t)view * 3 + 0) * HW + pix] : 0.f;
    g1[p] = inside ? dL_dpix[((size_t)view * 3 + 1) * HW + pix] : 0.f;
    g2[p] = inside ? dL_dpix[((size_t)view * 3 + 2) * HW + pix] : 0.f;
    const float gop = (inside && dL_dopa) ? dL_dopa[view * HW + pix] : 0.f;
    // d(out)/d(alpha_i) through the final transmittance: opacity map (+) and background (-)
    gob[p] = Tf[p] * (gop - (bg0 * g0[p] + bg1 * g1[p] + bg2 * g2[p]));
    a0[p] = a1[p] = a2[p] = lc0[p] = lc1[p] = lc2[p] = last_alpha[p] = 0.f;
  }
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

  // round r covers culled-list indices [lo_r, hi_r) with hi_r = mc - r*BATCH (back to front)
  auto issue = [&](int r) {
    const int hi = mc - r * BWD_BATCH;
    const int lo = max(0, hi - BWD_BATCH);
    const uint32_t bytes = (uint32_t)(hi - lo) * sizeof(Record);
    mbar_expect_tx(&s_bar[r & 1], bytes);
    bulk_g2s(&s_rec[r & 1][0], src + lo, bytes, &s_bar[r & 1]);
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    issue(0);
    if (rounds > 1) issue(1);
  }

  for (int r = 0; r < rounds; r++) {
    const int st = r & 1;
    const int hi = mc - r * BWD_BATCH;
    const int lo = max(0, hi - BWD_BATCH);
    const int cnt = hi - lo;
    mbar_wait(&s_bar[st], (r >> 1) & 1);
    const float4* rec4 = reinterpret_cast<const float4*>(&s_rec[st][0]);
    uint32_t touched = 0;  // warp-uniform: bit (j & 31) set when this warp reduced record j

    for (int j = cnt - 1; j >= 0; j--) {
      // pixel p blended this record iff its reference list position (b.z) <= n_contrib[p] and the tests pass
      const float4 a = rec4[3 * j], b = rec4[3 * j + 1];
      float dx[PPT], dy[PPT], G[PPT], alpha[PPT];
      bool ok[PPT];
      bool any = false;
#pragma unroll
      for (int p = 0; p < PPT; p++) {
        dx[p] = a.x - fx[p];
        dy[p] = a.y - fy[p];
        const float power = a.z * dx[p] * dx[p] + b.x * dy[p] * dy[p] + a.w * dx[p] * dy[p];  // log2 domain (scaled conic)
        G[p] = ex2_approx_b(power);
        alpha[p] = fminf(0.99f, b.y * G[p]);
        ok[p] = (int)__float_as_uint(b.z) <= nc[p] && power <= 0.0f && alpha[p] >= 1.0f / 255.0f;
        any = any || ok[p];
      }
      if (__any_sync(0xffffffffu, any)) {
        const float2 c = *reinterpret_cast<const float2*>(&rec4[3 * j + 2]);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float v8 = 0.f;
#pragma unroll
        for (int p = 0; p < PPT; p++) {
          if (!ok[p]) continue;
          const float rcp = __fdividef(1.f, 1.f - alpha[p]);
          T[p] *= rcp;
          const float w = alpha[p] * T[p];
          a0[p] = last_alpha[p] * lc0[p] + (1.f - last_alpha[p]) * a0[p];
          a1[p] = last_alpha[p] * lc1[p] + (1.f - last_alpha[p]) * a1[p];
          a2[p] = last_alpha[p] * lc2[p] + (1.f - last_alpha[p]) * a2[p];
          lc0[p] = b.w; lc1[p] = c.x; lc2[p] = c.y;
          float dL_dalpha = (b.w - a0[p]) * g0[p] + (c.x - a1[p]) * g1[p] + (c.y - a2[p]) * g2[p];
          dL_dalpha *= T[p];
          last_alpha[p] = alpha[p];
          dL_dalpha += gob[p] * rcp;
          // raw moments of t = dL/dG * G about the Gaussian's mean; the per-record factors (conic, -1/2,
          // NDC scale, 1/opacity) are applied once per record in the flush instead of once per pixel
          const float t = b.y * dL_dalpha * G[p];
          const float tdx = t * dx[p], tdy = t * dy[p];
          v[0] += t;
          v[1] += tdx;
          v[2] += tdy;
          v[3] = fmaf(tdx, dx[p], v[3]);
          v[4] = fmaf(tdx, dy[p], v[4]);
          v[5] = fmaf(tdy, dy[p], v[5]);
          v[6] = fmaf(w, g0[p], v[6]);
          v[7] = fmaf(w, g1[p], v[7]);
          v8 = fmaf(w, g2[p], v8);
        }
        const float r8 = butterfly8(v, lane);
        const float r1 = warp_sum(v8);
        if ((lane & 3) == 0) s_acc[warp][j][lane >> 2] = r8;
        if (lane == 1) s_acc[warp][j][8] = r1;
        touched |= 1u << (j & 31);
      }
      if ((j & 31) == 0) {
        if (lane == 0) s_touched[warp][j >> 5] = touched;
        touched = 0;
      }
    }
    __syncthreads();  // partial sums of this batch are complete
    // flush: one thread per record adds the valid warp rows and issues the global reductions
    for (int j = tid; j < cnt; j += NT) {
      float q[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      bool hit = false;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        if ((s_touched[w][j >> 5] >> (j & 31)) & 1u) {
          hit = true;
          const float4 q0 = *reinterpret_cast<const float4*>(&s_acc[w][j][0]);
          const float4 q1 = *reinterpret_cast<const float4*>(&s_acc[w][j][4]);
          q[0] += q0.x; q[1] += q0.y; q[2] += q0.z; q[3] += q0.w;
          q[4] += q1.x; q[5] += q1.y; q[6] += q1.z; q[7] += q1.w;
          q[8] += s_acc[w][j][8];
        }
      }
      if (!hit) continue;
      const Record rec = s_rec[st][j];
      const uint32_t id = rec.id;
      // q = (m0, mx, my, mxx, mxy, myy): dmean2D = -(conic . m1) * NDC scale, dconic = -1/2 m2, dopacity = m0 / opacity
      float cA, cB, cC;
      record_conic(rec.qa, rec.qb, rec.qc, cA, cB, cC);
      const float gmx = -(cA * q[1] + cB * q[2]) * ddelx_dx;
      const float gmy = -(cC * q[2] + cB * q[1]) * ddely_dy;
      double* gg = ggrad + ((size_t)view * P + id) * OCRF_GGRAD_STRIDE;
      atomicAdd(gg + 0, (double)gmx);
      atomicAdd(gg + 1, (double)gmy);
      atomicAdd(gg + 2, (double)(-0.5f * q[3]));
      atomicAdd(gg + 3, (double)(-0.5f * q[4]));
      atomicAdd(gg + 4, (double)(-0.5f * q[5]));
      atomicAdd(gg + 5, (double)__fdividef(q[0], rec.op));
      float* gc = dL_dcolors + ((size_t)(colors_per_view ? view : view / views_per_sample) * P + id) * 3;
      atomicAdd(gc, q[6]);
      atomicAdd(gc + 1, q[7]);
      atomicAdd(gc + 2, q[8]);
    }
    __syncthreads();  // stage `st` and the partial sums are consumed: only now may round r+2 land
    if (tid == 0 && r + 2 < rounds) issue(r + 2);
  }
}

// The production C == 3 kernel.  Same traversal, staging and cross-warp reduction as the scalar kernel above; the
// per-(pixel, Gaussian) arithmetic is restructured around three facts:
//   * a thread owns PPT pixels of ONE column, so dx is common to them: of the six raw moments of t = dL/dG * G only
//     s0 = sum t, s1 = sum t dy, s2 = sum t dy^2 are accumulated per pixel; (m_x, m_xx, m_xy) = (dx s0, dx^2 s0, dx s1)
//     follow once per (thread, record);
//   * only g . A of the colour A accumulated behind a pixel is needed, and it obeys the same recurrence as A
//     (S <- S + alpha (g . c - S), applied eagerly): one register and one FMA per pixel instead of seven registers and
//     thirteen instructions (the reference's lazy last_alpha / last_color form, backward.cu:480-490);
//   * pixels that do not take part get alpha = 0 and G = 0, after which every formula is exact for them (1/(1-0) = 1,
//     all contributions 0), so the gradient arithmetic runs unpredicated and PACKED two pixels per instruction
//     (FFMA2 / FMUL2 / FADD2; record scalars are broadcast operands).
template <int PPT>
__global__ void __launch_bounds__(TILE_PIX / PPT) render_backward_c3_kernel(
    int W, int H, int P, int views_per_sample, int colors_per_view, const uint2* __restrict__ ranges /* culled lists */,
    const Record* __restrict__ records, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ max_contrib,
    const float* __restrict__ dL_dpix, const float* __restrict__ dL_dopa, double* __restrict__ ggrad,
    float* __restrict__ dL_dcolors) {
  pdl_enter();
  static_assert(PPT == 2 || PPT == 4, "pixel pairs");
  constexpr int NT = TILE_PIX / PPT;
  constexpr int NW = NT / 32;
  constexpr int NQ = PPT / 2;
  constexpr int BWD_BATCH = BwdBatch<PPT>::value;
  __shared__ __align__(128) Record s_rec[2][BWD_BATCH];
  __shared__ __align__(16) float s_acc[NW][BWD_BATCH][BWD_ACC];
  __shared__ uint32_t s_touched[NW][BWD_BATCH / 32];
  __shared__ __align__(8) uint64_t s_bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x, tiles_per_view = gridDim.x * gridDim.y;
  const int view = blockIdx.z;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const size_t vt = (size_t)view * tiles_per_view + tile;
  const int mc = (int)max_contrib[vt];
  if (mc == 0) return;
  const uint2 range = ranges[vt];
  const Record* src = records + range.x;
  const int rounds = (mc + BWD_BATCH - 1) / BWD_BATCH;

  const size_t HW = (size_t)H * W;
  const int px = blockIdx.x * TILE + (warp & 1) * 8 + (lane & 7);
  const int py0 = blockIdx.y * TILE + (warp >> 1) * (4 * PPT) + (lane >> 3);
  const float fx = (float)px;
  float2 nfy[NQ], T[NQ], S[NQ], g0[NQ], g1[NQ], g2[NQ], gob[NQ];
  int nc[PPT];
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
#pragma unroll
  for (int p = 0; p < PPT; p++) {
    const int py = py0 + 4 * p;
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)py * W + px;
    const float Tf = inside ? final_T[view * HW + pix] : 0.f;
    nc[p] = inside ? (int)n_contrib[view * HW + pix] : 0;
    const float c0 = inside ? dL_dpix[((size_t)view * 3 + 0) * HW + pix] : 0.f;
    const float c1 = inside ? dL_dpix[((size_t)view * 3 + 1) * HW + pix] : 0.f;
    const float c2 = inside ? dL_dpix[((size_t)view * 3 + 2) * HW + pix] : 0.f;
    const float gop = (inside && dL_dopa) ? dL_dopa[view * HW + pix] : 0.f;
    // d(out)/d(alpha_i) through the final transmittance: opacity map (+) and background (-)
    const float gb = Tf * (gop - (bg0 * c0 + bg1 * c1 + bg2 * c2));
    const int q = p >> 1;
    if (p & 1) { nfy[q].y = -(float)py; T[q].y = Tf; g0[q].y = c0; g1[q].y = c1; g2[q].y = c2; gob[q].y = gb; }
    else       { nfy[q].x = -(float)py; T[q].x = Tf; g0[q].x = c0; g1[q].x = c1; g2[q].x = c2; gob[q].x = gb; }
  }
#pragma unroll
  for (int q = 0; q < NQ; q++) S[q] = make_float2(0.f, 0.f);
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

  // round r covers culled-list indices [lo_r, hi_r) with hi_r = mc - r*BATCH (back to front)
  auto issue = [&](int r) {
    const int hi = mc - r * BWD_BATCH;
    const int lo = max(0, hi - BWD_BATCH);
    const uint32_t bytes = (uint32_t)(hi - lo) * sizeof(Record);
    mbar_expect_tx(&s_bar[r & 1], bytes);
    bulk_g2s(&s_rec[r & 1][0], src + lo, bytes, &s_bar[r & 1]);
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    issue(0);
    if (rounds > 1) issue(1);
  }

  for (int r = 0; r < rounds; r++) {
    const int st = r & 1;
    const int hi = mc - r * BWD_BATCH;
    const int lo = max(0, hi - BWD_BATCH);
    const int cnt = hi - lo;
    mbar_wait(&s_bar[st], (r >> 1) & 1);
    const float4* rec4 = reinterpret_cast<const float4*>(&s_rec[st][0]);
    uint32_t touched = 0;  // warp-uniform: bit (j & 31) set when this warp reduced record j

    for (int j = cnt - 1; j >= 0; j--) {
      // pixel p blended this record iff its reference list position (b.z) <= n_contrib[p] and the tests pass
      const float4 a = rec4[3 * j], b = rec4[3 * j + 1];  // x, y, qa, qb | qc, op, orig, r
      const float dx = a.x - fx;
      const float u = a.z * dx * dx;  // qa dx^2
      const float w = a.w * dx;       // qb dx
      const int orig = (int)__float_as_uint(b.z);
      float2 dy[NQ], araw[NQ];
      bool ok[PPT];
      bool any = false;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        dy[q] = fadd2(bcast2(a.y), nfy[q]);
        const float2 t = ffma2(bcast2(b.x), dy[q], bcast2(w));
        const float2 pw = ffma2(t, dy[q], bcast2(u));
        const float2 G = make_float2(ex2_approx_b(pw.x), ex2_approx_b(pw.y));
        araw[q] = fmul2(bcast2(b.y), G);  // opacity * G: alpha before the 0.99 clamp
        ok[2 * q] = orig <= nc[2 * q] && pw.x <= 0.0f && araw[q].x >= 1.0f / 255.0f;
        ok[2 * q + 1] = orig <= nc[2 * q + 1] && pw.y <= 0.0f && araw[q].y >= 1.0f / 255.0f;
        any = any || ok[2 * q] || ok[2 * q + 1];
      }
      if (__any_sync(0xffffffffu, any)) {
        const float2 c = *reinterpret_cast<const float2*>(&rec4[3 * j + 2]);  // g, b
        float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, k0 = s0, k1 = s0, k2 = s0;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          float2 og;  // opacity * G, zero for a pixel that does not take part
          og.x = ok[2 * q] ? araw[q].x : 0.f;
          og.y = ok[2 * q + 1] ? araw[q].y : 0.f;
          const float2 al = make_float2(fminf(0.99f, og.x), fminf(0.99f, og.y));
          const float2 om = fadd2(bcast2(1.f), make_float2(-al.x, -al.y));
          const float2 rcp = make_float2(rcp_approx(om.x), rcp_approx(om.y));
          T[q] = fmul2(T[q], rcp);
          const float2 wgt = fmul2(al, T[q]);
          float2 gc = fmul2(g0[q], bcast2(b.w));  // g . colour of this Gaussian
          gc = ffma2(g1[q], bcast2(c.x), gc);
          gc = ffma2(g2[q], bcast2(c.y), gc);
          const float2 dot = fadd2(gc, make_float2(-S[q].x, -S[q].y));
          const float2 dL_dalpha = ffma2(dot, T[q], fmul2(gob[q], rcp));
          S[q] = ffma2(al, dot, S[q]);
          const float2 t = fmul2(og, dL_dalpha);  // dL/dG * G
          const float2 tdy = fmul2(t, dy[q]);
          s0 = fadd2(s0, t);
          s1 = fadd2(s1, tdy);
          s2 = ffma2(tdy, dy[q], s2);
          k0 = ffma2(wgt, g0[q], k0);
          k1 = ffma2(wgt, g1[q], k1);
          k2 = ffma2(wgt, g2[q], k2);
        }
        const float m0 = s0.x + s0.y, my = s1.x + s1.y, myy = s2.x + s2.y;
        float v[8];
        v[0] = m0;
        v[1] = m0 * dx;       // sum t dx
        v[2] = my;            // sum t dy
        v[3] = v[1] * dx;     // sum t dx^2
        v[4] = my * dx;       // sum t dx dy
        v[5] = myy;           // sum t dy^2
        v[6] = k0.x + k0.y;
        v[7] = k1.x + k1.y;
        const float v8 = k2.x + k2.y;
        const float r8 = butterfly8(v, lane);
        const float r1 = warp_sum(v8);
        if ((lane & 3) == 0) s_acc[warp][j][lane >> 2] = r8;
        if (lane == 1) s_acc[warp][j][8] = r1;
        touched |= 1u << (j & 31);
      }
      if ((j & 31) == 0) {
        if (lane == 0) s_touched[warp][j >> 5] = touched;
        touched = 0;
      }
    }
    __syncthreads();  // partial sums of this batch are complete
    // flush: one thread per record adds the valid warp rows and issues the global reductions
    for (int j = tid; j < cnt; j += NT) {
      float q[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      bool hit = false;
#pragma unroll
      for (int w2 = 0; w2 < NW; w2++) {
        if ((s_touched[w2][j >> 5] >> (j & 31)) & 1u) {
          hit = true;
          const float4 q0 = *reinterpret_cast<const float4*>(&s_acc[w2][j][0]);
          const float4 q1 = *reinterpret_cast<const float4*>(&s_acc[w2][j][4]);
          q[0] += q0.x; q[1] += q0.y; q[2] += q0.z; q[3] += q0.w;
          q[4] += q1.x; q[5] += q1.y; q[6] += q1.z; q[7] += q1.w;
          q[8] += s_acc[w2][j][8];
        }
      }
      if (!hit) continue;
      const Record rec = s_rec[st][j];
      const uint32_t id = rec.id;
      float cA, cB, cC;
      record_conic(rec.qa, rec.qb, rec.qc, cA, cB, cC);
      // q = (m0, mx, my, mxx, mxy, myy): dmean2D = -(conic . m1) * NDC scale, dconic = -1/2 m2, dopacity = m0 / opacity
      const float gmx = -(cA * q[1] + cB * q[2]) * ddelx_dx;
      const float gmy = -(cC * q[2] + cB * q[1]) * ddely_dy;
      double* gg = ggrad + ((size_t)view * P + id) * OCRF_GGRAD_STRIDE;
      atomicAdd(gg + 0, (double)gmx);
      atomicAdd(gg + 1, (double)gmy);
      atomicAdd(gg + 2, (double)(-0.5f * q[3]));
      atomicAdd(gg + 3, (double)(-0.5f * q[4]));
      atomicAdd(gg + 4, (double)(-0.5f * q[5]));
      atomicAdd(gg + 5, (double)__fdividef(q[0], rec.op));
      float* gcol = dL_dcolors + ((size_t)(colors_per_view ? view : view / views_per_sample) * P + id) * 3;
      atomicAdd(gcol, q[6]);
      atomicAdd(gcol + 1, q[7]);
      atomicAdd(gcol + 2, q[8]);
    }
    __syncthreads();  // stage `st` and the partial sums are consumed: only now may round r+2 land
    if (tid == 0 && r + 2 < rounds) issue(r + 2);
  }
}

// Generic channel count (C != 3, e.g. the 80-channel feature rendering of GeoEnhDet).
//
// dL/dalpha of a (pixel, Gaussian) pair needs the sum over ALL channels of (c_k - A_k) g_k before the
// geometry gradients can be formed, so the list is walked once with the whole upstream gradient g[CP] of the pixel
// in REGISTERS (CP = C rounded up to 16, fully unrolled); the batch's feature rows sit in shared memory and are read
// as broadcast LDS.128.  The colour accumulated behind the pixel, A_k, is never formed: only g . A is needed, and
// that scalar follows the same recurrence, S <- S + alpha (g . c - S), applied eagerly right after it is used (the
// reference's lazy `last_alpha * last_color + (1 - last_alpha) * accum`, backward.cu:480-490, one step earlier) --
// one register and one FMA instead of CP of each per (pixel, Gaussian).  The feature gradients
// dF[record][channel] = sum_pixels w[pixel][record] g[pixel][channel] are a dense product: every warp keeps the blend
// weights of its 32 pixels for 16 records in shared memory and multiplies them with its pixels' gradients on the tensor
// cores (3xTF32 mma.sync, helpers below) instead of CP multiplies and 5 butterfly reductions per record; the eight
// partial tiles are summed and ONE global reduction per (record, channel) is issued (they were 8x as many when every
// warp issued its own).  Geometry gradients follow the C = 3 scheme.  Known limit: the resident g tile (90 KB at
// C = 80) allows one CTA per SM, so the 2.4x fewer instructions buy only 6 % of time -- see DESIGN.md section 8.
constexpr int BWDG_BATCH = 64;
constexpr int BWDG_FB = 16;  // records per cross-warp reduction of the feature gradients

// reduce-scatter of 16 values: afterwards lane L holds the warp total of value (L >> 1) & 15 (both lanes of a pair)
__device__ __forceinline__ float butterfly16(const float* v, int lane) {
  float w[8], u[4], t[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float keep = h16 ? v[i + 8] : v[i], send = h16 ? v[i] : v[i + 8];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float keep = h8 ? w[i + 4] : w[i], send = h8 ? w[i] : w[i + 4];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float keep = h4 ? u[i + 2] : u[i], send = h4 ? u[i] : u[i + 2];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h2 ? t[1] : t[0], send = h2 ? t[0] : t[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;  // value index = h16*8 + h8*4 + h4*2 + h2
}

constexpr int BWDG_WP = 36;  // pitch of a record's 32 pixel weights (== 4 mod 32: conflict-free A fragments)

template <int CP>
__global__ void __launch_bounds__(TILE_PIX, 1) render_backward_generic_kernel(
    int W, int H, int C, int P, int views_per_sample, const uint2* __restrict__ ranges /* culled lists */,
    const Record* __restrict__ records, const float* __restrict__ feats, const float* __restrict__ bg,
    const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
    const uint32_t* __restrict__ max_contrib, const float* __restrict__ dL_dpix, const float* __restrict__ dL_dopa,
    double* __restrict__ ggrad, float* __restrict__ dL_dfeats) {
  constexpr int NW = TILE_PIX / 32;
  extern __shared__ __align__(16) unsigned char smem_g[];
  Record* s_rec = reinterpret_cast<Record*>(smem_g);                                   // [BATCH]
  float* s_feat = reinterpret_cast<float*>(smem_g + BWDG_BATCH * sizeof(Record));      // [BATCH][CP]
  float* s_acc = s_feat + BWDG_BATCH * CP;                                             // [NW][BATCH][8]
  float* s_facc = s_acc + NW * BWDG_BATCH * 8;                                         // [NW][BWDG_FB][CP] per-warp partial dF
  constexpr int GP = CP + 8;                                                           // g row pitch (== 8 mod 16: conflict-free B fragments)
  float* s_g = s_facc + NW * BWDG_FB * CP;                                             // [NW][32][GP] upstream gradients of the tile
  float* s_w = s_g + NW * 32 * GP;                                                     // [NW][BWDG_FB][BWDG_WP] blend weights
  __shared__ uint32_t s_touched[NW][BWDG_BATCH / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = gridDim.x, tiles_per_view = gridDim.x * gridDim.y;
  const int view = blockIdx.z;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const size_t vt = (size_t)view * tiles_per_view + tile;
  const int mc = (int)max_contrib[vt];
  if (mc == 0) return;
  const uint2 range = ranges[vt];
  const int s = view / views_per_sample;
  const float* fbase = feats + (size_t)s * P * C;
  float* gfbase = dL_dfeats + (size_t)s * P * C;
  const size_t HW = (size_t)H * W;
  const int px = blockIdx.x * TILE + (warp & 1) * 8 + (lane & 7);
  const int py = blockIdx.y * TILE + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const size_t pix = (size_t)py * W + px;
  const float fx = (float)px, fy = (float)py;
  const float Tf = inside ? final_T[view * HW + pix] : 0.f;
  const int nc = inside ? (int)n_contrib[view * HW + pix] : 0;
  float g[CP];
  float bgdot = 0.f;
#pragma unroll
  for (int k = 0; k < CP; k++) {
    g[k] = (inside && k < C) ? dL_dpix[((size_t)view * C + k) * HW + pix] : 0.f;
    bgdot += (k < C ? bg[k] : 0.f) * g[k];
  }
  // g . (colour accumulated behind this pixel): the reference keeps the accumulated colour per channel
  // (backward.cu:480-490) and dots it with dL/dpixel; the dot product obeys the same linear recurrence, so one scalar
  // replaces C registers and C multiply-adds per (pixel, Gaussian)
  float S = 0.f;
  {
    float* grow = s_g + ((size_t)warp * 32 + lane) * GP;
#pragma unroll
    for (int k = 0; k < CP; k += 4) *reinterpret_cast<float4*>(grow + k) = make_float4(g[k], g[k + 1], g[k + 2], g[k + 3]);
  }
  float* my_w = s_w + (size_t)warp * BWDG_FB * BWDG_WP + lane;
  const float gop = (inside && dL_dopa) ? dL_dopa[view * HW + pix] : 0.f;
  const float gob = Tf * (gop - bgdot);
  float T = Tf;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  const int rounds = (mc + BWDG_BATCH - 1) / BWDG_BATCH;

  for (int r = 0; r < rounds; r++) {
    const int hi = mc - r * BWDG_BATCH;
    const int lo = max(0, hi - BWDG_BATCH);
    const int cnt = hi - lo;
    __syncthreads();
    if (tid < cnt) s_rec[tid] = records[range.x + lo + tid];
    for (int e = tid; e < cnt * (CP / 4); e += TILE_PIX) {  // feature rows, 16 bytes per thread
      const int j = e / (CP / 4), k4 = (e - j * (CP / 4)) * 4;
      const uint32_t id = records[range.x + lo + j].id;
      float4 f;
      f.x = k4 + 0 < C ? __ldg(fbase + (size_t)id * C + k4 + 0) : 0.f;
      f.y = k4 + 1 < C ? __ldg(fbase + (size_t)id * C + k4 + 1) : 0.f;
      f.z = k4 + 2 < C ? __ldg(fbase + (size_t)id * C + k4 + 2) : 0.f;
      f.w = k4 + 3 < C ? __ldg(fbase + (size_t)id * C + k4 + 3) : 0.f;
      *reinterpret_cast<float4*>(s_feat + j * CP + k4) = f;
    }
    __syncthreads();
    uint32_t touched = 0;
    for (int j = cnt - 1; j >= 0; j--) {
      const float4 a = reinterpret_cast<const float4*>(&s_rec[j])[0];
      const float4 b = reinterpret_cast<const float4*>(&s_rec[j])[1];
      const float dx = a.x - fx, dy = a.y - fy;
      const float power = a.z * dx * dx + b.x * dy * dy + a.w * dx * dy;  // log2 domain (scaled conic)
      const float G = ex2_approx_b(power);
      const float alpha = fminf(0.99f, b.y * G);
      const bool ok = (int)__float_as_uint(b.z) <= nc && power <= 0.0f && alpha >= 1.0f / 255.0f;
      float rcp = 1.f, w = 0.f, al = 0.f;
      if (ok) {
        rcp = __fdividef(1.f, 1.f - alpha);
        T *= rcp;
        w = alpha * T;
        al = alpha;
      }
      my_w[(j & (BWDG_FB - 1)) * BWDG_WP] = w;  // column `lane` of the warp's w^T tile (0 for a pixel that did not blend)
      if (__any_sync(0xffffffffu, ok)) {
        const float* fj = s_feat + j * CP;
        // g . colour of this Gaussian: two channels per FFMA2, two independent accumulator pairs
        float2 d01 = make_float2(0.f, 0.f), d23 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CP; k += 4) {
          const float4 c = *reinterpret_cast<const float4*>(fj + k);
          d01 = ffma2(make_float2(c.x, c.y), make_float2(g[k], g[k + 1]), d01);
          d23 = ffma2(make_float2(c.z, c.w), make_float2(g[k + 2], g[k + 3]), d23);
        }
        const float dsum = (d01.x + d01.y) + (d23.x + d23.y);
        const float dot = dsum - S;
        S = fmaf(al, dot, S);  // lanes that did not blend have al = 0
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (ok) {
          const float dL_dalpha = dot * T + gob * rcp;
          const float dL_dG = b.y * dL_dalpha;
          const float gdx = G * dx, gdy = G * dy;
          // conic recovered from the scaled record: A = -2 ln2 qa, B = -ln2 qb, C = -2 ln2 qc
          v[0] = dL_dG * (LN2 * (2.f * gdx * a.z + gdy * a.w)) * ddelx_dx;
          v[1] = dL_dG * (LN2 * (2.f * gdy * b.x + gdx * a.w)) * ddely_dy;
          v[2] = -0.5f * gdx * dx * dL_dG;
          v[3] = -0.5f * gdx * dy * dL_dG;
          v[4] = -0.5f * gdy * dy * dL_dG;
          v[5] = G * dL_dalpha;
        }
        const float r8 = butterfly8(v, lane);
        if ((lane & 3) == 0) s_acc[(warp * BWDG_BATCH + j) * 8 + (lane >> 2)] = r8;
        touched |= 1u << (j & 31);
      }
      if ((j & (BWDG_FB - 1)) == 0) {
        // ---- records [j, j + BWDG_FB) are complete in every warp: one reduction over the warps per channel ----
        if (lane == 0) s_touched[warp][j >> 5] = touched;  // (the word keeps accumulating until its 32-record boundary)
        __syncwarp();
        {  // this warp's partial dF[16 records][CP] = w^T[16 x 32 px] . g[32 px x CP] on the tensor cores
          const int gid = lane >> 2, tig = lane & 3;
          const float* wt = s_w + (size_t)warp * BWDG_FB * BWDG_WP;
          const float* gt = s_g + (size_t)warp * 32 * GP;
          float cacc[CP / 8][4];
#pragma unroll
          for (int nt = 0; nt < CP / 8; nt++) cacc[nt][0] = cacc[nt][1] = cacc[nt][2] = cacc[nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            uint32_t ahi[4], alo[4];
            split_tf32(wt[gid * BWDG_WP + 8 * ks + tig], ahi[0], alo[0]);
            split_tf32(wt[(gid + 8) * BWDG_WP + 8 * ks + tig], ahi[1], alo[1]);
            split_tf32(wt[gid * BWDG_WP + 8 * ks + tig + 4], ahi[2], alo[2]);
            split_tf32(wt[(gid + 8) * BWDG_WP + 8 * ks + tig + 4], ahi[3], alo[3]);
#pragma unroll
            for (int nt = 0; nt < CP / 8; nt++) {
              uint32_t bhi[2], blo[2];
              split_tf32(gt[(8 * ks + tig) * GP + 8 * nt + gid], bhi[0], blo[0]);
              split_tf32(gt[(8 * ks + tig + 4) * GP + 8 * nt + gid], bhi[1], blo[1]);
              mma_tf32_16x8x8(cacc[nt], alo, bhi);
              mma_tf32_16x8x8(cacc[nt], ahi, blo);
              mma_tf32_16x8x8(cacc[nt], ahi, bhi);
            }
          }
          float* fo = s_facc + (size_t)warp * BWDG_FB * CP;
#pragma unroll
          for (int nt = 0; nt < CP / 8; nt++) {
            *reinterpret_cast<float2*>(fo + gid * CP + 8 * nt + 2 * tig) = make_float2(cacc[nt][0], cacc[nt][1]);
            *reinterpret_cast<float2*>(fo + (gid + 8) * CP + 8 * nt + 2 * tig) = make_float2(cacc[nt][2], cacc[nt][3]);
          }
        }
        __syncthreads();
        for (int e = tid; e < BWDG_FB * CP; e += TILE_PIX) {
          const int jj = e / CP, ch = e - jj * CP;
          const int rj = j + jj;
          if (rj >= cnt || ch >= C) continue;  // (rows past the end of the batch hold stale weights)
          float tot = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < NW; w2++) tot += s_facc[((size_t)w2 * BWDG_FB + jj) * CP + ch];
          if (tot != 0.f) atomicAdd(gfbase + (size_t)s_rec[rj].id * C + ch, tot);
        }
        __syncthreads();
      }
      if ((j & 31) == 0) touched = 0;
    }
    __syncthreads();
    for (int j = tid; j < cnt; j += TILE_PIX) {
      float q[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      bool hit = false;
#pragma unroll
      for (int w2 = 0; w2 < NW; w2++)
        if ((s_touched[w2][j >> 5] >> (j & 31)) & 1u) {
          hit = true;
          const float4 q0 = *reinterpret_cast<const float4*>(&s_acc[(w2 * BWDG_BATCH + j) * 8]);
          const float2 q1 = *reinterpret_cast<const float2*>(&s_acc[(w2 * BWDG_BATCH + j) * 8 + 4]);
          q[0] += q0.x; q[1] += q0.y; q[2] += q0.z; q[3] += q0.w; q[4] += q1.x; q[5] += q1.y;
        }
      if (!hit) continue;
      double* gg = ggrad + ((size_t)view * P + s_rec[j].id) * OCRF_GGRAD_STRIDE;
#pragma unroll
      for (int e = 0; e < 6; e++) atomicAdd(gg + e, (double)q[e]);
    }
  }
}

template <int CP>
static int launch_backward_generic(cudaStream_t st, dim3 grid, const OcrfShape* sh, const uint2* ranges, const Record* rec,
                                   const float* colors, const float* bg, const float* fT, const uint32_t* nc,
                                   const uint32_t* mc, const float* dL_dcolor, const float* dL_dopa, double* ggrad,
                                   float* dL_dcolors) {
  const size_t dyn = BWDG_BATCH * sizeof(Record) + (size_t)BWDG_BATCH * CP * 4 + (size_t)(TILE_PIX / 32) * BWDG_BATCH * 8 * 4 +
                     (size_t)(TILE_PIX / 32) * BWDG_FB * CP * 4 + (size_t)(TILE_PIX / 32) * 32 * (CP + 8) * 4 +
                     (size_t)(TILE_PIX / 32) * BWDG_FB * BWDG_WP * 4;
  cudaError_t e = cudaFuncSetAttribute(render_backward_generic_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)dyn);
  if (e != cudaSuccess) return (int)e;
  render_backward_generic_kernel<CP><<<grid, TILE_PIX, dyn, st>>>(sh->W, sh->H, sh->C, sh->P, sh->views_per_sample, ranges,
                                                                   rec, colors, bg, fT, nc, mc, dL_dcolor, dL_dopa, ggrad,
                                                                   dL_dcolors);
  return 0;
}

// Zeroes what ocrf_render_backward accumulates into: every ggrad row of a visible (view, Gaussian) pair -- rows of
// invisible pairs are never read or written downstream, so the 48-byte rows of ~80 % of the pairs are skipped --
// and the whole colour-gradient buffer.  Replaces the two torch::zeros of rasterize_points.cu:151-159 on this path.
__global__ void __launch_bounds__(256) clear_gradients_kernel(size_t n_pairs, const int32_t* __restrict__ radii,
                                                              double* __restrict__ ggrad, size_t n_color_vec4,
                                                              size_t n_color, float* __restrict__ dL_dcolors) {
  pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t o = t0; o < n_pairs; o += stride) {
    if (__ldg(radii + o) > 0) {
      double2* row = reinterpret_cast<double2*>(ggrad + o * OCRF_GGRAD_STRIDE);
      row[0] = make_double2(0., 0.);
      row[1] = make_double2(0., 0.);
      row[2] = make_double2(0., 0.);
    }
  }
  float4* c4 = reinterpret_cast<float4*>(dL_dcolors);
  for (size_t i = t0; i < n_color_vec4; i += stride) c4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = n_color_vec4 * 4 + t0; i < n_color; i += stride) dL_dcolors[i] = 0.f;
}

}  // namespace ocrf

using namespace ocrf;

static int env_int_b(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

extern "C" int ocrf_render_backward(void* stream, const OcrfShape* sh, uint64_t pair_capacity, const float* colors,
                                    int use_sh, const float* bg, const void* geom_ws, const void* bin_ws,
                                    const void* image_ws, const float* dL_dcolor, const float* dL_dopacity_map,
                                    double* ggrad, float* dL_dcolors) {
  if (!sh || !bg || !bin_ws || !image_ws || !dL_dcolor || !ggrad || !dL_dcolors) return OCRF_EINVAL;
  (void)geom_ws;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OcrfBinLayout B;
  OcrfImageLayout I;
  int rc = ocrf_bin_layout(sh, pair_capacity, &B);
  if (rc) return rc;
  ocrf_image_layout(sh, &I);
  const dim3 grid(tiles_x(*sh), tiles_y(*sh), sh->V);
  const uint2* ranges = at<uint2>(image_ws, I.ranges_render);
  const float* fT = at<float>(image_ws, I.final_T);
  const uint32_t* nc = at<uint32_t>(image_ws, I.n_contrib);
  const uint32_t* mc = at<uint32_t>(image_ws, I.max_contrib);
  if (sh->C == 3) {
    static const int ppt = env_int_b("OCRF_BWD_PPT", 4);
    static const int packed = env_int_b("OCRF_BWD_PACKED", 1);  // 0: the scalar kernel (A/B measurements)
    const Record* rec = at<Record>(bin_ws, B.records);
#define OCRF_BWD_ARGS sh->W, sh->H, sh->P, sh->views_per_sample, use_sh, ranges, rec, bg, fT, nc, mc, dL_dcolor, \
                      dL_dopacity_map, ggrad, dL_dcolors
    if (packed && ppt == 2) OCRF_LAUNCH(render_backward_c3_kernel<2>, dim3(grid), dim3(TILE_PIX / 2), 0, st, OCRF_BWD_ARGS);
    else if (packed) OCRF_LAUNCH(render_backward_c3_kernel<4>, dim3(grid), dim3(TILE_PIX / 4), 0, st, OCRF_BWD_ARGS);
    else if (ppt == 4) OCRF_LAUNCH(render_backward_c3_scalar_kernel<4>, dim3(grid), dim3(TILE_PIX / 4), 0, st, OCRF_BWD_ARGS);
    else if (ppt == 2) OCRF_LAUNCH(render_backward_c3_scalar_kernel<2>, dim3(grid), dim3(TILE_PIX / 2), 0, st, OCRF_BWD_ARGS);
    else OCRF_LAUNCH(render_backward_c3_scalar_kernel<1>, dim3(grid), dim3(TILE_PIX), 0, st, OCRF_BWD_ARGS);
#undef OCRF_BWD_ARGS
  } else {
    if (!colors) return OCRF_EINVAL;
    const Record* rec = at<Record>(bin_ws, B.records);
    int rc2;
#define OCRF_BWDG(CPV)                                                                                              \
  rc2 = launch_backward_generic<CPV>(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, dL_dcolor, dL_dopacity_map, \
                                     ggrad, dL_dcolors)
    static const bool tc_bwd = !(getenv("OCRF_TC_BWD") != nullptr && atoi(getenv("OCRF_TC_BWD")) == 0);
    if (tc_bwd && tc::forward_tc_supported(sh->C))
      rc2 = tc::launch_backward_tc(st, grid, sh, ranges, rec, colors, bg, fT, nc, mc, dL_dcolor, dL_dopacity_map, ggrad, dL_dcolors);
    else if (sh->C <= 16) OCRF_BWDG(16);
    else if (sh->C <= 32) OCRF_BWDG(32);
    else if (sh->C <= 48) OCRF_BWDG(48);
    else if (sh->C <= 64) OCRF_BWDG(64);
    else if (sh->C <= 80) OCRF_BWDG(80);
    else if (sh->C <= 96) OCRF_BWDG(96);
    else return OCRF_ECAPACITY;  // more than 96 feature channels: not supported by the register-resident backward
#undef OCRF_BWDG
    if (rc2) return rc2;
  }
  OCRF_CHECK_LAST();
  return 0;
}

extern "C" int ocrf_clear_gradients(void* stream, const OcrfShape* sh, int use_sh, const int32_t* radii, double* ggrad,
                                    float* dL_dcolors) {
  if (!sh || !radii || !ggrad || !dL_dcolors) return OCRF_EINVAL;
  if ((reinterpret_cast<uintptr_t>(ggrad) & 15) || (reinterpret_cast<uintptr_t>(dL_dcolors) & 15)) return OCRF_EINVAL;
  const size_t n_pairs = (size_t)sh->V * sh->P;
  const size_t n_color = use_sh ? n_pairs * 3 : (size_t)sh->S * sh->P * sh->C;
  if (n_pairs == 0) return 0;
  const size_t want = (n_pairs + 255) / 256;
  const unsigned grid = (unsigned)(want < (size_t)num_sms() * 8 ? want : (size_t)num_sms() * 8);
  OCRF_LAUNCH(clear_gradients_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), n_pairs, radii, ggrad,
               n_color / 4, n_color, dL_dcolors);
  OCRF_CHECK_LAST();
  return 0;
}
