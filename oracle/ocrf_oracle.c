/*
 * ocrf_oracle.c -- CPU restatement of the OcRFDet Gaussian render path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (ocrfdet_b200/) never does.
 *
 * PARITY PIN: the reference repository has no tests and no golden vectors for this path
 * (SURVEY.md section 4).  The oracle is pinned against outputs of the reference's own CUDA
 * rasterizer (oracle/_ref/libinria_ref.so, built from the vendored sources by oracle/Makefile),
 * recorded on a B200 by tests/golden/make_golden.py and committed under tests/golden/.
 * The median-depth output follows the w-depth fork's README only (its source is absent from the
 * reference tree): for that one output parity is UNPINNED.
 *
 * Reference files restated here (paths relative to
 * /root/reference/mmdet3d/models/necks/MVSGaussian/lib/submodules/diff-gaussian-rasterization/):
 *   preprocess          cuda_rasterizer/forward.cu:74-152,155-256, auxiliary.h:41-77,139-164
 *   binning             cuda_rasterizer/rasterizer_impl.cu:35-50,70-138,277-318
 *   forward blend       cuda_rasterizer/forward.cu:261-374 (+ /root/reference/diff-gaussian-rasterization-w-depth/README.md:8-13)
 *   backward blend      cuda_rasterizer/backward.cu:399-557
 *   preprocess backward cuda_rasterizer/backward.cu:144-396
 *   SH colours          cuda_rasterizer/forward.cu:20-71, backward.cu:20-139
 *   opacity mask (HOA)  /root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:230-242,1197-1199
 *
 * Floating point: the forward preprocess uses explicit fmaf() in exactly the places where nvcc
 * contracts the reference's expressions (read from the SASS of the reference build: a three-term
 * dot a*b + c*d + e*f becomes fma(e,f, fma(a,b, c*d)); division, reciprocal and sqrt are IEEE),
 * so radii, tile rectangles and depth bits -- everything the sort keys depend on -- are bit-exact
 * against the CUDA build.  Compile with -ffp-contract=off so gcc adds no contraction of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16

static inline float fmul(float a, float b) { return a * b; }
static inline float fadd(float a, float b) { return a + b; }
/* a*b + c*d + e*f as the CUDA build rounds it */
static inline float dot3(float a, float b, float c, float d, float e, float f) {
  return fmaf(e, f, fmaf(a, b, c * d));
}
/* m[0]*x + m[4]*y + m[8]*z + m[12] with stride-4 column-major element r (auxiliary.h:58-77) */
static inline float xform_row(const float* m, int r, float x, float y, float z) {
  return fmaf(m[8 + r], z, fmaf(m[r], x, m[4 + r] * y)) + m[12 + r];
}

/* ------------------------------------------------------------------------------------------
 * 3D covariance from scale and (un-normalised) quaternion: forward.cu:118-152.
 * q = (r, x, y, z).  Sigma = (S R)^T (S R); out = [S00 S01 S02 S11 S12 S22].
 * ------------------------------------------------------------------------------------------ */
static void cov3d_from_scale_rot(const float* s3, float mod, const float* q, float* out) {
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
  /* rotation entries R[c][k]: column c, row k of the glm matrix */
  float R[3][3];
  R[0][0] = 1.f - 2.f * (yy + zz);
  R[0][1] = 2.f * fmaf(x, y, -rz);
  R[0][2] = 2.f * fmaf(r, y, xz);
  R[1][0] = 2.f * fmaf(x, y, rz);
  R[1][1] = 1.f - 2.f * fmaf(x, x, zz);
  R[1][2] = 2.f * fmaf(y, z, -rx);
  R[2][0] = 2.f * fmaf(-r, y, xz);
  R[2][1] = 2.f * fmaf(y, z, rx);
  R[2][2] = 1.f - 2.f * fmaf(x, x, yy);
  const float s[3] = {s3[0] * mod, s3[1] * mod, s3[2] * mod};
  float M[3][3]; /* M[c][k] = s[k] * R[c][k] */
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) M[c][k] = s[k] * R[c][k];
  /* Sigma[c][k] = M[k][0]*M[c][0] + M[k][1]*M[c][1] + M[k][2]*M[c][2] */
#define SIG(c, k) dot3(M[k][0], M[c][0], M[k][1], M[c][1], M[k][2], M[c][2])
  out[0] = SIG(0, 0);
  out[1] = SIG(0, 1);
  out[2] = SIG(0, 2);
  out[3] = SIG(1, 1);
  out[4] = SIG(1, 2);
  out[5] = SIG(2, 2);
#undef SIG
}

/* EWA projection of the 3D covariance: forward.cu:74-113.  Returns (a, b, c) with the 0.3 low-pass. */
static void cov2d_ewa(const float* t_in, float fx, float fy, float tanx, float tany, const float* c6,
                      const float* v, float* a, float* b, float* c) {
  const float tz = t_in[2];
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  const float txtz = t_in[0] / tz, tytz = t_in[1] / tz;
  const float cx = fminf(limx, fmaxf(-limx, txtz)), cy = fminf(limy, fmaxf(-limy, tytz));
  const float J00 = fx / tz, J11 = fy / tz;
  const float tz2 = tz * tz;
  const float J02 = ((-(tz * cx)) * fx) / tz2;
  const float J12 = ((-(tz * cy)) * fy) / tz2;
  /* T[i][k], i in {0,1}: row i of J times the view rotation; v is column-major 4x4 */
  float T0[3], T1[3];
  for (int k = 0; k < 3; k++) {
    T0[k] = fmaf(v[4 * k + 2], J02, fmaf(v[4 * k + 0], J00, 0.f * v[4 * k + 1]));
    T1[k] = fmaf(v[4 * k + 2], J12, fmaf(0.f, v[4 * k + 0], v[4 * k + 1] * J11));
  }
  const float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
  float A0[3], A1[3]; /* A_i[c] = T_i . S[:,c] */
  for (int k = 0; k < 3; k++) {
    A0[k] = dot3(T0[0], S[0][k], T0[1], S[1][k], T0[2], S[2][k]);
    A1[k] = dot3(T1[0], S[0][k], T1[1], S[1][k], T1[2], S[2][k]);
  }
  *a = dot3(T0[0], A0[0], T0[1], A0[1], T0[2], A0[2]) + 0.3f;
  *b = dot3(T0[0], A1[0], T0[1], A1[1], T0[2], A1[2]);
  *c = dot3(T1[0], A1[0], T1[1], A1[1], T1[2], A1[2]) + 0.3f;
}

/* auxiliary.h:41-44 -- evaluated in double, one fused multiply-add, stored as float */
static inline float ndc_to_pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* float -> int conversion with CUDA semantics (saturating, NaN -> 0) */
static inline int f2i_trunc(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.f) return 2147483647;
  if (f <= -2147483648.f) return (int)0x80000000;
  return (int)f;
}

/* auxiliary.h:46-56 */
static void tile_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
  const float r = (float)radius;
  *x0 = clampi(f2i_trunc((px - r) * 0.0625f), 0, gx);
  *y0 = clampi(f2i_trunc((py - r) * 0.0625f), 0, gy);
  *x1 = clampi(f2i_trunc((((px + r) + 16.f) - 1.f) * 0.0625f), 0, gx);
  *y1 = clampi(f2i_trunc((((py + r) + 16.f) - 1.f) * 0.0625f), 0, gy);
}

/* SH basis constants: auxiliary.h:20-38 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/* number of basis functions of degree <= deg */
static inline int sh_count(int deg) { return (deg + 1) * (deg + 1); }

/* basis values b[0..n) for unit direction (x,y,z): forward.cu:31-60 */
static void sh_basis(int deg, float x, float y, float z, float* b) {
  b[0] = SH_C0;
  if (deg < 1) return;
  b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  b[4] = SH_C2[0] * xy; b[5] = SH_C2[1] * yz; b[6] = SH_C2[2] * (2.f * zz - xx - yy);
  b[7] = SH_C2[3] * xz; b[8] = SH_C2[4] * (xx - yy);
  if (deg < 3) return;
  b[9] = SH_C3[0] * y * (3.f * xx - yy); b[10] = SH_C3[1] * xy * z; b[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
  b[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); b[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
  b[14] = SH_C3[5] * z * (xx - yy); b[15] = SH_C3[6] * x * (xx - 3.f * yy);
}

/* d(basis)/d(x,y,z): backward.cu:58-128 restated per basis function */
static void sh_basis_grad(int deg, float x, float y, float z, float (*g)[3]) {
  for (int i = 0; i < 16; i++) g[i][0] = g[i][1] = g[i][2] = 0.f;
  if (deg < 1) return;
  g[1][1] = -SH_C1; g[2][2] = SH_C1; g[3][0] = -SH_C1;
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  g[4][0] = SH_C2[0] * y; g[4][1] = SH_C2[0] * x;
  g[5][1] = SH_C2[1] * z; g[5][2] = SH_C2[1] * y;
  g[6][0] = SH_C2[2] * 2.f * -x; g[6][1] = SH_C2[2] * 2.f * -y; g[6][2] = SH_C2[2] * 2.f * 2.f * z;
  g[7][0] = SH_C2[3] * z; g[7][2] = SH_C2[3] * x;
  g[8][0] = SH_C2[4] * 2.f * x; g[8][1] = SH_C2[4] * 2.f * -y;
  if (deg < 3) return;
  g[9][0] = SH_C3[0] * 3.f * 2.f * xy; g[9][1] = SH_C3[0] * 3.f * (xx - yy);
  g[10][0] = SH_C3[1] * yz; g[10][1] = SH_C3[1] * xz; g[10][2] = SH_C3[1] * xy;
  g[11][0] = SH_C3[2] * -2.f * xy; g[11][1] = SH_C3[2] * (-3.f * yy + 4.f * zz - xx); g[11][2] = SH_C3[2] * 4.f * 2.f * yz;
  g[12][0] = SH_C3[3] * -3.f * 2.f * xz; g[12][1] = SH_C3[3] * -3.f * 2.f * yz; g[12][2] = SH_C3[3] * 3.f * (2.f * zz - xx - yy);
  g[13][0] = SH_C3[4] * (-3.f * xx + 4.f * zz - yy); g[13][1] = SH_C3[4] * -2.f * xy; g[13][2] = SH_C3[4] * 4.f * 2.f * xz;
  g[14][0] = SH_C3[5] * 2.f * xz; g[14][1] = SH_C3[5] * -2.f * yz; g[14][2] = SH_C3[5] * (xx - yy);
  g[15][0] = SH_C3[6] * 3.f * (xx - yy); g[15][1] = SH_C3[6] * -3.f * 2.f * xy;
}

/* ------------------------------------------------------------------------------------------
 * Stage 1: per-Gaussian preprocess.  forward.cu:155-256.
 * Optional inputs may be NULL: cov3D_precomp (then scales/rots are used), shs (then rgb untouched).
 * Outputs for culled Gaussians: radii = 0, tiles_touched = 0, the rest untouched.
 * ------------------------------------------------------------------------------------------ */
void ocrf_oracle_preprocess(int P, int sh_deg, int sh_M, const float* means, const float* scales, float scale_modifier,
                            const float* rots, const float* opacities, const float* shs, const float* cov3D_precomp,
                            const float* view, const float* proj, const float* campos, int W, int H, float tanfovx,
                            float tanfovy, int32_t* radii, float* xy, float* depths, float* cov3D, float* conic_opacity,
                            uint32_t* tiles_touched, float* rgb, uint8_t* clamped) {
  const float fy = H / (2.0f * tanfovy), fx = W / (2.0f * tanfovx); /* rasterizer_impl.cu:222-223 */
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  for (int i = 0; i < P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    const float x = means[3 * i], y = means[3 * i + 1], z = means[3 * i + 2];
    float t[3];
    for (int r = 0; r < 3; r++) t[r] = xform_row(view, r, x, y, z);
    if (t[2] <= 0.2f) continue; /* auxiliary.h:154 */
    const float hx = xform_row(proj, 0, x, y, z), hy = xform_row(proj, 1, x, y, z), hw = xform_row(proj, 3, x, y, z);
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ndcx = hx * pw, ndcy = hy * pw;
    const float* c6;
    if (cov3D_precomp) {
      c6 = cov3D_precomp + 6 * i;
    } else {
      cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rots + 4 * i, cov3D + 6 * i);
      c6 = cov3D + 6 * i;
    }
    float a, b, c;
    cov2d_ewa(t, fx, fy, tanfovx, tanfovy, c6, view, &a, &b, &c);
    const float det = fmaf(a, c, -(b * b));
    if (det == 0.0f) continue;
    const float det_inv = 1.f / det;
    const float mid = 0.5f * (a + c);
    const float disc = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
    const float lam = fmaxf(mid + disc, mid - disc);
    const float rad_f = ceilf(3.f * sqrtf(lam));
    const int rad = f2i_trunc(rad_f);
    const float px = ndc_to_pix(ndcx, W), py = ndc_to_pix(ndcy, H);
    int x0, y0, x1, y1;
    tile_rect(px, py, rad, gx, gy, &x0, &y0, &x1, &y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;
    if (shs) { /* forward.cu:20-71 */
      float d[3] = {x - campos[0], y - campos[1], z - campos[2]};
      const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      d[0] /= len; d[1] /= len; d[2] /= len;
      float bas[16];
      sh_basis(sh_deg, d[0], d[1], d[2], bas);
      const int n = sh_count(sh_deg);
      for (int ch = 0; ch < 3; ch++) {
        float acc = 0.f;
        for (int k = 0; k < n; k++) acc += bas[k] * shs[((size_t)i * sh_M + k) * 3 + ch];
        acc += 0.5f;
        clamped[3 * i + ch] = acc < 0.f;
        rgb[3 * i + ch] = acc < 0.f ? 0.f : acc;
      }
    }
    depths[i] = t[2];
    radii[i] = rad;
    xy[2 * i] = px;
    xy[2 * i + 1] = py;
    conic_opacity[4 * i + 0] = c * det_inv;
    conic_opacity[4 * i + 1] = b * -det_inv;
    conic_opacity[4 * i + 2] = a * det_inv;
    conic_opacity[4 * i + 3] = opacities[i];
    tiles_touched[i] = (uint32_t)((y1 - y0) * (x1 - x0));
  }
}

/* rasterizer_impl.cu:54-66 (markVisible) */
void ocrf_oracle_mark_visible(int P, const float* means, const float* view, const float* proj, uint8_t* present) {
  (void)proj;
  for (int i = 0; i < P; i++)
    present[i] = xform_row(view, 2, means[3 * i], means[3 * i + 1], means[3 * i + 2]) > 0.2f;
}

/* rasterizer_impl.cu:35-50: position of the highest set bit, plus one (0 -> 0) */
uint32_t ocrf_oracle_higher_msb(uint32_t n) {
  uint32_t b = 0;
  while (b < 32 && (n >> b)) b++;
  return b == 0 ? 1 : b; /* the reference's binary search never returns 0: n=0 gives 1 */
}

/* ------------------------------------------------------------------------------------------
 * Stage 2: binning.  rasterizer_impl.cu:70-138, 277-318.
 * offsets = inclusive scan of tiles_touched.  Returns N = offsets[P-1].
 * Call with keys == NULL to obtain N only (after filling offsets).
 * ------------------------------------------------------------------------------------------ */
uint64_t ocrf_oracle_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets) {
  uint32_t acc = 0;
  for (int i = 0; i < P; i++) { acc += tiles_touched[i]; offsets[i] = acc; }
  return P ? acc : 0;
}

void ocrf_oracle_duplicate(int P, const float* xy, const float* depths, const int32_t* radii, const uint32_t* offsets,
                           int W, int H, uint64_t* keys, uint32_t* values) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  for (int i = 0; i < P; i++) {
    if (radii[i] <= 0) continue;
    uint32_t off = i ? offsets[i - 1] : 0;
    int x0, y0, x1, y1;
    tile_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
    uint32_t dbits;
    memcpy(&dbits, depths + i, 4);
    for (int ty = y0; ty < y1; ty++)
      for (int tx = x0; tx < x1; tx++) {
        keys[off] = ((uint64_t)(uint32_t)(ty * gx + tx) << 32) | dbits;
        values[off] = (uint32_t)i;
        off++;
      }
  }
}

/* Stable LSD radix sort on bits [0, end_bit) -- same contract as cub::DeviceRadixSort::SortPairs
 * (rasterizer_impl.cu:303-308).  Bits at or above end_bit do not take part in the ordering. */
void ocrf_oracle_sort_pairs(uint64_t n, const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                            uint32_t* vals_out, int end_bit) {
  if (n == 0) return;
  uint64_t* ka = (uint64_t*)malloc(n * 8);
  uint64_t* kb = (uint64_t*)malloc(n * 8);
  uint32_t* va = (uint32_t*)malloc(n * 4);
  uint32_t* vb = (uint32_t*)malloc(n * 4);
  memcpy(ka, keys_in, n * 8);
  memcpy(va, vals_in, n * 4);
  for (int shift = 0; shift < end_bit; shift += 8) {
    const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
    const uint64_t mask = (1ull << bits) - 1;
    uint64_t cnt[257] = {0};
    for (uint64_t i = 0; i < n; i++) cnt[((ka[i] >> shift) & mask) + 1]++;
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (uint64_t i = 0; i < n; i++) {
      const uint64_t d = (ka[i] >> shift) & mask;
      kb[cnt[d]] = ka[i];
      vb[cnt[d]] = va[i];
      cnt[d]++;
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  memcpy(keys_out, ka, n * 8);
  memcpy(vals_out, va, n * 4);
  free(ka); free(kb); free(va); free(vb);
}

/* rasterizer_impl.cu:116-138 + the memset at :310.  ranges is [tiles][2]. */
void ocrf_oracle_tile_ranges(uint64_t n, const uint64_t* keys_sorted, int tiles, uint32_t* ranges) {
  memset(ranges, 0, (size_t)tiles * 8);
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
    if (i == 0) ranges[2 * cur] = 0;
    else {
      const uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
    }
    if (i == n - 1) ranges[2 * cur + 1] = (uint32_t)n;
  }
}

/* ------------------------------------------------------------------------------------------
 * Stage 3: forward blend.  forward.cu:261-374, median depth per the w-depth README (default 15),
 * opacity = 1 - final_T (north-star extension).
 * `ambiguous` (optional) marks pixels where a discrete decision of the blend (the power>0 skip,
 * the alpha<1/255 skip, the T<1e-4 stop, the 0.5 median crossing) came within `amb_eps` (relative)
 * of its threshold: a second implementation whose exp() differs in the last bits may legitimately
 * take the other branch there.
 * ------------------------------------------------------------------------------------------ */
void ocrf_oracle_render_forward(int W, int H, int C, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                                const float* depths, const float* conic_opacity, const float* colors, const float* bg,
                                float* out_color, float* out_depth, float* out_opacity, float* final_T,
                                uint32_t* n_contrib, uint8_t* ambiguous, float amb_eps) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < gx * gy; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    float* acc = (float*)malloc(sizeof(float) * (size_t)(C > 0 ? C : 1));
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const float pxf = (float)px, pyf = (float)py;
        float T = 1.f, D = 15.f;
        uint32_t contributor = 0, last = 0;
        uint8_t amb = 0;
        for (int ch = 0; ch < C; ch++) acc[ch] = 0.f;
        for (uint32_t k = r0; k < r1; k++) {
          contributor++;
          const uint32_t g = point_list[k];
          const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
          const float cA = conic_opacity[4 * g], cB = conic_opacity[4 * g + 1], cC = conic_opacity[4 * g + 2],
                      op = conic_opacity[4 * g + 3];
          const float power = -0.5f * (cA * dx * dx + cC * dy * dy) - cB * dx * dy;
          if (fabsf(power) <= amb_eps) amb = 1;
          if (power > 0.f) continue;
          const float alpha = fminf(0.99f, op * expf(power));
          if (fabsf(alpha - 1.0f / 255.0f) <= amb_eps * (1.0f / 255.0f)) amb = 1;
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = T * (1 - alpha);
          if (fabsf(test_T - 0.0001f) <= amb_eps * 0.0001f) amb = 1;
          if (test_T < 0.0001f) break; /* pixel done; this Gaussian is not blended */
          for (int ch = 0; ch < C; ch++) acc[ch] += colors[(size_t)g * C + ch] * alpha * T;
          if (fabsf(T - 0.5f) <= amb_eps || fabsf(test_T - 0.5f) <= amb_eps) amb = 1;
          if (T > 0.5f && test_T < 0.5f) D = depths[g];
          T = test_T;
          last = contributor;
        }
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        for (int ch = 0; ch < C; ch++) out_color[(size_t)ch * H * W + pix] = acc[ch] + T * bg[ch];
        if (out_depth) out_depth[pix] = D;
        if (out_opacity) out_opacity[pix] = 1.f - T;
        if (ambiguous) ambiguous[pix] = amb;
      }
    free(acc);
  }
}

/* ------------------------------------------------------------------------------------------
 * Stage 4: backward blend.  backward.cu:399-557.  Per-term arithmetic in float as the reference;
 * the sums over pixels (global atomics in the reference, order undefined) are accumulated in double.
 * dL_dout_opacity (optional) is the gradient of the opacity-map extension: d(1-T_final)/d(alpha_i)
 * = T_final / (1 - alpha_i).  Depth has no backward (w-depth README:13).
 * Outputs are [P,2] mean2D (NDC-scaled: * 0.5 W, * 0.5 H), [P,3] conic (A, B, C), [P] opacity, [P,C] colours.
 * ------------------------------------------------------------------------------------------ */
void ocrf_oracle_render_backward(int P, int W, int H, int C, const uint32_t* ranges, const uint32_t* point_list,
                                 const float* xy, const float* conic_opacity, const float* colors, const float* bg,
                                 const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                                 const float* dL_dout_opacity, double* dL_dmean2D, double* dL_dconic,
                                 double* dL_dopacity, double* dL_dcolors) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  memset(dL_dmean2D, 0, sizeof(double) * 2 * (size_t)P);
  memset(dL_dconic, 0, sizeof(double) * 3 * (size_t)P);
  memset(dL_dopacity, 0, sizeof(double) * (size_t)P);
  memset(dL_dcolors, 0, sizeof(double) * (size_t)C * P);
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  float* accum = (float*)malloc(sizeof(float) * (size_t)(3 * C + 3));
  float* lastc = accum + C;
  float* gpix = lastc + C;
  for (int tile = 0; tile < gx * gy; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const size_t pix = (size_t)py * W + px;
        const float pxf = (float)px, pyf = (float)py;
        const float T_final = final_T[pix];
        float T = T_final;
        float last_alpha = 0.f;
        float bg_dot = 0.f;
        for (int ch = 0; ch < C; ch++) {
          accum[ch] = 0.f; lastc[ch] = 0.f;
          gpix[ch] = dL_dpix[(size_t)ch * H * W + pix];
          bg_dot += bg[ch] * gpix[ch];
        }
        const float gop = dL_dout_opacity ? dL_dout_opacity[pix] : 0.f;
        for (int64_t k = (int64_t)n_contrib[pix] - 1; k >= 0; k--) {
          const uint32_t g = point_list[r0 + k];
          const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
          const float cA = conic_opacity[4 * g], cB = conic_opacity[4 * g + 1], cC = conic_opacity[4 * g + 2],
                      op = conic_opacity[4 * g + 3];
          const float power = -0.5f * (cA * dx * dx + cC * dy * dy) - cB * dx * dy;
          if (power > 0.f) continue;
          const float G = expf(power);
          const float alpha = fminf(0.99f, op * G);
          if (alpha < 1.0f / 255.0f) continue;
          T = T / (1.f - alpha);
          const float w = alpha * T;
          float dL_dalpha = 0.f;
          for (int ch = 0; ch < C; ch++) {
            const float c = colors[(size_t)g * C + ch];
            accum[ch] = last_alpha * lastc[ch] + (1.f - last_alpha) * accum[ch];
            lastc[ch] = c;
            dL_dalpha += (c - accum[ch]) * gpix[ch];
            dL_dcolors[(size_t)g * C + ch] += (double)(w * gpix[ch]);
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
          dL_dalpha += (T_final / (1.f - alpha)) * gop;
          const float dL_dG = op * dL_dalpha;
          const float gdx = G * dx, gdy = G * dy;
          const float dG_ddelx = -gdx * cA - gdy * cB;
          const float dG_ddely = -gdy * cC - gdx * cB;
          dL_dmean2D[2 * g] += (double)(dL_dG * dG_ddelx * ddelx_dx);
          dL_dmean2D[2 * g + 1] += (double)(dL_dG * dG_ddely * ddely_dy);
          dL_dconic[3 * g] += (double)(-0.5f * gdx * dx * dL_dG);
          dL_dconic[3 * g + 1] += (double)(-0.5f * gdx * dy * dL_dG);
          dL_dconic[3 * g + 2] += (double)(-0.5f * gdy * dy * dL_dG);
          dL_dopacity[g] += (double)(G * dL_dalpha);
        }
      }
  }
  free(accum);
}

/* ------------------------------------------------------------------------------------------
 * Preprocess backward.  backward.cu:144-274 (conic -> cov2D -> cov3D and mean), :346-396 (projection
 * part of the mean gradient, SH), :278-341 (cov3D -> scale, quaternion; no normalisation Jacobian).
 * Inputs are float gradients [P,2] mean2D, [P,3] conic (A,B,C), [P,C] colours (only read for SH).
 * Any of dL_dscales/dL_drots (when cov3D_precomp is given) or dL_dcov3D/dL_dshs may be NULL.
 * Gaussians with radii <= 0 receive zeros.
 * ------------------------------------------------------------------------------------------ */
#define REAL float
#define RMIN fminf
#define RMAX fmaxf
#define RSQRT sqrtf
#define OCRF_PB_NAME ocrf_oracle_preprocess_backward
#include "preprocess_backward.inc"
#undef REAL
#undef RMIN
#undef RMAX
#undef RSQRT
#undef OCRF_PB_NAME
/* Same chain evaluated in double from double screen-space gradients: the conic -> covariance ->
 * quaternion chain cancels heavily in float32 (e.g. (denom - a*c) is -b*b), so two float32 builds of
 * the reference formulas differ from each other by ~1e-4 of the largest gradient. */
#define REAL double
#define RMIN fmin
#define RMAX fmax
#define RSQRT sqrt
#define OCRF_PB_NAME ocrf_oracle_preprocess_backward_f64
#include "preprocess_backward.inc"
#undef REAL
#undef RMIN
#undef RMAX
#undef RSQRT
#undef OCRF_PB_NAME

/* ------------------------------------------------------------------------------------------
 * Stage 5: opacity mask of the HOA lift (view_transformer_ocrf.py:230-242 `ObatinOpacityMask`,
 * applied at :1197-1199): mask = sigmoid(conv7x7([mean_c(x), max_c(x)]) + opacity_bev); out = x * mask.
 * x [B,C,H,W], w [2,K,K] (no bias, zero padding K/2), opacity_bev [B,1,H,W].
 * ------------------------------------------------------------------------------------------ */
void ocrf_oracle_opacity_mask_forward(int B, int C, int H, int W, int K, const float* x, const float* w,
                                      const float* opacity_bev, float* out, float* mask, float* stats) {
  const int pad = K / 2;
  const size_t HW = (size_t)H * W;
  for (int b = 0; b < B; b++) {
    float* st = stats + (size_t)b * 2 * HW; /* [2,H,W] mean, max */
    for (size_t p = 0; p < HW; p++) {
      float sum = 0.f, mx = -INFINITY;
      for (int c = 0; c < C; c++) {
        const float v = x[((size_t)b * C + c) * HW + p];
        sum += v;
        mx = fmaxf(mx, v);
      }
      st[p] = sum / (float)C;
      st[HW + p] = mx;
    }
    for (int yy = 0; yy < H; yy++)
      for (int xx = 0; xx < W; xx++) {
        float acc = 0.f;
        for (int ch = 0; ch < 2; ch++)
          for (int ky = 0; ky < K; ky++)
            for (int kx = 0; kx < K; kx++) {
              const int sy = yy + ky - pad, sx = xx + kx - pad;
              if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
              acc += w[(ch * K + ky) * K + kx] * st[ch * HW + (size_t)sy * W + sx];
            }
        const size_t p = (size_t)yy * W + xx;
        const float m = 1.f / (1.f + expf(-(acc + opacity_bev[b * HW + p])));
        mask[b * HW + p] = m;
        for (int c = 0; c < C; c++) out[((size_t)b * C + c) * HW + p] = x[((size_t)b * C + c) * HW + p] * m;
      }
  }
}

/* Backward of the above.  Ties in max_c send the gradient to the first maximal channel (torch.max). */
void ocrf_oracle_opacity_mask_backward(int B, int C, int H, int W, int K, const float* x, const float* w,
                                       const float* mask, const float* stats, const float* g_out, float* g_x,
                                       float* g_w, float* g_opacity_bev) {
  const int pad = K / 2;
  const size_t HW = (size_t)H * W;
  double* gw = (double*)calloc((size_t)2 * K * K, sizeof(double));
  float* gz = (float*)malloc(sizeof(float) * HW);
  float* gst = (float*)malloc(sizeof(float) * 2 * HW);
  for (int b = 0; b < B; b++) {
    const float* st = stats + (size_t)b * 2 * HW;
    for (size_t p = 0; p < HW; p++) {
      float gm = 0.f;
      const float m = mask[b * HW + p];
      for (int c = 0; c < C; c++) {
        const size_t idx = ((size_t)b * C + c) * HW + p;
        gm += g_out[idx] * x[idx];
        g_x[idx] = g_out[idx] * m;
      }
      gz[p] = gm * m * (1.f - m);
      g_opacity_bev[b * HW + p] = gz[p];
    }
    memset(gst, 0, sizeof(float) * 2 * HW);
    for (int yy = 0; yy < H; yy++)
      for (int xx = 0; xx < W; xx++) {
        const float g = gz[(size_t)yy * W + xx];
        for (int ch = 0; ch < 2; ch++)
          for (int ky = 0; ky < K; ky++)
            for (int kx = 0; kx < K; kx++) {
              const int sy = yy + ky - pad, sx = xx + kx - pad;
              if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
              gw[(ch * K + ky) * K + kx] += (double)(g * st[ch * HW + (size_t)sy * W + sx]);
              gst[ch * HW + (size_t)sy * W + sx] += g * w[(ch * K + ky) * K + kx];
            }
      }
    for (size_t p = 0; p < HW; p++) {
      const float gmean = gst[p] / (float)C;
      int arg = 0;
      float mx = -INFINITY;
      for (int c = 0; c < C; c++) {
        const float v = x[((size_t)b * C + c) * HW + p];
        if (v > mx) { mx = v; arg = c; }
      }
      for (int c = 0; c < C; c++) g_x[((size_t)b * C + c) * HW + p] += gmean + (c == arg ? gst[HW + p] : 0.f);
    }
  }
  for (int i = 0; i < 2 * K * K; i++) g_w[i] = (float)gw[i];
  free(gw); free(gz); free(gst);
}

/* ------------------------------------------------------------------------------------------
 * Stage 0: OcRF Gaussian construction heads (view_transformer_ocrf.py:272-320; built with
 * input_dim 80, hidden_dim 4 at :611-622; evaluated at :1130-1133).  Parameters are given per head
 * in torch's nn.Linear layout (weight [out,in], bias [out]), NOT packed:
 *   head order S (softplus, 3 outputs), R (normalize, 4), A (sigmoid, 1), C (sigmoid, 3; input = cat(feat, rgb)).
 * fc1_w[h] [4, in_h], fc1_b[h] [4], fc2_w[h] [out_h, 4], fc2_b[h] [out_h];  in_h = F (S,R,A) or F+3 (C).
 * ------------------------------------------------------------------------------------------ */
static const int GH_OUTS[4] = {3, 4, 1, 3};

static float gh_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); } /* nn.Softplus(beta=1, threshold=20) */
static float gh_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

void ocrf_oracle_gaussian_heads_forward(long long n, int F, const float* feat, const float* rgb,
                                        const float* const* fc1_w, const float* const* fc1_b,
                                        const float* const* fc2_w, const float* const* fc2_b, float* opacity,
                                        float* scales, float* rotations, float* colors) {
  for (long long g = 0; g < n; g++) {
    float z[4][4];
    for (int hd = 0; hd < 4; hd++) {
      const int in = hd == 3 ? F + 3 : F;
      float hdn[4];
      for (int j = 0; j < 4; j++) {
        float acc = fc1_b[hd][j];
        for (int k = 0; k < F; k++) acc += feat[g * F + k] * fc1_w[hd][j * in + k];
        if (hd == 3)
          for (int c = 0; c < 3; c++) acc += rgb[g * 3 + c] * fc1_w[hd][j * in + F + c];
        hdn[j] = acc > 0.f ? acc : 0.f;
      }
      for (int i = 0; i < GH_OUTS[hd]; i++) {
        float acc = fc2_b[hd][i];
        for (int j = 0; j < 4; j++) acc += fc2_w[hd][i * 4 + j] * hdn[j];
        z[hd][i] = acc;
      }
    }
    for (int c = 0; c < 3; c++) scales[g * 3 + c] = gh_softplus(z[0][c]);
    float nrm = sqrtf(z[1][0] * z[1][0] + z[1][1] * z[1][1] + z[1][2] * z[1][2] + z[1][3] * z[1][3]);
    if (nrm < 1e-12f) nrm = 1e-12f; /* F.normalize eps */
    for (int c = 0; c < 4; c++) rotations[g * 4 + c] = z[1][c] / nrm;
    opacity[g] = gh_sigmoid(z[2][0]);
    for (int c = 0; c < 3; c++) colors[g * 3 + c] = gh_sigmoid(z[3][c]);
  }
}

/* Backward in double accumulation for the parameter sums (order-independent reference values).
 * g_fc1_w[h] etc. are written (not accumulated). */
void ocrf_oracle_gaussian_heads_backward(long long n, int F, const float* feat, const float* rgb,
                                         const float* const* fc1_w, const float* const* fc1_b,
                                         const float* const* fc2_w, const float* const* fc2_b,
                                         const float* g_opacity, const float* g_scales, const float* g_rotations,
                                         const float* g_colors, float* g_feat, float* const* g_fc1_w,
                                         float* const* g_fc1_b, float* const* g_fc2_w, float* const* g_fc2_b) {
  double* a1w[4];
  double a1b[4][4] = {{0}}, a2w[4][16] = {{0}}, a2b[4][4] = {{0}};
  for (int hd = 0; hd < 4; hd++) a1w[hd] = (double*)calloc((size_t)4 * (F + 3), sizeof(double));
  for (long long g = 0; g < n; g++) {
    for (int k = 0; k < F; k++) g_feat[g * F + k] = 0.f;
    for (int hd = 0; hd < 4; hd++) {
      const int in = hd == 3 ? F + 3 : F;
      float pre[4], hdn[4], z[4], gz[4] = {0, 0, 0, 0};
      for (int j = 0; j < 4; j++) {
        float acc = fc1_b[hd][j];
        for (int k = 0; k < F; k++) acc += feat[g * F + k] * fc1_w[hd][j * in + k];
        if (hd == 3)
          for (int c = 0; c < 3; c++) acc += rgb[g * 3 + c] * fc1_w[hd][j * in + F + c];
        pre[j] = acc;
        hdn[j] = acc > 0.f ? acc : 0.f;
      }
      for (int i = 0; i < GH_OUTS[hd]; i++) {
        float acc = fc2_b[hd][i];
        for (int j = 0; j < 4; j++) acc += fc2_w[hd][i * 4 + j] * hdn[j];
        z[i] = acc;
      }
      if (hd == 0) {
        for (int c = 0; c < 3; c++) gz[c] = g_scales[g * 3 + c] * (z[c] > 20.f ? 1.f : gh_sigmoid(z[c]));
      } else if (hd == 1) {
        const float n2 = z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3];
        const float nrm = sqrtf(n2);
        const float* gr = g_rotations + g * 4;
        if (nrm > 1e-12f) {
          const float dotv = (gr[0] * z[0] + gr[1] * z[1] + gr[2] * z[2] + gr[3] * z[3]) / n2;
          for (int c = 0; c < 4; c++) gz[c] = (gr[c] - z[c] * dotv) / nrm;
        } else {
          for (int c = 0; c < 4; c++) gz[c] = gr[c] * 1e12f;
        }
      } else if (hd == 2) {
        const float y = gh_sigmoid(z[0]);
        gz[0] = g_opacity[g] * y * (1.f - y);
      } else {
        for (int c = 0; c < 3; c++) {
          const float y = gh_sigmoid(z[c]);
          gz[c] = g_colors[g * 3 + c] * y * (1.f - y);
        }
      }
      float gh[4] = {0, 0, 0, 0};
      for (int i = 0; i < GH_OUTS[hd]; i++) {
        a2b[hd][i] += gz[i];
        for (int j = 0; j < 4; j++) {
          a2w[hd][i * 4 + j] += (double)gz[i] * hdn[j];
          gh[j] += fc2_w[hd][i * 4 + j] * gz[i];
        }
      }
      for (int j = 0; j < 4; j++) {
        if (!(pre[j] > 0.f)) continue;
        a1b[hd][j] += gh[j];
        for (int k = 0; k < F; k++) {
          a1w[hd][j * in + k] += (double)gh[j] * feat[g * F + k];
          g_feat[g * F + k] += gh[j] * fc1_w[hd][j * in + k];
        }
        if (hd == 3)
          for (int c = 0; c < 3; c++) a1w[hd][j * in + F + c] += (double)gh[j] * rgb[g * 3 + c];
      }
    }
  }
  for (int hd = 0; hd < 4; hd++) {
    const int in = hd == 3 ? F + 3 : F;
    for (int i = 0; i < 4 * in; i++) g_fc1_w[hd][i] = (float)a1w[hd][i];
    for (int j = 0; j < 4; j++) g_fc1_b[hd][j] = (float)a1b[hd][j];
    for (int i = 0; i < GH_OUTS[hd] * 4; i++) g_fc2_w[hd][i] = (float)a2w[hd][i];
    for (int i = 0; i < GH_OUTS[hd]; i++) g_fc2_b[hd][i] = (float)a2b[hd][i];
    free(a1w[hd]);
  }
}

/* ------------------------------------------------------------------------------------------
 * BEV pooling v2 (mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu).  Same loops, same summation order; the products
 * are accumulated with fmaf because nvcc contracts `psum += a * b` of the reference kernels (this file is compiled
 * with -ffp-contract=off, so the contraction is spelled out).
 * ------------------------------------------------------------------------------------------ */
/* bev_pool_cuda.cu:21-49 */
void ocrf_oracle_bev_pool_forward(int c, int n_intervals, const float* depth, const float* feat, const int* ranks_depth,
                                  const int* ranks_feat, const int* ranks_bev, const int* interval_starts,
                                  const int* interval_lengths, float* out) {
  for (int index = 0; index < n_intervals; index++) {
    const int start = interval_starts[index], len = interval_lengths[index];
    for (int cur_c = 0; cur_c < c; cur_c++) {
      float psum = 0.f;
      for (int i = 0; i < len; i++)
        psum = fmaf(feat[(size_t)ranks_feat[start + i] * c + cur_c], depth[ranks_depth[start + i]], psum);
      out[(size_t)ranks_bev[start] * c + cur_c] = psum;
    }
  }
}

/* bev_pool_cuda.cu:69-121; the rank arrays and intervals are the feature-sorted ones of bev_pool.py:47-60 */
void ocrf_oracle_bev_pool_backward(int c, int n_intervals, const float* out_grad, const float* depth, const float* feat,
                                   const int* ranks_depth, const int* ranks_feat, const int* ranks_bev,
                                   const int* interval_starts, const int* interval_lengths, float* depth_grad,
                                   float* feat_grad) {
  for (int idx = 0; idx < n_intervals; idx++) {
    const int start = interval_starts[idx], len = interval_lengths[idx];
    for (int i = 0; i < len; i++) {
      const float* g = out_grad + (size_t)ranks_bev[start + i] * c;
      const float* f = feat + (size_t)ranks_feat[start + i] * c;
      float grad_sum = 0.f;
      for (int cur_c = 0; cur_c < c; cur_c++) grad_sum = fmaf(g[cur_c], f[cur_c], grad_sum);
      depth_grad[ranks_depth[start + i]] = grad_sum;
    }
    for (int cur_c = 0; cur_c < c; cur_c++) {
      float grad_sum = 0.f;
      for (int i = 0; i < len; i++)
        grad_sum = fmaf(out_grad[(size_t)ranks_bev[start + i] * c + cur_c], depth[ranks_depth[start + i]], grad_sum);
      feat_grad[(size_t)ranks_feat[start] * c + cur_c] = grad_sum;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Voxel colouring and the sparse supervision image (view_transformer_ocrf.py:924-971, 1004-1022).
 * ------------------------------------------------------------------------------------------ */
/* F.grid_sample(bilinear, zeros, align_corners=True) of one point after the caller's normalisation (:929-930);
 * ATen grid_sampler_2d: ix = (xn + 1) / 2 * (W - 1), corners nw/ne/sw/se accumulated in that order. */
static void vc_bilinear(const float* img, int C, int H, int W, float x, float y, float* out) {
  const float xn = x / (float)(W - 1) * 2.f - 1.f, yn = y / (float)(H - 1) * 2.f - 1.f;
  const float ix = (xn + 1.f) / 2.f * (float)(W - 1), iy = (yn + 1.f) / 2.f * (float)(H - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy), fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
  const float w[4] = {(fx1 - ix) * (fy1 - iy), (ix - fx0) * (fy1 - iy), (fx1 - ix) * (iy - fy0), (ix - fx0) * (iy - fy0)};
  const float cx[4] = {fx0, fx1, fx0, fx1}, cy[4] = {fy0, fy0, fy1, fy1};
  for (int c = 0; c < C; c++) {
    float v = 0.f;
    for (int k = 0; k < 4; k++)
      if (cx[k] >= 0.f && cx[k] <= (float)(W - 1) && cy[k] >= 0.f && cy[k] <= (float)(H - 1))
        v += img[((size_t)c * H + (int)cy[k]) * W + (int)cx[k]] * w[k];
    out[c] = v;
  }
}

/* lidar_points_to_image_values (:924-942) + color_voxels (:945-962): avg [B,M,C], valid [B,M] */
void ocrf_oracle_color_voxels(int B, int N, long long M, int C, int H, int W, const float* coords, const uint8_t* mask,
                              const float* imgs, float divisor, float* avg, uint8_t* valid) {
  for (int b = 0; b < B; b++)
    for (long long m = 0; m < M; m++) {
      float sum[8] = {0};
      int count = 0;
      for (int n = 0; n < N; n++) {
        const size_t o = ((size_t)b * N + n) * M + m;
        if (!mask[o]) continue;
        float s[8];
        vc_bilinear(imgs + ((size_t)b * N + n) * C * H * W, C, H, W, coords[2 * o], coords[2 * o + 1], s);
        for (int c = 0; c < C; c++) sum[c] += s[c];
        count++;
      }
      for (int c = 0; c < C; c++) {
        float v = sum[c] / (float)(count > 0 ? count : 1);
        if (divisor != 1.f) v = v / divisor;
        avg[((size_t)b * M + m) * C + c] = v;
      }
      valid[(size_t)b * M + m] = count > 0;
    }
}

/* retain_valid_pixels (:1004-1022) */
void ocrf_oracle_retain_valid_pixels(int V, long long M, int C, int H, int W, const float* coords, const uint8_t* mask,
                                     const float* img, float fill, float* out) {
  const size_t HW = (size_t)H * W;
  for (size_t i = 0; i < (size_t)V * C * HW; i++) out[i] = fill;
  const long long lim = (W > H ? W : H) - 1;
  for (int v = 0; v < V; v++)
    for (long long m = 0; m < M; m++) {
      const size_t o = (size_t)v * M + m;
      if (!mask[o] || coords[2 * o] == -1.f) continue;
      long long xi = (long long)coords[2 * o], yi = (long long)coords[2 * o + 1];
      xi = xi < 0 ? 0 : (xi > lim ? lim : xi);
      yi = yi < 0 ? 0 : (yi > lim ? lim : yi);
      if (xi >= W || yi >= H) continue;
      for (int c = 0; c < C; c++) out[((size_t)v * C + c) * HW + yi * W + xi] = img[((size_t)v * C + c) * HW + yi * W + xi];
    }
}
