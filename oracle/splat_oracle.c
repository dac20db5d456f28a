/*
 * splat_oracle.c — CPU restatement of the differentiable Gaussian-splat rasterizer that
 * SplatFields calls through gaussian_renderer.render() (reference
 * gaussian_renderer/__init__.py:14,59-72,94-102,106-114).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under splatfields_b200/ may import, link or execute this
 * file.  Allowed users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * PARITY UNPINNED (rasterizer part): the arithmetic of this path lives in the third-party
 * dependency  ingra14m/depth-diff-gaussian-rasterization @ f2d8fa9921ea9a6cb9ac1c33a34ebd1b11510657
 * (pinned at reference README.md:28), whose source is NOT vendored under /root/reference and
 * ships no tests or golden vectors.  This file restates that rasterizer's published algorithm
 * (3DGS, Kerbl et al. 2023: EWA projection, 16x16 tile binning on (tile<<32 | depth bits) keys,
 * front-to-back alpha compositing, analytic backward) as surveyed in SURVEY.md Appendix A.
 * What IS pinned, against the reference's own in-tree Python helpers (tests/golden/, generated
 * by tests/golden/make_golden.py): camera/projection conventions (utils/graphics_utils.py:24-31,
 * 42-76; scene/cameras.py:62-74), SH evaluation (utils/sh_utils.py:26-112), quaternion->R and
 * Sigma = (R S)(R S)^T with the (xx,xy,xz,yy,yz,zz) packing (utils/general_utils.py:122-171).
 * The analytic backward is additionally cross-checked against fp64 autograd through an
 * independent dense restatement (oracle/torch_naive.py).
 *
 * Numerics: fp32 throughout, compiled with -ffp-contract=off so that a*b+c is two roundings and
 * fmaf() is one; the CUDA preprocess kernel is compiled with -fmad=false and follows the same
 * formula sheet (DESIGN.md "canonical op order"), which is what makes radii / tile rectangles /
 * depth key bits comparable bit-for-bit.  Backward sums over pixels are accumulated in fp64
 * (the reference accumulates with fp32 atomics in nondeterministic order; fp64 is the value any
 * such order approximates).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK_X 16
#define BLOCK_Y 16
#define BLOCK_SIZE (BLOCK_X * BLOCK_Y)

/* SH constants: reference utils/sh_utils.py:26-43 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

int so_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void so_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* dot of two 3-vectors in the canonical order fma(a2,b2, fma(a1,b1, a0*b0)) */
static inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

/* float -> int conversion with CUDA cvt.rzi.s32.f32 semantics (saturating, NaN -> 0) */
static inline int f2i_rz(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 2147483647;
  if (v <= -2147483648.0f) return (-2147483647 - 1);
  return (int)v;
}

/* p (row vector) times the transposed-storage matrix, rows 0..2 (+ translation).
 * m[4*c + r] is element (r,c) of the mathematical matrix (scene/cameras.py:68 stores W2C^T).
 * Order: ((m0*x + m4*y) + m8*z) + m12 with nvcc's contraction fma(m8,z, fma(m0,x, m4*y)) + m12. */
static inline void xform4x3(const float* m, const float* p, float* o) {
  o[0] = fmaf(m[8], p[2], fmaf(m[0], p[0], m[4] * p[1])) + m[12];
  o[1] = fmaf(m[9], p[2], fmaf(m[1], p[0], m[5] * p[1])) + m[13];
  o[2] = fmaf(m[10], p[2], fmaf(m[2], p[0], m[6] * p[1])) + m[14];
}
static inline void xform4x4(const float* m, const float* p, float* o) {
  o[0] = fmaf(m[8], p[2], fmaf(m[0], p[0], m[4] * p[1])) + m[12];
  o[1] = fmaf(m[9], p[2], fmaf(m[1], p[0], m[5] * p[1])) + m[13];
  o[2] = fmaf(m[10], p[2], fmaf(m[2], p[0], m[6] * p[1])) + m[14];
  o[3] = fmaf(m[11], p[2], fmaf(m[3], p[0], m[7] * p[1])) + m[15];
}

/* NDC -> pixel, evaluated in double like the external rasterizer's double literals
 * (cf. scene/dataset_readers.py:515-516 for the in-tree restatement). */
static inline float ndc2pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }

static inline void get_rect(float px, float py, int max_radius, int gx, int gy, int* rmin, int* rmax) {
  float r = (float)max_radius;
  int a;
  a = f2i_rz((px - r) / (float)BLOCK_X);                   rmin[0] = a < 0 ? 0 : (a > gx ? gx : a);
  a = f2i_rz((py - r) / (float)BLOCK_Y);                   rmin[1] = a < 0 ? 0 : (a > gy ? gy : a);
  a = f2i_rz((px + r + (float)(BLOCK_X - 1)) / (float)BLOCK_X); rmax[0] = a < 0 ? 0 : (a > gx ? gx : a);
  a = f2i_rz((py + r + (float)(BLOCK_Y - 1)) / (float)BLOCK_Y); rmax[1] = a < 0 ? 0 : (a > gy ? gy : a);
}

/* quaternion (r,x,y,z), NOT normalised (the rasterizer takes q as given; callers pass unit q:
 * scene/gaussian_model.py:72) -> rotation, entries as utils/general_utils.py:150-158 */
static inline void quat_to_R(const float* q, float* R) {
  float r = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = fmaf(-2.f, fmaf(y, y, z * z), 1.f);
  R[1] = 2.f * fmaf(x, y, -(r * z));
  R[2] = 2.f * fmaf(x, z, r * y);
  R[3] = 2.f * fmaf(x, y, r * z);
  R[4] = fmaf(-2.f, fmaf(x, x, z * z), 1.f);
  R[5] = 2.f * fmaf(y, z, -(r * x));
  R[6] = 2.f * fmaf(x, z, -(r * y));
  R[7] = 2.f * fmaf(y, z, r * x);
  R[8] = fmaf(-2.f, fmaf(x, x, y * y), 1.f);
}

/* Sigma = (R S)(R S)^T packed (xx,xy,xz,yy,yz,zz): utils/general_utils.py:162-171,
 * scene/gaussian_model.py:33-37 */
static inline void cov3d_from_scale_rot(const float* scale, float mod, const float* q, float* cov6) {
  float R[9], L[9];
  quat_to_R(q, R);
  float s0 = mod * scale[0], s1 = mod * scale[1], s2 = mod * scale[2];
  for (int i = 0; i < 3; i++) {
    L[3 * i + 0] = R[3 * i + 0] * s0;
    L[3 * i + 1] = R[3 * i + 1] * s1;
    L[3 * i + 2] = R[3 * i + 2] * s2;
  }
  cov6[0] = dot3(L[0], L[0], L[1], L[1], L[2], L[2]);
  cov6[1] = dot3(L[0], L[3], L[1], L[4], L[2], L[5]);
  cov6[2] = dot3(L[0], L[6], L[1], L[7], L[2], L[8]);
  cov6[3] = dot3(L[3], L[3], L[4], L[4], L[5], L[5]);
  cov6[4] = dot3(L[3], L[6], L[4], L[7], L[5], L[8]);
  cov6[5] = dot3(L[6], L[6], L[7], L[7], L[8], L[8]);
}

/* The pieces of the EWA projection shared by forward and backward. */
typedef struct {
  float t[3];       /* view-space position with clamped x,y */
  float xmul, ymul; /* 0 when the tan-fov clamp was active (gradient mask), else 1 */
  float J00, J02, J11, J12;
  float m0[3], m1[3]; /* rows of J * Rw */
} ewa_t;

static inline void ewa_setup(const float* mean, const float* view, float fx, float fy, float tanx,
                             float tany, ewa_t* e) {
  float t[3];
  xform4x3(view, mean, t);
  float limx = 1.3f * tanx, limy = 1.3f * tany;
  float txtz = t[0] / t[2], tytz = t[1] / t[2];
  e->xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  e->ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
  t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
  e->t[0] = t[0]; e->t[1] = t[1]; e->t[2] = t[2];
  float tz2 = t[2] * t[2];
  e->J00 = fx / t[2];
  e->J02 = -(fx * t[0]) / tz2;
  e->J11 = fy / t[2];
  e->J12 = -(fy * t[1]) / tz2;
  /* Rw(i,j) = view[4*j + i] */
  for (int j = 0; j < 3; j++) {
    e->m0[j] = fmaf(e->J02, view[4 * j + 2], e->J00 * view[4 * j + 0]);
    e->m1[j] = fmaf(e->J12, view[4 * j + 2], e->J11 * view[4 * j + 1]);
  }
}

static inline void ewa_cov2d(const ewa_t* e, const float* c6, float* a, float* b, float* c,
                             float* v0, float* v1) {
  /* v0 = Sigma m0, v1 = Sigma m1 */
  const float* m0 = e->m0; const float* m1 = e->m1;
  v0[0] = dot3(c6[0], m0[0], c6[1], m0[1], c6[2], m0[2]);
  v0[1] = dot3(c6[1], m0[0], c6[3], m0[1], c6[4], m0[2]);
  v0[2] = dot3(c6[2], m0[0], c6[4], m0[1], c6[5], m0[2]);
  v1[0] = dot3(c6[0], m1[0], c6[1], m1[1], c6[2], m1[2]);
  v1[1] = dot3(c6[1], m1[0], c6[3], m1[1], c6[4], m1[2]);
  v1[2] = dot3(c6[2], m1[0], c6[4], m1[1], c6[5], m1[2]);
  *a = dot3(m0[0], v0[0], m0[1], v0[1], m0[2], v0[2]) + 0.3f;
  *b = dot3(m1[0], v0[0], m1[1], v0[1], m1[2], v0[2]);
  *c = dot3(m1[0], v1[0], m1[1], v1[1], m1[2], v1[2]) + 0.3f;
}

/* SH -> RGB, formula and signs of utils/sh_utils.py:57-112, then +0.5 and clamp at 0
 * (extract_geo.py:40-44 is the in-tree restatement of that colour step). */
static inline void sh_to_rgb(int deg, int M, const float* sh /* [M][3] */, const float* mean,
                             const float* campos, float* rgb, uint8_t* clamped) {
  float dx = mean[0] - campos[0], dy = mean[1] - campos[1], dz = mean[2] - campos[2];
  float len = sqrtf(dot3(dx, dx, dy, dy, dz, dz));
  float x = dx / len, y = dy / len, z = dz / len;
  (void)M;
  for (int c = 0; c < 3; c++) {
    float r = SH_C0 * sh[0 * 3 + c];
    if (deg > 0) {
      r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
            SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
            SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
        if (deg > 2) {
          r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
              SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
              SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
              SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] +
              SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] + SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
        }
      }
    }
    r += 0.5f;
    clamped[c] = r < 0.f;
    rgb[c] = r < 0.f ? 0.f : r;
  }
}

/* ------------------------------------------------------------------------------------------
 * K1: per-Gaussian preprocess.  All output arrays are caller-allocated; radii and
 * tiles_touched are zero for culled Gaussians, other outputs are then left untouched (zeros
 * if the caller zero-filled).
 * ---------------------------------------------------------------------------------------- */
void so_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                   const float* rotations, const float* opacities, const float* shs,
                   const float* colors_precomp, const float* cov3D_precomp, const float* viewmatrix,
                   const float* projmatrix, const float* campos, int W, int H, float tan_fovx,
                   float tan_fovy, int* radii, float* means2D, float* depths, float* cov3Ds, float* rgb,
                   float* conic_opacity, uint32_t* tiles_touched, uint8_t* clamped) {
  const float focal_x = (float)W / (2.0f * tan_fovx);
  const float focal_y = (float)H / (2.0f * tan_fovy);
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < P; idx++) {
    radii[idx] = 0;
    tiles_touched[idx] = 0;
    const float* p = means3D + 3 * idx;
    float p_view[3];
    xform4x3(viewmatrix, p, p_view);
    if (p_view[2] <= 0.2f) continue; /* near cull; no x/y frustum test */
    float p_hom[4];
    xform4x4(projmatrix, p, p_hom);
    float p_w = 1.0f / (p_hom[3] + 0.0000001f); /* utils/graphics_utils.py:30 */
    float p_proj_x = p_hom[0] * p_w, p_proj_y = p_hom[1] * p_w;

    float c6[6];
    if (cov3D_precomp) {
      memcpy(c6, cov3D_precomp + 6 * idx, sizeof(c6));
    } else {
      cov3d_from_scale_rot(scales + 3 * idx, scale_modifier, rotations + 4 * idx, c6);
    }
    memcpy(cov3Ds + 6 * idx, c6, sizeof(c6));

    ewa_t e;
    ewa_setup(p, viewmatrix, focal_x, focal_y, tan_fovx, tan_fovy, &e);
    float a, b, c, v0[3], v1[3];
    ewa_cov2d(&e, c6, &a, &b, &c, v0, v1);
    float det = fmaf(a, c, -(b * b));
    if (det == 0.0f) continue;
    float det_inv = 1.f / det;
    float conic[3] = {c * det_inv, -b * det_inv, a * det_inv};
    float mid = 0.5f * (a + c);
    float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
    float lambda1 = mid + sq, lambda2 = mid - sq;
    float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    float pix = ndc2pix(p_proj_x, W), piy = ndc2pix(p_proj_y, H);
    int rmin[2], rmax[2];
    get_rect(pix, piy, f2i_rz(my_radius), gx, gy, rmin, rmax);
    if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;

    if (colors_precomp) {
      rgb[3 * idx + 0] = colors_precomp[3 * idx + 0];
      rgb[3 * idx + 1] = colors_precomp[3 * idx + 1];
      rgb[3 * idx + 2] = colors_precomp[3 * idx + 2];
      clamped[3 * idx + 0] = clamped[3 * idx + 1] = clamped[3 * idx + 2] = 0;
    } else {
      sh_to_rgb(D, M, shs + (size_t)idx * M * 3, p, campos, rgb + 3 * idx, clamped + 3 * idx);
    }
    depths[idx] = p_view[2];
    radii[idx] = f2i_rz(my_radius);
    means2D[2 * idx + 0] = pix;
    means2D[2 * idx + 1] = piy;
    conic_opacity[4 * idx + 0] = conic[0];
    conic_opacity[4 * idx + 1] = conic[1];
    conic_opacity[4 * idx + 2] = conic[2];
    conic_opacity[4 * idx + 3] = opacities[idx];
    tiles_touched[idx] = (uint32_t)((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]));
  }
}

/* K10 markVisible: view z > 0.2 */
void so_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present) {
  for (int idx = 0; idx < P; idx++) {
    float pv[3];
    xform4x3(viewmatrix, means3D + 3 * idx, pv);
    present[idx] = pv[2] > 0.2f;
  }
}

/* ------------------------------------------------------------------------------------------
 * K2..K5: scan, duplicateWithKeys, stable sort on (tile<<32 | depth bits), tile ranges
 * (SURVEY A.3, A.4).  Two-call protocol: so_count_rendered gives R, the caller allocates.
 * ---------------------------------------------------------------------------------------- */
int64_t so_count_rendered(int P, const uint32_t* tiles_touched) {
  int64_t r = 0;
  for (int i = 0; i < P; i++) r += tiles_touched[i];
  return r;
}

/* Stable LSD radix sort of (key, value) pairs on key bits [0, bits), 8 bits per pass, parallel over contiguous
 * chunks of the input: per-chunk digit histograms, offsets in (digit, chunk) order — so equal digits keep their
 * input order across chunks — then every chunk scatters its own items.  The number of passes is made even so that
 * the result lands back in keys / vals. */
static void radix_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint64_t* ktmp, uint32_t* vtmp, size_t n,
                                 int bits) {
  int npass = (bits + 7) / 8;
  if (npass & 1) npass++;
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  if (n < (size_t)(1 << 16)) nth = 1;
  size_t* cnt = (size_t*)malloc(sizeof(size_t) * 256 * (size_t)nth);
  for (int pass = 0; pass < npass; pass++) {
    const int shift = 8 * pass;
    memset(cnt, 0, sizeof(size_t) * 256 * (size_t)nth);
#pragma omp parallel num_threads(nth)
    {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      const size_t lo = n * (size_t)t / (size_t)nth, hi = n * (size_t)(t + 1) / (size_t)nth;
      size_t* c = cnt + 256 * (size_t)t;
      for (size_t i = lo; i < hi; i++) c[(keys[i] >> shift) & 0xFF]++;
#pragma omp barrier
#pragma omp single
      {
        size_t run = 0;
        for (int b = 0; b < 256; b++)
          for (int u = 0; u < nth; u++) { size_t v = cnt[256 * (size_t)u + b]; cnt[256 * (size_t)u + b] = run; run += v; }
      }
      for (size_t i = lo; i < hi; i++) {
        const size_t d = c[(keys[i] >> shift) & 0xFF]++;
        ktmp[d] = keys[i];
        vtmp[d] = vals[i];
      }
    }
    uint64_t* kt = keys; keys = ktmp; ktmp = kt;
    uint32_t* vt = vals; vals = vtmp; vtmp = vt;
  }
  free(cnt);
}

void so_bin(int P, int W, int H, const float* means2D, const float* depths, const int* radii,
            const uint32_t* tiles_touched, int64_t R, uint64_t* keys_sorted, uint32_t* vals_sorted,
            uint32_t* ranges /* [T][2] */) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
  const int T = gx * gy;
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)T);
  if (R == 0) return;
  uint64_t* ktmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)R);
  uint32_t* vtmp = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
  /* K2: exclusive scan of tiles_touched = where each Gaussian's instances start (index order, like the reference's
   * InclusiveSum); K3: every Gaussian writes its own run, so the emission is parallel. */
  size_t* start = (size_t*)malloc(sizeof(size_t) * ((size_t)P + 1));
  start[0] = 0;
  for (int i = 0; i < P; i++) start[i + 1] = start[i] + (radii[i] > 0 ? tiles_touched[i] : 0);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int idx = 0; idx < P; idx++) {
    if (radii[idx] <= 0) continue;
    int rmin[2], rmax[2];
    get_rect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], gx, gy, rmin, rmax);
    uint32_t dbits;
    memcpy(&dbits, depths + idx, 4);
    size_t off = start[idx];
    for (int y = rmin[1]; y < rmax[1]; y++)
      for (int x = rmin[0]; x < rmax[0]; x++) {
        uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
        key = (key << 32) | dbits;
        keys_sorted[off] = key;
        vals_sorted[off] = (uint32_t)idx;
        off++;
      }
  }
  free(start);
  /* K4: stable LSD sort on bits [0, 32 + ceil(log2 T)) — the range the reference hands to CUB */
  int tbits = 0;
  while ((1 << tbits) < T) tbits++;
  radix_sort_pairs_u64(keys_sorted, vals_sorted, ktmp, vtmp, (size_t)R, 32 + tbits);
  free(ktmp);
  free(vtmp);
  /* K5: every boundary between two tiles writes one end and one start; disjoint entries, so parallel */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < R; i++) {
    uint32_t tile = (uint32_t)(keys_sorted[i] >> 32);
    if (i == 0) ranges[2 * tile] = 0;
    else {
      uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
      if (tile != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * tile] = (uint32_t)i; }
    }
    if (i == R - 1) ranges[2 * tile + 1] = (uint32_t)R;
  }
}

/* ------------------------------------------------------------------------------------------
 * K6: forward compositing.  margin[pix] (optional) = smallest relative distance of
 * any decision this pixel took (alpha vs 1/255, T' vs 1e-4, power vs 0) from its threshold — the
 * parity tests use it to exclude pixels whose outcome legitimately depends on the last ulp of
 * expf (glibc vs CUDA libdevice).
 * ---------------------------------------------------------------------------------------- */
void so_render_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                       const float* means2D, const float* rgb, const float* depths,
                       const float* conic_opacity, const float* bg, float* out_color /* [3][H][W] */,
                       float* out_depth /* [H][W] */, float* final_T, uint32_t* n_contrib,
                       float* margin /* may be NULL */) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    int tx = tile % gx, ty = tile / gx;
    uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < BLOCK_Y; ly++)
      for (int lx = 0; lx < BLOCK_X; lx++) {
        int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
        if (px >= W || py >= H) continue;
        size_t pix_id = (size_t)W * py + px;
        float pixfx = (float)px, pixfy = (float)py;
        float T = 1.0f, C[3] = {0, 0, 0}, Dp = 0.f;
        uint32_t contributor = 0, last_contributor = 0;
        float mg = 1e30f;
        for (uint32_t k = r0; k < r1; k++) {
          contributor++;
          uint32_t g = point_list[k];
          float dx = means2D[2 * g] - pixfx, dy = means2D[2 * g + 1] - pixfy;
          const float* co = conic_opacity + 4 * g;
          /* -0.5f*(A*dx*dx + C*dy*dy) - B*dx*dy with nvcc's contraction */
          float s = fmaf(co[0] * dx, dx, (co[2] * dy) * dy);
          float power = fmaf(s, -0.5f, -((co[1] * dx) * dy));
          if (power > 0.0f) continue;
          float ex = expf(power);
          float alpha = fminf(0.99f, co[3] * ex);
          if (margin) {
            float m = fabsf(alpha - 1.0f / 255.0f) * 255.0f;
            if (m < mg) mg = m;
          }
          if (alpha < 1.0f / 255.0f) continue;
          float test_T = T * (1 - alpha);
          if (margin) {
            float m = fabsf(test_T - 0.0001f) * 10000.0f;
            if (m < mg) mg = m;
          }
          if (test_T < 0.0001f) break;
          float w = alpha * T;
          for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(rgb[3 * g + ch], w, C[ch]);
          Dp = fmaf(depths[g], w, Dp);
          T = test_T;
          last_contributor = contributor;
        }
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pix_id] = fmaf(T, bg[ch], C[ch]);
        out_depth[pix_id] = Dp;
        if (margin) margin[pix_id] = mg;
      }
  }
}

/* ------------------------------------------------------------------------------------------
 * K7: backward compositing (SURVEY A.5 second half).  Accumulators are fp64:
 *   dL_dmean2D [P][2]  (gradient w.r.t. the NDC-scaled screen mean, i.e. includes 0.5*W, 0.5*H —
 *                       this is what lands in viewspace_points.grad, scene/gaussian_model.py:429)
 *   dL_dconic  [P][3]  TRUE derivatives w.r.t. (A,B,C) of power = -0.5(A dx^2 + C dy^2) - B dx dy
 *   dL_dopacity[P], dL_dcolor [P][3]
 * ---------------------------------------------------------------------------------------- */
void so_render_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                        const float* means2D, const float* rgb, const float* conic_opacity,
                        const float* bg, const float* final_T, const uint32_t* n_contrib,
                        const float* dL_dpixels /* [3][H][W] */, double* dL_dmean2D, double* dL_dconic,
                        double* dL_dopacity, double* dL_dcolor,
                        const float* depths /* [P] view-space z, or NULL */,
                        const float* dL_ddepth_img /* [H][W] cotangent of the depth image, or NULL */,
                        double* dL_ddepth /* [P] out (with dL_ddepth_img) */) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  /* The pixel sums of one tile are collected per list position in a thread-local fp64 table and added to the
   * per-Gaussian totals once per (tile, Gaussian) — 9 atomics per instance instead of 9 per (pixel, contributor),
   * which is what makes this port usable as a CPU baseline at 1M Gaussians.  (fp64 sums: the grouping of the
   * additions is immaterial at the tolerances the tests use.) */
  uint32_t max_len = 0;
  for (int tile = 0; tile < gx * gy; tile++) {
    uint32_t len = ranges[2 * tile + 1] - ranges[2 * tile];
    if (len > max_len) max_len = len;
  }
#pragma omp parallel
  {
  /* depth cotangent (the depth image is one more composited channel, "colour" = the splat's view-space z, no
   * background): tenth sum per list position */
  const int with_depth = depths && dL_ddepth_img && dL_ddepth;
  double* loc = (double*)malloc(sizeof(double) * 10 * ((size_t)max_len + 1));
#pragma omp for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    int tx = tile % gx, ty = tile / gx;
    uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    uint32_t deepest = 0;
    for (int ly = 0; ly < BLOCK_Y; ly++)
      for (int lx = 0; lx < BLOCK_X; lx++) {
        int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
        if (px < W && py < H && n_contrib[(size_t)W * py + px] > deepest) deepest = n_contrib[(size_t)W * py + px];
      }
    if (deepest > r1 - r0) deepest = r1 - r0;
    memset(loc, 0, sizeof(double) * 10 * (size_t)deepest);
    for (int ly = 0; ly < BLOCK_Y; ly++)
      for (int lx = 0; lx < BLOCK_X; lx++) {
        int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
        if (px >= W || py >= H) continue;
        size_t pix_id = (size_t)W * py + px;
        float pixfx = (float)px, pixfy = (float)py;
        const float T_final = final_T[pix_id];
        float T = T_final;
        uint32_t last = n_contrib[pix_id];
        float dLp[3], accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f;
        for (int ch = 0; ch < 3; ch++) dLp[ch] = dL_dpixels[(size_t)ch * H * W + pix_id];
        float bg_dot = 0.f;
        for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dLp[ch];
        const float dLpd = with_depth ? dL_ddepth_img[pix_id] : 0.f;
        float accum_depth = 0.f, last_depth = 0.f;
        /* walk positions last-1 .. 0 of this tile's list (position = contributor index - 1) */
        for (int64_t pos = (int64_t)last - 1; pos >= 0; pos--) {
          uint32_t g = point_list[r0 + pos];
          (void)r1;
          float dx = means2D[2 * g] - pixfx, dy = means2D[2 * g + 1] - pixfy;
          const float* co = conic_opacity + 4 * g;
          float s = fmaf(co[0] * dx, dx, (co[2] * dy) * dy);
          float power = fmaf(s, -0.5f, -((co[1] * dx) * dy));
          if (power > 0.0f) continue;
          float G = expf(power);
          float alpha = fminf(0.99f, co[3] * G);
          if (alpha < 1.0f / 255.0f) continue;
          T = T / (1.f - alpha);
          float dchannel_dcolor = alpha * T;
          float dL_dalpha = 0.f;
          for (int ch = 0; ch < 3; ch++) {
            float c = rgb[3 * g + ch];
            accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
            last_color[ch] = c;
            dL_dalpha += (c - accum_rec[ch]) * dLp[ch];
            loc[10 * pos + 6 + ch] += (double)(dchannel_dcolor * dLp[ch]);
          }
          if (with_depth) {
            const float z = depths[g];
            accum_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_depth;
            last_depth = z;
            dL_dalpha += (z - accum_depth) * dLpd;
            loc[10 * pos + 9] += (double)(dchannel_dcolor * dLpd);
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
          float dL_dG = co[3] * dL_dalpha;
          float gdx = G * dx, gdy = G * dy;
          float dG_ddelx = -gdx * co[0] - gdy * co[1];
          float dG_ddely = -gdy * co[2] - gdx * co[1];
          double v0 = (double)(dL_dG * dG_ddelx * ddelx_dx), v1 = (double)(dL_dG * dG_ddely * ddely_dy);
          double cA = (double)(-0.5f * gdx * dx * dL_dG), cB = (double)(-gdx * dy * dL_dG),
                 cC = (double)(-0.5f * gdy * dy * dL_dG);
          double vo = (double)(G * dL_dalpha);
          double* l = loc + 10 * pos;
          l[0] += v0; l[1] += v1; l[2] += cA; l[3] += cB; l[4] += cC; l[5] += vo;
        }
      }
    for (uint32_t pos = 0; pos < deepest; pos++) {
      const double* l = loc + 10 * (size_t)pos;
      uint32_t g = point_list[r0 + pos];
      int any = 0;
      for (int k = 0; k < 10; k++) any |= (l[k] != 0.0);
      if (!any) continue;
      if (with_depth) {
#pragma omp atomic
        dL_ddepth[g] += l[9];
      }
#pragma omp atomic
      dL_dmean2D[2 * g] += l[0];
#pragma omp atomic
      dL_dmean2D[2 * g + 1] += l[1];
#pragma omp atomic
      dL_dconic[3 * g] += l[2];
#pragma omp atomic
      dL_dconic[3 * g + 1] += l[3];
#pragma omp atomic
      dL_dconic[3 * g + 2] += l[4];
#pragma omp atomic
      dL_dopacity[g] += l[5];
#pragma omp atomic
      dL_dcolor[3 * g] += l[6];
#pragma omp atomic
      dL_dcolor[3 * g + 1] += l[7];
#pragma omp atomic
      dL_dcolor[3 * g + 2] += l[8];
    }
  }
  free(loc);
  }
}

/* ------------------------------------------------------------------------------------------
 * K8 + K9: per-Gaussian backward.  Inputs are the fp32-rounded pixel sums.
 * Outputs: dL_dmeans3D [P][3], dL_dcov3D [P][6], dL_dsh [P][M][3] (if shs), dL_dscales [P][3],
 * dL_drot [P][4] (if scales/rotations).  All outputs must be zero-filled by the caller.
 * ---------------------------------------------------------------------------------------- */
void so_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                            const uint8_t* clamped, const float* scales, const float* rotations,
                            float scale_modifier, const float* cov3Ds, int cov_is_precomp,
                            const float* viewmatrix, const float* projmatrix, const float* campos, int W,
                            int H, float tan_fovx, float tan_fovy, const float* dL_dmean2D /* [P][2] */,
                            const float* dL_dconic /* [P][3] */, const float* dL_dcolor /* [P][3] */,
                            float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscales,
                            float* dL_drot, const float* dL_ddepth /* [P] or NULL: gradient w.r.t. the view-space z */) {
  const float fx = (float)W / (2.0f * tan_fovx), fy = (float)H / (2.0f * tan_fovy);
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < P; idx++) {
    if (!(radii[idx] > 0)) continue;
    const float* mean = means3D + 3 * idx;
    const float* c6 = cov3Ds + 6 * idx;
    float dmean[3] = {0, 0, 0};

    /* ---- conic -> cov2D -> (Sigma3D, t) ---- */
    ewa_t e;
    ewa_setup(mean, viewmatrix, fx, fy, tan_fovx, tan_fovy, &e);
    float a, b, c, v0[3], v1[3];
    ewa_cov2d(&e, c6, &a, &b, &c, v0, v1);
    float gA = dL_dconic[3 * idx], gB = dL_dconic[3 * idx + 1], gC = dL_dconic[3 * idx + 2];
    float denom = a * c - b * b;
    float denom2inv = 1.0f / ((denom * denom) + 0.0000001f); /* external rasterizer's guard */
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    float dcov[6] = {0, 0, 0, 0, 0, 0};
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * gA + b * c * gB + (denom - a * c) * gC);
      dL_dc = denom2inv * (-a * a * gC + a * b * gB + (denom - a * c) * gA);
      dL_db = denom2inv * (2 * b * c * gA - (denom + 2 * b * b) * gB + 2 * a * b * gC);
      const float* m0 = e.m0; const float* m1 = e.m1;
      dcov[0] = m0[0] * m0[0] * dL_da + m0[0] * m1[0] * dL_db + m1[0] * m1[0] * dL_dc;
      dcov[3] = m0[1] * m0[1] * dL_da + m0[1] * m1[1] * dL_db + m1[1] * m1[1] * dL_dc;
      dcov[5] = m0[2] * m0[2] * dL_da + m0[2] * m1[2] * dL_db + m1[2] * m1[2] * dL_dc;
      dcov[1] = 2 * m0[0] * m0[1] * dL_da + (m0[0] * m1[1] + m0[1] * m1[0]) * dL_db + 2 * m1[0] * m1[1] * dL_dc;
      dcov[2] = 2 * m0[0] * m0[2] * dL_da + (m0[0] * m1[2] + m0[2] * m1[0]) * dL_db + 2 * m1[0] * m1[2] * dL_dc;
      dcov[4] = 2 * m0[2] * m0[1] * dL_da + (m0[1] * m1[2] + m0[2] * m1[1]) * dL_db + 2 * m1[1] * m1[2] * dL_dc;
    }
    for (int k = 0; k < 6; k++) dL_dcov3D[6 * idx + k] = dcov[k];
    /* d/dm0 = 2 dL_da Sigma m0 + dL_db Sigma m1 ; d/dm1 = 2 dL_dc Sigma m1 + dL_db Sigma m0 */
    float dm0[3], dm1[3];
    for (int j = 0; j < 3; j++) {
      dm0[j] = 2 * v0[j] * dL_da + v1[j] * dL_db;
      dm1[j] = 2 * v1[j] * dL_dc + v0[j] * dL_db;
    }
    float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
    for (int j = 0; j < 3; j++) {
      dJ00 += viewmatrix[4 * j + 0] * dm0[j];
      dJ02 += viewmatrix[4 * j + 2] * dm0[j];
      dJ11 += viewmatrix[4 * j + 1] * dm1[j];
      dJ12 += viewmatrix[4 * j + 2] * dm1[j];
    }
    float tz = 1.f / e.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    float dtx = e.xmul * -fx * tz2 * dJ02;
    float dty = e.ymul * -fy * tz2 * dJ12;
    float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * e.t[0]) * tz3 * dJ02 + (2 * fy * e.t[1]) * tz3 * dJ12;
    /* dL/dmean = Rw^T dL/dt :  Rw(i,j) = view[4*j+i] */
    for (int j = 0; j < 3; j++)
      dmean[j] = viewmatrix[4 * j + 0] * dtx + viewmatrix[4 * j + 1] * dty + viewmatrix[4 * j + 2] * dtz;
    /* depth = view[2] x + view[6] y + view[10] z + view[14] */
    if (dL_ddepth)
      for (int j = 0; j < 3; j++) dmean[j] += viewmatrix[4 * j + 2] * dL_ddepth[idx];

    /* ---- screen mean (NDC) -> mean3D ---- */
    {
      float m_hom[4];
      xform4x4(projmatrix, mean, m_hom);
      float m_w = 1.0f / (m_hom[3] + 0.0000001f);
      float mul1 = m_hom[0] * m_w * m_w, mul2 = m_hom[1] * m_w * m_w;
      float g0 = dL_dmean2D[2 * idx], g1 = dL_dmean2D[2 * idx + 1];
      for (int j = 0; j < 3; j++) {
        dmean[j] += (projmatrix[4 * j + 0] * m_w - projmatrix[4 * j + 3] * mul1) * g0 +
                    (projmatrix[4 * j + 1] * m_w - projmatrix[4 * j + 3] * mul2) * g1;
      }
    }

    /* ---- colour -> SH coefficients and view direction ---- */
    if (shs) {
      const float* sh = shs + (size_t)idx * M * 3;
      float* dsh = dL_dsh + (size_t)idx * M * 3;
      float vx = mean[0] - campos[0], vy = mean[1] - campos[1], vz = mean[2] - campos[2];
      float len = sqrtf(dot3(vx, vx, vy, vy, vz, vz));
      float x = vx / len, y = vy / len, z = vz / len;
      float g[3];
      for (int ch = 0; ch < 3; ch++) g[ch] = clamped[3 * idx + ch] ? 0.f : dL_dcolor[3 * idx + ch];
      float ddx = 0, ddy = 0, ddz = 0; /* dL/ddir */
      for (int ch = 0; ch < 3; ch++) {
        float gc = g[ch];
        float rx = 0, ry = 0, rz = 0; /* dRGB_ch/d{x,y,z} */
        dsh[0 * 3 + ch] = SH_C0 * gc;
        if (D > 0) {
          dsh[1 * 3 + ch] = -SH_C1 * y * gc;
          dsh[2 * 3 + ch] = SH_C1 * z * gc;
          dsh[3 * 3 + ch] = -SH_C1 * x * gc;
          rx = -SH_C1 * sh[3 * 3 + ch];
          ry = -SH_C1 * sh[1 * 3 + ch];
          rz = SH_C1 * sh[2 * 3 + ch];
          if (D > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            dsh[4 * 3 + ch] = SH_C2[0] * xy * gc;
            dsh[5 * 3 + ch] = SH_C2[1] * yz * gc;
            dsh[6 * 3 + ch] = SH_C2[2] * (2.f * zz - xx - yy) * gc;
            dsh[7 * 3 + ch] = SH_C2[3] * xz * gc;
            dsh[8 * 3 + ch] = SH_C2[4] * (xx - yy) * gc;
            rx += SH_C2[0] * y * sh[4 * 3 + ch] + SH_C2[2] * 2.f * -x * sh[6 * 3 + ch] +
                  SH_C2[3] * z * sh[7 * 3 + ch] + SH_C2[4] * 2.f * x * sh[8 * 3 + ch];
            ry += SH_C2[0] * x * sh[4 * 3 + ch] + SH_C2[1] * z * sh[5 * 3 + ch] +
                  SH_C2[2] * 2.f * -y * sh[6 * 3 + ch] + SH_C2[4] * 2.f * -y * sh[8 * 3 + ch];
            rz += SH_C2[1] * y * sh[5 * 3 + ch] + SH_C2[2] * 2.f * 2.f * z * sh[6 * 3 + ch] +
                  SH_C2[3] * x * sh[7 * 3 + ch];
            if (D > 2) {
              dsh[9 * 3 + ch] = SH_C3[0] * y * (3.f * xx - yy) * gc;
              dsh[10 * 3 + ch] = SH_C3[1] * xy * z * gc;
              dsh[11 * 3 + ch] = SH_C3[2] * y * (4.f * zz - xx - yy) * gc;
              dsh[12 * 3 + ch] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * gc;
              dsh[13 * 3 + ch] = SH_C3[4] * x * (4.f * zz - xx - yy) * gc;
              dsh[14 * 3 + ch] = SH_C3[5] * z * (xx - yy) * gc;
              dsh[15 * 3 + ch] = SH_C3[6] * x * (xx - 3.f * yy) * gc;
              rx += SH_C3[0] * sh[9 * 3 + ch] * 3.f * 2.f * xy + SH_C3[1] * sh[10 * 3 + ch] * yz +
                    SH_C3[2] * sh[11 * 3 + ch] * -2.f * xy + SH_C3[3] * sh[12 * 3 + ch] * -3.f * 2.f * xz +
                    SH_C3[4] * sh[13 * 3 + ch] * (-3.f * xx + 4.f * zz - yy) +
                    SH_C3[5] * sh[14 * 3 + ch] * 2.f * xz + SH_C3[6] * sh[15 * 3 + ch] * 3.f * (xx - yy);
              ry += SH_C3[0] * sh[9 * 3 + ch] * 3.f * (xx - yy) + SH_C3[1] * sh[10 * 3 + ch] * xz +
                    SH_C3[2] * sh[11 * 3 + ch] * (-3.f * yy + 4.f * zz - xx) +
                    SH_C3[3] * sh[12 * 3 + ch] * -3.f * 2.f * yz + SH_C3[4] * sh[13 * 3 + ch] * -2.f * xy +
                    SH_C3[5] * sh[14 * 3 + ch] * -2.f * yz + SH_C3[6] * sh[15 * 3 + ch] * -3.f * 2.f * xy;
              rz += SH_C3[1] * sh[10 * 3 + ch] * xy + SH_C3[2] * sh[11 * 3 + ch] * 4.f * 2.f * yz +
                    SH_C3[3] * sh[12 * 3 + ch] * 3.f * (2.f * zz - xx - yy) +
                    SH_C3[4] * sh[13 * 3 + ch] * 4.f * 2.f * xz + SH_C3[5] * sh[14 * 3 + ch] * (xx - yy);
            }
          }
        }
        ddx += rx * gc; ddy += ry * gc; ddz += rz * gc;
      }
      /* through dir = v/|v| */
      float sum2 = vx * vx + vy * vy + vz * vz;
      float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmean[0] += ((sum2 - vx * vx) * ddx - vy * vx * ddy - vz * vx * ddz) * invsum32;
      dmean[1] += (-vx * vy * ddx + (sum2 - vy * vy) * ddy - vz * vy * ddz) * invsum32;
      dmean[2] += (-vx * vz * ddx - vy * vz * ddy + (sum2 - vz * vz) * ddz) * invsum32;
    }
    dL_dmeans3D[3 * idx + 0] = dmean[0];
    dL_dmeans3D[3 * idx + 1] = dmean[1];
    dL_dmeans3D[3 * idx + 2] = dmean[2];

    /* ---- Sigma3D -> scale, quaternion ---- */
    if (!cov_is_precomp) {
      const float* q = rotations + 4 * idx;
      float R[9];
      quat_to_R(q, R);
      float s[3] = {scale_modifier * scales[3 * idx], scale_modifier * scales[3 * idx + 1],
                    scale_modifier * scales[3 * idx + 2]};
      /* G = dL/dSigma as a full symmetric matrix (off-diagonals halved) */
      float Gm[9] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                     0.5f * dcov[2], 0.5f * dcov[4], dcov[5]};
      float dLm[9]; /* dL/dL = 2 G L, L = R diag(s) */
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          float acc = 0;
          for (int k = 0; k < 3; k++) acc += Gm[3 * i + k] * (R[3 * k + j] * s[j]);
          dLm[3 * i + j] = 2.f * acc;
        }
      float Dr[9];
      for (int j = 0; j < 3; j++) {
        float ds = 0;
        for (int i = 0; i < 3; i++) {
          ds += dLm[3 * i + j] * R[3 * i + j];
          Dr[3 * i + j] = dLm[3 * i + j] * s[j];
        }
        dL_dscales[3 * idx + j] = scale_modifier * ds;
      }
      float r = q[0], x = q[1], y = q[2], z = q[3];
      dL_drot[4 * idx + 0] = 2.f * (z * (Dr[3] - Dr[1]) + y * (Dr[2] - Dr[6]) + x * (Dr[7] - Dr[5]));
      dL_drot[4 * idx + 1] = 2.f * (y * (Dr[1] + Dr[3]) + z * (Dr[2] + Dr[6]) + r * (Dr[7] - Dr[5])) - 4.f * x * (Dr[4] + Dr[8]);
      dL_drot[4 * idx + 2] = 2.f * (x * (Dr[1] + Dr[3]) + r * (Dr[2] - Dr[6]) + z * (Dr[5] + Dr[7])) - 4.f * y * (Dr[0] + Dr[8]);
      dL_drot[4 * idx + 3] = 2.f * (r * (Dr[3] - Dr[1]) + x * (Dr[2] + Dr[6]) + y * (Dr[5] + Dr[7])) - 4.f * z * (Dr[0] + Dr[4]);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * simple_knn's distCUDA2 (un-vendored dependency pinned at reference README.md:29,
 * gitlab.inria.fr/bkerbl/simple-knn @ 44f7642; only call site scene/gaussian_model.py:105).
 * PARITY UNPINNED (no source, tests or vectors in the reference tree); published behaviour restated:
 * for every point the mean of the squared distances to its 3 nearest OTHER points (self excluded by
 * index; the running best-3 list starts at FLT_MAX, so P < 4 leaves FLT_MAX terms in the mean).
 * Brute force O(P^2); the squared distance is evaluated as fma(dz,dz, fma(dy,dy, dx*dx)) in fp32 —
 * the contraction nvcc applies to simple_knn's `d.x*d.x + d.y*d.y + d.z*d.z` — so the CUDA kernel,
 * which prunes exactly, must agree bit for bit. */
#include <float.h>
void so_knn3_mean_dist2(int P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    const float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    for (int j = 0; j < P; j++) {
      if (j == i) continue;
      const float dx = pts[3 * j] - px, dy = pts[3 * j + 1] - py, dz = pts[3 * j + 2] - pz;
      float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      for (int k = 0; k < 3; k++) {
        if (best[k] > d) { const float t = best[k]; best[k] = d; d = t; }
      }
    }
    out[i] = (best[0] + best[1] + best[2]) / 3.0f;
  }
}
