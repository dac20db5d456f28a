// preprocess.cu — K1: per-Gaussian EWA projection, frustum cull, tile rectangle, SH -> RGB.
// Replaces the external rasterizer's forward preprocess; the
// conventions it must honour are the reference's own: row-vector matrices (scene/cameras.py:68-73),
// w + 1e-7 (utils/graphics_utils.py:30), quaternion/covariance (utils/general_utils.py:138-171),
// SH basis (utils/sh_utils.py:26-112).
//
// THIS FILE IS COMPILED WITH -fmad=false: a*b+c is two roundings, fmaf() is the only fused op.  The
// radius, the tile rectangle and the depth bits feed the integer tile keys, which have to be
// bit-identical with the CPU oracle (gcc -ffp-contract=off) — DESIGN.md "canonical op order".
// The kernel is HBM-bound (≈236 B in, ≈110 B out per Gaussian); the extra FMULs are free.
#include "common.cuh"
#include <cstdlib>

namespace sfb {

__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

__device__ __forceinline__ float3 xform4x3(const float* __restrict__ m, float3 p) {
  float3 o;
  o.x = fmaf(m[8], p.z, fmaf(m[0], p.x, m[4] * p.y)) + m[12];
  o.y = fmaf(m[9], p.z, fmaf(m[1], p.x, m[5] * p.y)) + m[13];
  o.z = fmaf(m[10], p.z, fmaf(m[2], p.x, m[6] * p.y)) + m[14];
  return o;
}
__device__ __forceinline__ float4 xform4x4(const float* __restrict__ m, float3 p) {
  float4 o;
  o.x = fmaf(m[8], p.z, fmaf(m[0], p.x, m[4] * p.y)) + m[12];
  o.y = fmaf(m[9], p.z, fmaf(m[1], p.x, m[5] * p.y)) + m[13];
  o.z = fmaf(m[10], p.z, fmaf(m[2], p.x, m[6] * p.y)) + m[14];
  o.w = fmaf(m[11], p.z, fmaf(m[3], p.x, m[7] * p.y)) + m[15];
  return o;
}

__device__ __forceinline__ float ndc2pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
__constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

// The camera block is tiny and read by every thread: keep it in shared memory.
struct Cam {
  float view[16], proj[16], campos[3];
};

template <int D, bool VEC>
__device__ __forceinline__ void load_sh(const float* __restrict__ shs, int idx, int M, float* sh, bool wide256) {
  constexpr int NF = 3 * (D + 1) * (D + 1);
  const float* base = shs + (size_t)idx * M * 3;
  if (VEC && wide256 && NF % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NF / 8; i++) ldg256(base + 8 * i, sh + 8 * i);
  } else if (VEC) {
    constexpr int NV = (NF + 3) / 4;
    const float4* b4 = reinterpret_cast<const float4*>(base);
#pragma unroll
    for (int i = 0; i < NV; i++) {
      float4 v = __ldg(b4 + i);
      sh[4 * i + 0] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NF; i++) sh[i] = __ldg(base + i);
  }
}

template <int D>
__device__ __forceinline__ void sh_to_rgb(const float* sh, float3 mean, const float* campos, float* rgb,
                                          uint8_t& clamp_mask) {
  float dx = mean.x - campos[0], dy = mean.y - campos[1], dz = mean.z - campos[2];
  float len = sqrtf(dot3(dx, dx, dy, dy, dz, dz));
  float x = dx / len, y = dy / len, z = dz / len;
  clamp_mask = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float r = SH_C0 * sh[0 * 3 + c];
    if (D > 0) {
      r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
      if (D > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
            SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
            SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
        if (D > 2) {
          r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
              SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
              SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
              SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] +
              SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] + SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
        }
      }
    }
    r += 0.5f;
    if (r < 0.f) clamp_mask |= (uint8_t)(1u << c);
    rgb[c] = r < 0.f ? 0.f : r;
  }
}

// (A bulk-copy (cp.async.bulk) staging of the block's SH slab was measured equal to the L2-prefetch + 256-bit-load path
// used here — 85 vs 86 us at 1M splats, round 1 — while costing 48 KB of shared memory per CTA; it was removed.)
// (Clipping the tile rectangle to the alpha >= 1/255 footprint box — fewer instances, same image — was an opt-in of round 1;
// it changes tiles_touched / R / the key and index buffers, i.e. it is not reference-identical, and was removed.)
template <int D, bool VEC_SH>
__device__ __forceinline__ void preprocess_body(const FwdParams& p, const GeomState& g, int* __restrict__ radii) {
  __shared__ Cam cam;
  __shared__ uint32_t s_tiles[8];
  __shared__ uint32_t s_nkey[8];
  if (p.zero_ptr) {   // clear the depth sort's scratch on the way (it runs right behind this kernel): one memset node less
    const uint32_t per = (p.zero_words + gridDim.x - 1) / gridDim.x;
    const uint32_t z0 = blockIdx.x * per, z1 = min(z0 + per, p.zero_words);
    for (uint32_t i = z0 + threadIdx.x; i < z1; i += 256) p.zero_ptr[i] = 0u;
  }
  if (threadIdx.x < 16) {
    cam.view[threadIdx.x] = p.viewmatrix[threadIdx.x];
    cam.proj[threadIdx.x] = p.projmatrix[threadIdx.x];
  }
  if (threadIdx.x < 3) cam.campos[threadIdx.x] = p.campos[threadIdx.x];
  __syncthreads();

  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  // (A look-ahead L2 prefetch of the inputs of the blocks two waves further on was measured in round 2: no gain here,
  //  +8 us in the geometry backward — the first round trip is not where the remaining time goes.)
  uint32_t my_tiles = 0, my_key = 0xFFFFFFFFu;
  if (idx < p.P) {
    int out_radius = 0;
    uint32_t key = 0xFFFFFFFFu;
    uint2 rect_packed = make_uint2(0u, 0u);
    uint8_t clamp_mask = 0;
    float3 mean = make_float3(__ldg(p.means3D + 3 * idx), __ldg(p.means3D + 3 * idx + 1),
                              __ldg(p.means3D + 3 * idx + 2));
    // Rotation, scale and opacity are requested together with the mean, before the cull decision that depends on it:
    // one memory round trip instead of three dependent ones (ncu, round 2: 23 % of the kernel's stall samples sat on the
    // first use of the scale / rotation and of the opacity).  A culled splat costs 32 more bytes of traffic (~2 %).
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (!p.cov3D_precomp) {
      q0 = __ldg(p.rotations + 4 * (size_t)idx); q1 = __ldg(p.rotations + 4 * (size_t)idx + 1);
      q2 = __ldg(p.rotations + 4 * (size_t)idx + 2); q3 = __ldg(p.rotations + 4 * (size_t)idx + 3);
      s0 = __ldg(p.scales + 3 * (size_t)idx); s1 = __ldg(p.scales + 3 * (size_t)idx + 1);
      s2 = __ldg(p.scales + 3 * (size_t)idx + 2);
    }
    const float opac = __ldg(p.opacities + idx);
    float3 p_view = xform4x3(cam.view, mean);
    if (p_view.z > 0.2f) {
      if (p.shs) {   // start pulling this splat's SH row towards L2 while the projection math runs
        const char* row = reinterpret_cast<const char*>(p.shs + (size_t)idx * p.M * 3);
        prefetch_l2(row);
        prefetch_l2(row + 128);
      }
      float4 p_hom = xform4x4(cam.proj, mean);
      float p_w = 1.0f / (p_hom.w + 0.0000001f);
      float ndc_x = p_hom.x * p_w, ndc_y = p_hom.y * p_w;

      float c6[6];
      if (p.cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = __ldg(p.cov3D_precomp + 6 * (size_t)idx + k);
      } else {
        float r = q0, x = q1, y = q2, z = q3;
        float R[9];
        R[0] = fmaf(-2.f, fmaf(y, y, z * z), 1.f);
        R[1] = 2.f * fmaf(x, y, -(r * z));
        R[2] = 2.f * fmaf(x, z, r * y);
        R[3] = 2.f * fmaf(x, y, r * z);
        R[4] = fmaf(-2.f, fmaf(x, x, z * z), 1.f);
        R[5] = 2.f * fmaf(y, z, -(r * x));
        R[6] = 2.f * fmaf(x, z, -(r * y));
        R[7] = 2.f * fmaf(y, z, r * x);
        R[8] = fmaf(-2.f, fmaf(x, x, y * y), 1.f);
        s0 = p.scale_modifier * s0; s1 = p.scale_modifier * s1; s2 = p.scale_modifier * s2;
        float L[9];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          L[3 * i + 0] = R[3 * i + 0] * s0;
          L[3 * i + 1] = R[3 * i + 1] * s1;
          L[3 * i + 2] = R[3 * i + 2] * s2;
        }
        c6[0] = dot3(L[0], L[0], L[1], L[1], L[2], L[2]);
        c6[1] = dot3(L[0], L[3], L[1], L[4], L[2], L[5]);
        c6[2] = dot3(L[0], L[6], L[1], L[7], L[2], L[8]);
        c6[3] = dot3(L[3], L[3], L[4], L[4], L[5], L[5]);
        c6[4] = dot3(L[3], L[6], L[4], L[7], L[5], L[8]);
        c6[5] = dot3(L[6], L[6], L[7], L[7], L[8], L[8]);
      }

      // EWA: cov2D = (J Rw) Sigma (J Rw)^T + 0.3 I
      const float focal_x = (float)p.W / (2.0f * p.tan_fovx);
      const float focal_y = (float)p.H / (2.0f * p.tan_fovy);
      float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
      float txtz = p_view.x / p_view.z, tytz = p_view.y / p_view.z;
      float tx = fminf(limx, fmaxf(-limx, txtz)) * p_view.z;
      float ty = fminf(limy, fmaxf(-limy, tytz)) * p_view.z;
      float tz = p_view.z, tz2 = tz * tz;
      float J00 = focal_x / tz, J02 = -(focal_x * tx) / tz2;
      float J11 = focal_y / tz, J12 = -(focal_y * ty) / tz2;
      float m0[3], m1[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        m0[j] = fmaf(J02, cam.view[4 * j + 2], J00 * cam.view[4 * j + 0]);
        m1[j] = fmaf(J12, cam.view[4 * j + 2], J11 * cam.view[4 * j + 1]);
      }
      float v0[3], v1[3];
      v0[0] = dot3(c6[0], m0[0], c6[1], m0[1], c6[2], m0[2]);
      v0[1] = dot3(c6[1], m0[0], c6[3], m0[1], c6[4], m0[2]);
      v0[2] = dot3(c6[2], m0[0], c6[4], m0[1], c6[5], m0[2]);
      v1[0] = dot3(c6[0], m1[0], c6[1], m1[1], c6[2], m1[2]);
      v1[1] = dot3(c6[1], m1[0], c6[3], m1[1], c6[4], m1[2]);
      v1[2] = dot3(c6[2], m1[0], c6[4], m1[1], c6[5], m1[2]);
      float a = dot3(m0[0], v0[0], m0[1], v0[1], m0[2], v0[2]) + 0.3f;
      float b = dot3(m1[0], v0[0], m1[1], v0[1], m1[2], v0[2]);
      float c = dot3(m1[0], v1[0], m1[1], v1[1], m1[2], v1[2]) + 0.3f;
      float det = fmaf(a, c, -(b * b));
      if (det != 0.0f) {
        float det_inv = 1.f / det;
        float conA = c * det_inv, conB = -b * det_inv, conC = a * det_inv;
        float mid = 0.5f * (a + c);
        float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        float lambda1 = mid + sq, lambda2 = mid - sq;
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float pix = ndc2pix(ndc_x, p.W), piy = ndc2pix(ndc_y, p.H);
        const int gx = (p.W + TILE_X - 1) / TILE_X, gy = (p.H + TILE_Y - 1) / TILE_Y;
        int irad = (int)my_radius;
        float fr = (float)irad;
        int x0 = clampi((int)((pix - fr) / (float)TILE_X), 0, gx);
        int y0 = clampi((int)((piy - fr) / (float)TILE_Y), 0, gy);
        int x1 = clampi((int)((pix + fr + (float)(TILE_X - 1)) / (float)TILE_X), 0, gx);
        int y1 = clampi((int)((piy + fr + (float)(TILE_Y - 1)) / (float)TILE_Y), 0, gy);
        int area = (x1 - x0) * (y1 - y0);
        if (area != 0) {
          float rgb[3];
          if (p.colors_precomp) {
            rgb[0] = __ldg(p.colors_precomp + 3 * (size_t)idx);
            rgb[1] = __ldg(p.colors_precomp + 3 * (size_t)idx + 1);
            rgb[2] = __ldg(p.colors_precomp + 3 * (size_t)idx + 2);
          } else {
            float sh[((3 * (D + 1) * (D + 1) + 3) / 4) * 4];
            load_sh<D, VEC_SH>(p.shs, idx, p.M, sh, p.wide256 != 0);
            sh_to_rgb<D>(sh, mean, cam.campos, rgb, clamp_mask);
          }
          // Conservative half-extents (pixels) of the region where this splat can reach alpha >= 1/255:
          // alpha = o*exp(-q/2) >= 1/255  <=>  q <= tau = 2 ln(255 o); the AABB of {d^T Sigma^-1 d <= tau} is
          // sqrt(tau * Sigma_xx), sqrt(tau * Sigma_yy) with Sigma the dilated 2D covariance (a, c above).
          // +2 % and +0.5 px absorb the rounding of the fp32 conic inversion; ill-conditioned splats
          // (lambda1/lambda2 > 1e4) and NaN opacities are never culled.  The render kernels use this ONLY to
          // skip (splat, 8x4-pixel patch) pairs that provably fail the reference's own alpha test, so the
          // composited result is unchanged bit for bit.
          float hx, hy;
          if (opac != opac) { hx = hy = 3.0e38f; }
          else if (opac < 1.0f / 255.0f) { hx = hy = -1.f; }   // alpha <= o < 1/255 everywhere
          else {
            const float tau = fmaxf(0.f, 2.f * logf(255.f * opac));
            const bool ill = !(lambda2 > 0.f) || lambda1 > 1.0e4f * lambda2;
            hx = ill ? 3.0e38f : 1.02f * sqrtf(tau * a) + 0.5f;
            hy = ill ? 3.0e38f : 1.02f * sqrtf(tau * c) + 0.5f;
          }
          float4* rp = reinterpret_cast<float4*>(g.rec + idx);
          rp[0] = make_float4(pix, piy, conA, conB);
          rp[1] = make_float4(conC, opac, p_view.z, rgb[0]);
          rp[2] = make_float4(rgb[1], rgb[2], hx, hy);
          out_radius = irad;
          my_tiles = (uint32_t)area;
          key = __float_as_uint(p_view.z);
          rect_packed = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
        }
      }
    }
    radii[idx] = out_radius;
    g.tiles_touched[idx] = my_tiles;
    g.rect[idx] = rect_packed;
    g.clamped[idx] = clamp_mask;
    g.depth_key[0][idx] = key;
    g.depth_idx[0][idx] = (uint32_t)idx;
    my_key = key;
  }
  // num_rendered = sum of tiles_touched: order-independent, so one atomic per block is exact.
  // ... and counters[2] = max(~key) = ~(smallest depth key of a visible splat): lets the depth sort rebase
  // its keys so that the top digit pass degenerates to a copy for bounded scenes.
  uint32_t v = my_tiles;
  uint32_t nk = my_tiles ? ~my_key : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
    nk = max(nk, __shfl_xor_sync(0xffffffffu, nk, o));
  }
  if ((threadIdx.x & 31) == 0) { s_tiles[threadIdx.x >> 5] = v; s_nkey[threadIdx.x >> 5] = nk; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0, m = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { t += s_tiles[w]; m = max(m, s_nkey[w]); }
    if (t) { atomicAdd(g.counters, t); atomicMax(g.counters + 2, m); }
    if (p.nr_host) {
      // last block to get here publishes num_rendered straight into pinned host memory: no D2H copy node in the stream
      __threadfence();
      const uint32_t done = atomicAdd(g.counters + 1, 1u);
      if (done == gridDim.x - 1) {
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(p.nr_host) = atomicAdd(g.counters, 0u);
        __threadfence_system();
      }
    }
  }
}

template <int D, bool VEC_SH>
__global__ void __launch_bounds__(256)
preprocess_kernel(FwdParams p, GeomState g, int* __restrict__ radii) {
  preprocess_body<D, VEC_SH>(p, g, radii);
}

// (A variant of the same body under __launch_bounds__(256, 4) — 64 registers, ~50 bytes of spills, 4 instead of 3 resident
// CTAs per SM — was measured in round 2: 0.086 vs 0.087 ms, no gain, removed.)
template <int D>
static void launch_pre_d(const FwdParams& p, const GeomState& g, int* radii, cudaStream_t s) {
  int blocks = (p.P + 255) / 256;
  bool vec = p.shs && ((p.M * 3) % 4 == 0) && ((reinterpret_cast<size_t>(p.shs) & 15) == 0);
  if (vec) {
    preprocess_kernel<D, true><<<blocks, 256, 0, s>>>(p, g, radii);
  } else {
    preprocess_kernel<D, false><<<blocks, 256, 0, s>>>(p, g, radii);
  }
}

void launch_preprocess(const FwdParams& p, const GeomState& g, int* radii, cudaStream_t s) {
  switch (p.shs ? p.D : 0) {
    case 0: launch_pre_d<0>(p, g, radii, s); break;
    case 1: launch_pre_d<1>(p, g, radii, s); break;
    case 2: launch_pre_d<2>(p, g, radii, s); break;
    default: launch_pre_d<3>(p, g, radii, s); break;
  }
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  float3 m = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
  float3 pv = xform4x3(view, m);
  present[idx] = pv.z > 0.2f ? 1 : 0;
}

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s) {
  if (P > 0) mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
}

// cov3D is not part of the forward state any more (the backward recomputes it from scale / rotation): the
// export recomputes it here with the SAME -fmad=false formula sheet the preprocess kernel uses.
__device__ void cov3d_from_scale_rot(const float* scales, const float* rotations, float mod, size_t idx, float* c6) {
  float r = rotations[4 * idx], x = rotations[4 * idx + 1], y = rotations[4 * idx + 2], z = rotations[4 * idx + 3];
  float R[9];
  R[0] = fmaf(-2.f, fmaf(y, y, z * z), 1.f);
  R[1] = 2.f * fmaf(x, y, -(r * z));
  R[2] = 2.f * fmaf(x, z, r * y);
  R[3] = 2.f * fmaf(x, y, r * z);
  R[4] = fmaf(-2.f, fmaf(x, x, z * z), 1.f);
  R[5] = 2.f * fmaf(y, z, -(r * x));
  R[6] = 2.f * fmaf(x, z, -(r * y));
  R[7] = 2.f * fmaf(y, z, r * x);
  R[8] = fmaf(-2.f, fmaf(x, x, y * y), 1.f);
  float s0 = mod * scales[3 * idx], s1 = mod * scales[3 * idx + 1], s2 = mod * scales[3 * idx + 2];
  float L[9];
  for (int i = 0; i < 3; i++) { L[3 * i] = R[3 * i] * s0; L[3 * i + 1] = R[3 * i + 1] * s1; L[3 * i + 2] = R[3 * i + 2] * s2; }
  c6[0] = dot3(L[0], L[0], L[1], L[1], L[2], L[2]);
  c6[1] = dot3(L[0], L[3], L[1], L[4], L[2], L[5]);
  c6[2] = dot3(L[0], L[6], L[1], L[7], L[2], L[8]);
  c6[3] = dot3(L[3], L[3], L[4], L[4], L[5], L[5]);
  c6[4] = dot3(L[3], L[6], L[4], L[7], L[5], L[8]);
  c6[5] = dot3(L[6], L[6], L[7], L[7], L[8], L[8]);
}

__global__ void export_geom_kernel(int P, GeomState g, const float* scales, const float* rotations, float mod,
                                   const float* cov3D_precomp, float* means2D, float* depths, float* cov3D,
                                   float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  bool vis = g.tiles_touched[idx] > 0;
  SplatRec r;
  if (vis) r = g.rec[idx];
  if (means2D) { means2D[2 * idx] = vis ? r.x : 0.f; means2D[2 * idx + 1] = vis ? r.y : 0.f; }
  if (depths) depths[idx] = vis ? r.depth : 0.f;
  if (conic_opacity) {
    conic_opacity[4 * idx + 0] = vis ? r.conA : 0.f; conic_opacity[4 * idx + 1] = vis ? r.conB : 0.f;
    conic_opacity[4 * idx + 2] = vis ? r.conC : 0.f; conic_opacity[4 * idx + 3] = vis ? r.opacity : 0.f;
  }
  if (rgb) { rgb[3 * idx] = vis ? r.r : 0.f; rgb[3 * idx + 1] = vis ? r.g : 0.f; rgb[3 * idx + 2] = vis ? r.b : 0.f; }
  if (cov3D) {
    float c6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (vis) {
      if (cov3D_precomp) for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[6 * (size_t)idx + k];
      else if (scales && rotations) cov3d_from_scale_rot(scales, rotations, mod, (size_t)idx, c6);
    }
    for (int k = 0; k < 6; k++) cov3D[6 * idx + k] = c6[k];
  }
  if (clamped) {
    uint8_t m = vis ? g.clamped[idx] : 0;
    clamped[3 * idx] = m & 1; clamped[3 * idx + 1] = (m >> 1) & 1; clamped[3 * idx + 2] = (m >> 2) & 1;
  }
  if (tiles_touched) tiles_touched[idx] = g.tiles_touched[idx];
}

void launch_export_geom(int P, const GeomState& g, const float* scales, const float* rotations, float mod,
                        const float* cov3D_precomp, float* means2D, float* depths, float* cov3D,
                        float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched,
                        cudaStream_t s) {
  if (P > 0)
    export_geom_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g, scales, rotations, mod, cov3D_precomp, means2D, depths,
                                                       cov3D, conic_opacity, rgb, clamped, tiles_touched);
}

}  // namespace sfb
