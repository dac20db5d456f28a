// render_bwd.cu — K7: per-pixel back-to-front gradient of the compositing, scattered onto the splats.
// Restates the external rasterizer's backward render: starting from
// final_T and the last contributor, walk the tile list backwards, rebuild alpha and T, and emit
//   d/d rgb_i, d/d opacity_i, d/d conic_i (A,B,C), d/d mean2D_i (NDC-scaled: includes 0.5*W, 0.5*H).
//
// The reference issues 9 global float atomics per (pixel, contributor).  Here each pixel only forms the
// raw moments of s = dL/dG*G about the splat centre plus the three colour terms; the 32 pixels of a warp
// are summed with a recursive-halving shuffle reduction (14 shuffles for 9 values, the results land on 9
// different lanes), those lanes add into a per-tile shared-memory accumulator in one conflict-free
// atomic, and at the end of a batch one thread per splat turns the moments into gradients and issues one
// vectorised global reduction (2x red.global.add.v4.f32 + 1 scalar) per (tile, splat) instance: global
// atomic traffic drops from 9 * pixels * contributors to 3 * R.
// Accumulation order differs from the reference's (as it does between two runs of the reference), so
// parity here is tolerance-based: 1e-3 relative on every per-splat gradient.
#include "common.cuh"
#include <cstdlib>

namespace sfb {

constexpr int BB = 256;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// CULL: same per-warp footprint culling as the forward render (SplatRec::hx/hy): a warp skips splats that
// cannot reach alpha >= 1/255 on any of its 32 pixels — pairs whose contribution is exactly zero.
// ALPHA: the cotangent of the fused coverage image (see render_fwd.cu) enters as a fourth channel with
// colour 1 and background 0: it only adds to dL/dalpha, exactly the sum the reference gets from the backward
// of its second (alpha) pass.
template <bool CULL, bool ALPHA>
__global__ void __launch_bounds__(BB)
render_backward_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                       const uint32_t* __restrict__ point_list, uint32_t idx_mask,
                      const SplatRec* __restrict__ rec,
                       const float* __restrict__ bg, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                       const float* __restrict__ dL_dalpha_img, const uint8_t* __restrict__ hit,
                       GradRec* __restrict__ grad) {
  // one struct = one base register: every access below is base + immediate (+ j * stride)
  struct Smem {
    float4 q0[BB];
    float4 q1[BB];
    float2 q2[BB];
    uint32_t id[BB];
    float acc[BB * 9];
    uint32_t maxc[BB / 32];
    uint8_t mask[CULL ? BB : 1];
    uint8_t list[CULL ? BB / 32 : 1][CULL ? BB : 1];
  };
  __shared__ Smem sm;
  float4* const s_q0 = sm.q0;
  float4* const s_q1 = sm.q1;
  float2* const s_q2 = sm.q2;
  uint32_t* const s_id = sm.id;
  float* const s_acc = sm.acc;
  uint32_t* const s_max = sm.maxc;
  uint8_t* const s_mask = sm.mask;
  uint8_t (*const s_list)[CULL ? BB : 1] = sm.list;
  const float tx0 = (float)((blockIdx.x % grid_x) * TILE_X), ty0 = (float)((blockIdx.x / grid_x) * TILE_Y);

  const int tile = blockIdx.x;
  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = tile_x * TILE_X + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pixfx = (float)px, pixfy = (float)py;
  const size_t pix = (size_t)py * W + px;
  const size_t HW = (size_t)H * W;

  const uint2 range = ranges[tile];
  const float T_final = inside ? final_T[pix] : 0.f;
  const uint32_t my_last = inside ? n_contrib[pix] : 0u;
  float T = T_final;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
  if (inside) { dLp0 = dL_dpixels[pix]; dLp1 = dL_dpixels[HW + pix]; dLp2 = dL_dpixels[2 * HW + pix]; }
  float dLpa = 0.f, acca = 0.f;          // alpha channel: cotangent, accumulated "colour" (= 1) behind
  if (ALPHA && inside) dLpa = dL_dalpha_img[pix];
  const float bg_dot = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
  const float Tf_bg = T_final * bg_dot;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

  // nothing behind the deepest last-contributor of the tile matters to any pixel
  uint32_t m = __reduce_max_sync(0xffffffffu, my_last);
  if (lane == 0) s_max[warp] = m;
  __syncthreads();
  uint32_t hi = 0;
#pragma unroll
  for (int w = 0; w < BB / 32; w++) hi = max(hi, s_max[w]);

  for (int top = (int)hi; top > 0; top -= BB) {
    // smem slot j holds list position top-1-j (back to front)
    const int n = top < BB ? top : BB;
    __syncthreads();
    uint32_t mask = 0u;
    if ((int)threadIdx.x < n) {
      uint32_t id = point_list[range.x + (uint32_t)(top - 1 - (int)threadIdx.x)] & idx_mask;
      const float4* rp = reinterpret_cast<const float4*>(rec + id);
      float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
      s_q0[threadIdx.x] = a;
      s_q1[threadIdx.x] = b;
      s_q2[threadIdx.x] = make_float2(c.x, c.y);
      s_id[threadIdx.x] = id;
      // the forward recorded which warps accumulated this entry: exact, and cheaper than the footprint box
      if (CULL) mask = hit ? (uint32_t)hit[range.x + (uint32_t)(top - 1 - (int)threadIdx.x)]
                           : refine_patch_mask(patch_mask(a.x, a.y, c.z, c.w, tx0, ty0), a.x, a.y, a.z, a.w, b.x, b.y, c.z,
                                               tx0, ty0);
    }
    if (CULL) s_mask[threadIdx.x] = (uint8_t)mask;
#pragma unroll
    for (int k = 0; k < 9; k++) s_acc[k * BB + threadIdx.x] = 0.f;   // 9*BB floats, any order
    __syncthreads();
    int nsweep = n;
    if (CULL) {
      int cnt = 0;
      const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
      for (int c8 = 0; c8 < BB / 32; c8++) {
        const int idx = c8 * 32 + lane;
        const bool hit = (s_mask[idx] >> warp) & 1;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (hit) s_list[warp][cnt + __popc(bal & lt)] = (uint8_t)idx;
        cnt += __popc(bal);
      }
      __syncwarp();
      nsweep = cnt;
    }

    for (int kk = 0; kk < nsweep; kk++) {
      const int j = CULL ? (int)s_list[warp][kk] : kk;
      const uint32_t pos = (uint32_t)(top - 1 - j);
      // Per-lane work is kept to the raw moments of  s = dL/dG * G  about the splat centre
      //   S0 = s, Sx = s dx, Sy = s dy, Sxx = s dx^2, Sxy = s dx dy, Syy = s dy^2   and   w dL/dC_c,
      // everything that is per-splat (conic, opacity, 0.5 W / 0.5 H) is applied once at flush time.
      bool contrib = false;
      float dx = 0.f, dy = 0.f, sG = 0.f, wgt = 0.f;
      if (pos < my_last) {
        const float4 q0 = s_q0[j];
        const float4 q1 = s_q1[j];
        dx = q0.x - pixfx; dy = q0.y - pixfy;
        const float s = __fmaf_rn(__fmul_rn(q0.z, dx), dx, __fmul_rn(__fmul_rn(q1.x, dy), dy));
        const float power = __fmaf_rn(s, -0.5f, -__fmul_rn(__fmul_rn(q0.w, dx), dy));
        if (power <= 0.0f) {
          const float G = expf(power);
          const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
          if (alpha >= 1.0f / 255.0f) {
            contrib = true;
            const float2 q2 = s_q2[j];
            float inv;   // 1 - alpha is in [0.01, 1]: the bare approximate reciprocal (1 ulp) is safe
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - alpha));
            T *= inv;
            wgt = alpha * T;
            acc0 = fmaf(last_alpha, lc0 - acc0, acc0);
            acc1 = fmaf(last_alpha, lc1 - acc1, acc1);
            acc2 = fmaf(last_alpha, lc2 - acc2, acc2);
            lc0 = q1.w; lc1 = q2.x; lc2 = q2.y;
            float dL_dalpha = (lc0 - acc0) * dLp0;
            dL_dalpha = fmaf(lc1 - acc1, dLp1, dL_dalpha);
            dL_dalpha = fmaf(lc2 - acc2, dLp2, dL_dalpha);
            if (ALPHA) {
              acca = fmaf(last_alpha, 1.f - acca, acca);   // every splat's "colour" is 1 (last_alpha = 0 at the first)
              dL_dalpha = fmaf(1.f - acca, dLpa, dL_dalpha);
            }
            dL_dalpha = fmaf(dL_dalpha, T, -Tf_bg * inv);
            last_alpha = alpha;
            sG = q1.y * dL_dalpha * G;
          }
        }
      }
      if (!__any_sync(0xffffffffu, contrib)) continue;
      float v[8], v8;
      {
        const float sx = sG * dx, sy = sG * dy;     // sG == 0 on lanes that do not contribute
        v[0] = sG; v[1] = sx; v[2] = sy; v[3] = sx * dx; v[4] = sx * dy; v[5] = sy * dy;
        v[6] = wgt * dLp0; v[7] = wgt * dLp1; v8 = wgt * dLp2;
      }
      // recursive halving: 8 values over 32 lanes in 4+2+1+2 shuffles
      float w4[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool up = lane & 16;
        float send = up ? v[k] : v[k + 4];
        float keep = up ? v[k + 4] : v[k];
        w4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      float w2[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const bool up = lane & 8;
        float send = up ? w4[k] : w4[k + 2];
        float keep = up ? w4[k + 2] : w4[k];
        w2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      float u;
      {
        const bool up = lane & 4;
        float send = up ? w2[0] : w2[1];
        float keep = up ? w2[1] : w2[0];
        u = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      u += __shfl_xor_sync(0xffffffffu, u, 2);
      u += __shfl_xor_sync(0xffffffffu, u, 1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(0xffffffffu, v8, o);
      // lanes 0,4,..,28 hold values 0..7, lane 1 holds value 8: ONE atomic site, nine distinct banks
      if ((lane & 3) == 0 || lane == 1) {
        const int k = lane == 1 ? 8 : ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        atomicAdd(&s_acc[j * 9 + k], lane == 1 ? v8 : u);
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < n) {
      float a[9];
      bool nz = false;
#pragma unroll
      for (int k = 0; k < 9; k++) { a[k] = s_acc[threadIdx.x * 9 + k]; nz |= (a[k] != 0.f); }
      if (nz) {
        const float4 q0 = s_q0[threadIdx.x];
        const float4 q1 = s_q1[threadIdx.x];
        const float conA = q0.z, conB = q0.w, conC = q1.x, op = q1.y;
        // moments -> gradients: d/dmean (NDC-scaled), d/dconic (true derivatives), d/dopacity
        const float gx = -(conA * a[1] + conB * a[2]) * ddelx_dx;
        const float gy = -(conC * a[2] + conB * a[1]) * ddely_dy;
        const float gA = -0.5f * a[3], gB = -a[4], gC = -0.5f * a[5];
        const float gop = a[0] / op;      // sum of G * dL/dalpha  (op >= 1/255 whenever a[0] != 0)
        float* gp = reinterpret_cast<float*>(grad + s_id[threadIdx.x]);
        red_add_v4(gp, gx, gy, gA, gB);
        red_add_v4(gp + 4, gC, a[0] != 0.f ? gop : 0.f, a[6], a[7]);
        atomicAdd(gp + 8, a[8]);
      }
    }
  }
}

void launch_render_backward(int W, int H, const uint2* ranges, const uint32_t* point_list, uint32_t idx_mask,
                            const SplatRec* rec,
                            const float* bg, const float* final_T, const uint32_t* n_contrib,
                            const float* dL_dpixels, const float* dL_dalpha_img, const uint8_t* hit, GradRec* grad,
                            cudaStream_t s) {
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  static int cull = -1;
  if (cull < 0) { const char* e = getenv("SFB_NO_CULL"); cull = (e && e[0] == '1') ? 0 : 1; }
#define SFB_RB(C, A)                                                                                          \
  render_backward_kernel<C, A><<<gx * gy, BB, 0, s>>>(W, H, gx, ranges, point_list, idx_mask, rec, bg, final_T, n_contrib, \
                                                      dL_dpixels, dL_dalpha_img, hit, grad)
  if (cull) { if (dL_dalpha_img) SFB_RB(true, true); else SFB_RB(true, false); }
  else      { if (dL_dalpha_img) SFB_RB(false, true); else SFB_RB(false, false); }
#undef SFB_RB
}

}  // namespace sfb
