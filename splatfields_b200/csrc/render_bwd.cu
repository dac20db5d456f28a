// render_bwd.cu — K7: per-pixel back-to-front gradient of the compositing, scattered onto the splats.
// Restates the external rasterizer's backward render: starting from
// final_T and the last contributor, walk the tile list backwards, rebuild alpha and T, and emit
//   d/d rgb_i, d/d opacity_i, d/d conic_i (A,B,C), d/d mean2D_i (NDC-scaled: includes 0.5*W, 0.5*H).
//
// The reference issues 9 global float atomics per (pixel, contributor).  Here the 32 pixels of a warp
// are first summed with a recursive-halving shuffle reduction (14 shuffles for 9 values, and the 8
// results end up on 8 different lanes), those lanes add into a per-tile shared-memory accumulator, and
// only one vectorised global reduction (2x red.global.add.v4.f32 + 1 scalar) is issued per
// (tile, splat) instance: global atomic traffic drops from 9 * pixels * contributors to 3 * R.
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
template <bool CULL>
__global__ void __launch_bounds__(BB)
render_backward_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                       const uint32_t* __restrict__ point_list, const SplatRec* __restrict__ rec,
                       const float* __restrict__ bg, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                       GradRec* __restrict__ grad) {
  __shared__ float4 s_q0[BB];
  __shared__ float4 s_q1[BB];
  __shared__ float2 s_q2[BB];
  __shared__ uint32_t s_id[BB];
  __shared__ float s_acc[BB * 9];
  __shared__ uint32_t s_max[BB / 32];
  __shared__ uint8_t s_mask[CULL ? BB : 1];
  __shared__ uint8_t s_list[CULL ? BB / 32 : 1][CULL ? BB : 1];
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
  const float bg_dot = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
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
      uint32_t id = point_list[range.x + (uint32_t)(top - 1 - (int)threadIdx.x)];
      const float4* rp = reinterpret_cast<const float4*>(rec + id);
      float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
      s_q0[threadIdx.x] = a;
      s_q1[threadIdx.x] = b;
      s_q2[threadIdx.x] = make_float2(c.x, c.y);
      s_id[threadIdx.x] = id;
      if (CULL) mask = patch_mask(a.x, a.y, c.z, c.w, tx0, ty0);
    }
    if (CULL) s_mask[threadIdx.x] = (uint8_t)mask;
#pragma unroll
    for (int k = 0; k < 9; k++) s_acc[k * BB + threadIdx.x] = 0.f;
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
      float v[8], v8 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = 0.f;
      bool contrib = false;
      if (pos < my_last) {
        const float4 q0 = s_q0[j];
        const float4 q1 = s_q1[j];
        const float dx = q0.x - pixfx, dy = q0.y - pixfy;
        const float s = __fmaf_rn(__fmul_rn(q0.z, dx), dx, __fmul_rn(__fmul_rn(q1.x, dy), dy));
        const float power = __fmaf_rn(s, -0.5f, -__fmul_rn(__fmul_rn(q0.w, dx), dy));
        if (power <= 0.0f) {
          const float G = expf(power);
          const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
          if (alpha >= 1.0f / 255.0f) {
            contrib = true;
            const float2 q2 = s_q2[j];
            T = T / (1.f - alpha);
            const float dchannel_dcolor = alpha * T;
            acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
            acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
            acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
            lc0 = q1.w; lc1 = q2.x; lc2 = q2.y;
            float dL_dalpha = (lc0 - acc0) * dLp0 + (lc1 - acc1) * dLp1 + (lc2 - acc2) * dLp2;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            const float dL_dG = q1.y * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
            const float dG_ddely = -gdy * q1.x - gdx * q0.w;
            v[0] = dL_dG * dG_ddelx * ddelx_dx;
            v[1] = dL_dG * dG_ddely * ddely_dy;
            v[2] = -0.5f * gdx * dx * dL_dG;
            v[3] = -gdx * dy * dL_dG;
            v[4] = -0.5f * gdy * dy * dL_dG;
            v[5] = G * dL_dalpha;
            v[6] = dchannel_dcolor * dLp0;
            v[7] = dchannel_dcolor * dLp1;
            v8 = dchannel_dcolor * dLp2;
          }
        }
      }
      if (!__any_sync(0xffffffffu, contrib)) continue;
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
      if ((lane & 3) == 0) {
        const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        atomicAdd(&s_acc[k * BB + j], u);
      }
      if (lane == 1) atomicAdd(&s_acc[8 * BB + j], v8);
    }
    __syncthreads();
    if ((int)threadIdx.x < n) {
      float a[9];
      bool nz = false;
#pragma unroll
      for (int k = 0; k < 9; k++) { a[k] = s_acc[k * BB + threadIdx.x]; nz |= (a[k] != 0.f); }
      if (nz) {
        float* gp = reinterpret_cast<float*>(grad + s_id[threadIdx.x]);
        red_add_v4(gp, a[0], a[1], a[2], a[3]);
        red_add_v4(gp + 4, a[4], a[5], a[6], a[7]);
        atomicAdd(gp + 8, a[8]);
      }
    }
  }
}

void launch_render_backward(int W, int H, const uint2* ranges, const uint32_t* point_list, const SplatRec* rec,
                            const float* bg, const float* final_T, const uint32_t* n_contrib,
                            const float* dL_dpixels, GradRec* grad, cudaStream_t s) {
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  static int cull = -1;
  if (cull < 0) { const char* e = getenv("SFB_NO_CULL"); cull = (e && e[0] == '1') ? 0 : 1; }
  if (cull)
    render_backward_kernel<true><<<gx * gy, BB, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib,
                                                         dL_dpixels, grad);
  else
    render_backward_kernel<false><<<gx * gy, BB, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T,
                                                          n_contrib, dL_dpixels, grad);
}

}  // namespace sfb
