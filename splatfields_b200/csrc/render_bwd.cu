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
          const float G = splat_exp(power);
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

// ---------------------------------------------------------------------------------------------------------
// Tensor-core variant of the pixel -> splat reduction.
//
// For one list entry the nine sums over the 256 pixels of the tile are
//     m_k = sum_p sG_p * phi_k(X_p, Y_p),  phi = (1, X, Y, X^2, XY, Y^2)      (raw moments of s = dL/dG * G)
//     c_k = sum_p w_p  * dL/dC_k(p),       k = 0..2                           (w = alpha * T)
// with (X, Y) the pixel position relative to the tile CENTRE: a matrix product  [entries x pixels] * [pixels x 9]
// whose right-hand side does not depend on the entry.  Each warp therefore stages (sG, w) of 16 consecutive
// entries of its list in shared memory ([entry][pixel], 4 KB per warp) and multiplies with
// mma.sync.m16n8k8 (tf32 inputs, fp32 accumulate): 4 k-steps of 8 pixels cover its 32 pixels.
//   * phi holds half-integers |X|,|Y| <= 7.5 and their products (<= 56.25): exact in tf32.  sG is split into
//     hi = top 19 bits, lo = sG - hi (exact), two MMAs -> ~2^-20 relative, fp32-grade sums.
//   * dL/dC is arbitrary: split both operands, three MMAs (hi*hi + hi*lo + lo*hi).
// The shuffle tree this replaces cost ~65 of the ~140 instructions of an iteration; this costs ~12 per entry
// (1 STS, 1.5 LDS, 4 split ops, 1.25 MMA, 0.4 shared atomics).  The moments come out about the tile centre
// and are moved to the splat centre at flush time (dx = cx - X):  Sx = cx m0 - mX,  Sxx = cx^2 m0 - 2 cx mX + mXX, ...
// tcgen05 is the wrong tool here: its smallest tile (M = 64/128 rows from shared memory, one issuing thread,
// TMEM round trip) cannot follow eight ragged per-warp lists; the warp-level MMA consumes exactly the
// (warp, entry) pairs the hit masks select.
constexpr int MG = 16;   // entries per MMA group = M of m16n8k8

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// hi = the 19 bits the tensor core reads (truncation), lo = the exact remainder
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

struct SmemBwdMma {
  float4 q0[BB];
  float4 q1[BB];
  float4 q2[BB];            // g, b, -, - (16-byte stride like q0 / q1: one address register + immediates)
  uint32_t id[BB];
  float acc[BB * 9];
  uint32_t maxc[BB / 32];
  uint8_t mask[BB];
  uint8_t list[BB / 32][BB];
  float2 stage[BB / 32][MG * 32];   // per warp: [entry row][pixel ^ swizzle] = (sG, w)
  float2 dlp[BB / 32][4][32];       // per warp: (hi, lo) of dL/dC_c per pixel; channel 3 = zeros
};

template <bool ALPHA>
__global__ void __launch_bounds__(BB, 3)
render_backward_mma_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                           const uint32_t* __restrict__ point_list, uint32_t idx_mask,
                           const SplatRec* __restrict__ rec,
                           const float* __restrict__ bg, const float* __restrict__ final_T,
                           const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                           const float* __restrict__ dL_dalpha_img, const uint8_t* __restrict__ hit,
                           GradRec* __restrict__ grad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemBwdMma& sm = *reinterpret_cast<SmemBwdMma*>(smem_raw);

  const int tile = blockIdx.x;
  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const float tx0 = (float)(tile_x * TILE_X), ty0 = (float)(tile_y * TILE_Y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;          // MMA fragment coordinates
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
  float dLpa = 0.f, acca = 0.f;
  if (ALPHA && inside) dLpa = dL_dalpha_img[pix];
  const float bg_dot = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
  const float Tf_bg = T_final * bg_dot;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

  // B operand of the colour product: this warp's dL/dC per pixel (row k = lane), split for tf32
  {
    uint32_t h, l;
    split_tf32(dLp0, h, l); sm.dlp[warp][0][lane] = make_float2(__uint_as_float(h), __uint_as_float(l));
    split_tf32(dLp1, h, l); sm.dlp[warp][1][lane] = make_float2(__uint_as_float(h), __uint_as_float(l));
    split_tf32(dLp2, h, l); sm.dlp[warp][2][lane] = make_float2(__uint_as_float(h), __uint_as_float(l));
    sm.dlp[warp][3][lane] = make_float2(0.f, 0.f);
  }
  // B operand of the moment product: phi_n at pixel k, n = gid, k = ks*8 + tig (+4); constant per lane.
  // Output column n lands in accumulator slot n (0..5); the colour sums use columns 6, 7, 5 -> slots 6, 7, 8.
  uint32_t bS0[4], bS1[4];
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    const float Y = (float)((warp >> 1) * 4 + ks) - 7.5f;
    const float Xa = (float)((warp & 1) * 8 + tig) - 7.5f, Xb = Xa + 4.0f;
    // branch-free on purpose: a switch here compiles to an indirect branch (BRX) and the lanes of one case
    // were observed to stay split off for the rest of the kernel, which breaks the warp-wide MMAs below
    const float k1 = gid == 0 ? 1.f : 0.f, kx = gid == 1 ? 1.f : 0.f, ky = gid == 2 ? 1.f : 0.f;
    const float kxx = gid == 3 ? 1.f : 0.f, kxy = gid == 4 ? 1.f : 0.f, kyy = gid == 5 ? 1.f : 0.f;
    const float fa = k1 + kx * Xa + ky * Y + kxx * (Xa * Xa) + kxy * (Xa * Y) + kyy * (Y * Y);
    const float fb = k1 + kx * Xb + ky * Y + kxx * (Xb * Xb) + kxy * (Xb * Y) + kyy * (Y * Y);
    bS0[ks] = __float_as_uint(fa);
    bS1[ks] = __float_as_uint(fb);
  }
  // colour channel of this lane's B column: 6 -> 0, 7 -> 1, 5 -> 2, else 3 (zeros); arithmetic, not a jump table
  const int wch = 3 - 3 * (int)(gid == 6) - 2 * (int)(gid == 7) - (int)(gid == 5);
  const float2* const my_dlp = sm.dlp[warp][wch];
  float2* const st = sm.stage[warp];
  const int swz = (gid & 3) << 2;                       // fragment loads: rows gid and gid+8 share (row & 3)

  uint32_t m = __reduce_max_sync(0xffffffffu, my_last);
  if (lane == 0) sm.maxc[warp] = m;
  __syncthreads();
  uint32_t hi = 0;
#pragma unroll
  for (int w = 0; w < BB / 32; w++) hi = max(hi, sm.maxc[w]);

  // mma.sync needs the whole warp converged; fail loudly rather than reduce garbage if it ever is not
  if (__activemask() != 0xffffffffu) __trap();

  for (int top = (int)hi; top > 0; top -= BB) {
    const int n = top < BB ? top : BB;
    __syncthreads();
    uint32_t mask = 0u;
    if ((int)threadIdx.x < n) {
      uint32_t id = point_list[range.x + (uint32_t)(top - 1 - (int)threadIdx.x)] & idx_mask;
      const float4* rp = reinterpret_cast<const float4*>(rec + id);
      float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
      sm.q0[threadIdx.x] = a;
      sm.q1[threadIdx.x] = b;
      sm.q2[threadIdx.x] = c;
      sm.id[threadIdx.x] = id;
      mask = hit ? (uint32_t)hit[range.x + (uint32_t)(top - 1 - (int)threadIdx.x)]
                 : refine_patch_mask(patch_mask(a.x, a.y, c.z, c.w, tx0, ty0), a.x, a.y, a.z, a.w, b.x, b.y, c.z,
                                     tx0, ty0);
    }
    sm.mask[threadIdx.x] = (uint8_t)mask;
#pragma unroll
    for (int k = 0; k < 9; k++) sm.acc[k * BB + threadIdx.x] = 0.f;
    __syncthreads();
    int nsweep = 0;
    {
      const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
      for (int c8 = 0; c8 < BB / 32; c8++) {
        const int idx = c8 * 32 + lane;
        const bool h = (sm.mask[idx] >> warp) & 1;
        const uint32_t bal = __ballot_sync(0xffffffffu, h);
        if (h) sm.list[warp][nsweep + __popc(bal & lt)] = (uint8_t)idx;
        nsweep += __popc(bal);
      }
      __syncwarp();
    }

    // list position of slot j is top-1-j; it is in front of this pixel's last contributor iff j >= top - my_last
    const int jmin = top - (int)my_last;
    for (int g0 = 0; g0 < nsweep; g0 += MG) {
      const int gn = nsweep - g0 < MG ? nsweep - g0 : MG;
      // ---- phase A: per-pixel chain over up to 16 entries; (sG, w) of every pixel go to the staging rows
      auto entry = [&](const int i, const int j) {
        float sG = 0.f, wgt = 0.f;
        if (j >= jmin) {
          const float4 q0 = sm.q0[j];
          const float4 q1 = sm.q1[j];
          const float dx = q0.x - pixfx, dy = q0.y - pixfy;
          const float s = __fmaf_rn(__fmul_rn(q0.z, dx), dx, __fmul_rn(__fmul_rn(q1.x, dy), dy));
          const float power = __fmaf_rn(s, -0.5f, -__fmul_rn(__fmul_rn(q0.w, dx), dy));
          if (power <= 0.0f) {
            const float G = splat_exp(power);
            const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
            if (alpha >= 1.0f / 255.0f) {
              const float2 q2 = make_float2(sm.q2[j].x, sm.q2[j].y);
              float inv;
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - alpha));
              T *= inv;
              wgt = alpha * T;
              // colour accumulated BEHIND this splat: acc <- alpha c + (1 - alpha) acc, applied after its use below
              // (the reference carries last_alpha / last_color to the next iteration; same values, fewer registers)
              const float d0 = q1.w - acc0, d1 = q2.x - acc1, d2 = q2.y - acc2;
              float dL_dalpha = d0 * dLp0;
              dL_dalpha = fmaf(d1, dLp1, dL_dalpha);
              dL_dalpha = fmaf(d2, dLp2, dL_dalpha);
              acc0 = fmaf(alpha, d0, acc0);
              acc1 = fmaf(alpha, d1, acc1);
              acc2 = fmaf(alpha, d2, acc2);
              if (ALPHA) {
                const float da = 1.f - acca;
                dL_dalpha = fmaf(da, dLpa, dL_dalpha);
                acca = fmaf(alpha, da, acca);
              }
              dL_dalpha = fmaf(dL_dalpha, T, -Tf_bg * inv);
              sG = q1.y * dL_dalpha * G;
            }
          }
        }
        st[i * 32 + (lane ^ ((i & 3) << 2))] = make_float2(sG, wgt);
      };
      if (gn == MG) {
        // full group (the common case): straight-line code, the 16 list bytes come with one 16-byte load and every
        // staging address is base + immediate
        const uint4 lw = *reinterpret_cast<const uint4*>(&sm.list[warp][g0]);
        const uint32_t lws[4] = {lw.x, lw.y, lw.z, lw.w};
#pragma unroll
        for (int i = 0; i < MG; i++) entry(i, (int)((lws[i >> 2] >> (8 * (i & 3))) & 0xffu));
      } else {
        for (int i = 0; i < gn; i++) entry(i, (int)sm.list[warp][g0 + i]);
      }
      for (int i = gn; i < MG; i++) st[i * 32 + lane] = make_float2(0.f, 0.f);   // rows without an entry (last group)
      __syncwarp();
      // ---- phase B: [16 entries x 32 pixels] x [32 pixels x 8] on the tensor cores
      float dS[4] = {0.f, 0.f, 0.f, 0.f}, dW[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const int p0 = (ks * 8 + tig) ^ swz, p1 = (ks * 8 + tig + 4) ^ swz;
        const float2 v0 = st[gid * 32 + p0], v1 = st[(gid + 8) * 32 + p0];
        const float2 v2 = st[gid * 32 + p1], v3 = st[(gid + 8) * 32 + p1];
        uint32_t sh0, sl0, sh1, sl1, sh2, sl2, sh3, sl3, wh0, wl0, wh1, wl1, wh2, wl2, wh3, wl3;
        split_tf32(v0.x, sh0, sl0); split_tf32(v1.x, sh1, sl1); split_tf32(v2.x, sh2, sl2); split_tf32(v3.x, sh3, sl3);
        split_tf32(v0.y, wh0, wl0); split_tf32(v1.y, wh1, wl1); split_tf32(v2.y, wh2, wl2); split_tf32(v3.y, wh3, wl3);
        mma_tf32(dS, sh0, sh1, sh2, sh3, bS0[ks], bS1[ks]);
        mma_tf32(dS, sl0, sl1, sl2, sl3, bS0[ks], bS1[ks]);
        const float2 d0 = my_dlp[ks * 8 + tig], d1 = my_dlp[ks * 8 + tig + 4];
        mma_tf32(dW, wh0, wh1, wh2, wh3, __float_as_uint(d0.x), __float_as_uint(d1.x));
        mma_tf32(dW, wh0, wh1, wh2, wh3, __float_as_uint(d0.y), __float_as_uint(d1.y));
        mma_tf32(dW, wl0, wl1, wl2, wl3, __float_as_uint(d0.x), __float_as_uint(d1.x));
      }
      // accumulator rows gid / gid+8 = entries g0+gid / g0+gid+8; columns 2*tig, 2*tig+1.
      // moments: slots 0..5 (tig 0..2); colours: columns 6,7 -> slots 6,7 (tig 3), column 5 -> slot 8 (tig 2).
      {
        const bool mom = tig < 3;
        if (gid < gn) {
          float* a = &sm.acc[(int)sm.list[warp][g0 + gid] * 9];
          atomicAdd(a + 2 * tig, mom ? dS[0] : dW[0]);
          atomicAdd(a + 2 * tig + 1, mom ? dS[1] : dW[1]);
          if (tig == 2) atomicAdd(a + 8, dW[1]);
        }
        if (gid + 8 < gn) {
          float* a = &sm.acc[(int)sm.list[warp][g0 + gid + 8] * 9];
          atomicAdd(a + 2 * tig, mom ? dS[2] : dW[2]);
          atomicAdd(a + 2 * tig + 1, mom ? dS[3] : dW[3]);
          if (tig == 2) atomicAdd(a + 8, dW[3]);
        }
      }
      __syncwarp();   // staging rows are rewritten by the next group
    }
    __syncthreads();
    if ((int)threadIdx.x < n) {
      float a[9];
      bool nz = false;
#pragma unroll
      for (int k = 0; k < 9; k++) { a[k] = sm.acc[threadIdx.x * 9 + k]; nz |= (a[k] != 0.f); }
      if (nz) {
        const float4 q0 = sm.q0[threadIdx.x];
        const float4 q1 = sm.q1[threadIdx.x];
        const float conA = q0.z, conB = q0.w, conC = q1.x, op = q1.y;
        // tile-centre moments -> splat-centre moments: dx = cx - X, dy = cy - Y
        const float cx = q0.x - (tx0 + 7.5f), cy = q0.y - (ty0 + 7.5f);
        const float S0 = a[0];
        const float Sx = fmaf(cx, S0, -a[1]);
        const float Sy = fmaf(cy, S0, -a[2]);
        const float Sxx = fmaf(cx, fmaf(cx, S0, -2.f * a[1]), a[3]);
        const float Sxy = fmaf(cx, fmaf(cy, S0, -a[2]), fmaf(-cy, a[1], a[4]));
        const float Syy = fmaf(cy, fmaf(cy, S0, -2.f * a[2]), a[5]);
        const float gx = -(conA * Sx + conB * Sy) * ddelx_dx;
        const float gy = -(conC * Sy + conB * Sx) * ddely_dy;
        const float gA = -0.5f * Sxx, gB = -Sxy, gC = -0.5f * Syy;
        const float gop = S0 / op;
        float* gp = reinterpret_cast<float*>(grad + sm.id[threadIdx.x]);
        red_add_v4(gp, gx, gy, gA, gB);
        red_add_v4(gp + 4, gC, S0 != 0.f ? gop : 0.f, a[6], a[7]);
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
  static int use_mma = -1;   // SFB_BWD_SHFL=1: shuffle-tree reduction instead of the tensor-core one (A/B knob)
  if (use_mma < 0) {
    const char* e = getenv("SFB_BWD_SHFL");
    use_mma = (e && e[0] == '1') ? 0 : 1;
    cudaFuncSetAttribute(render_backward_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(SmemBwdMma));
    cudaFuncSetAttribute(render_backward_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(SmemBwdMma));
  }
  if (cull && use_mma) {
    if (dL_dalpha_img)
      render_backward_mma_kernel<true><<<gx * gy, BB, sizeof(SmemBwdMma), s>>>(
          W, H, gx, ranges, point_list, idx_mask, rec, bg, final_T, n_contrib, dL_dpixels, dL_dalpha_img, hit, grad);
    else
      render_backward_mma_kernel<false><<<gx * gy, BB, sizeof(SmemBwdMma), s>>>(
          W, H, gx, ranges, point_list, idx_mask, rec, bg, final_T, n_contrib, dL_dpixels, dL_dalpha_img, hit, grad);
  } else if (cull) { if (dL_dalpha_img) SFB_RB(true, true); else SFB_RB(true, false); }
  else      { if (dL_dalpha_img) SFB_RB(false, true); else SFB_RB(false, false); }
#undef SFB_RB
}

}  // namespace sfb
