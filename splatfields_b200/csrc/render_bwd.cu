// render_bwd.cu — K7: per-pixel back-to-front gradient of the compositing, scattered onto the splats.
// Restates the external rasterizer's backward render: starting from
// final_T and the last contributor, walk the tile list backwards, rebuild alpha and T, and emit
//   d/d rgb_i, d/d opacity_i, d/d conic_i (A,B,C), d/d mean2D_i (NDC-scaled: includes 0.5*W, 0.5*H).
//
// The reference issues 9 global float atomics per (pixel, contributor).  Here each pixel only forms
// s = dL/dG*G and w = alpha*T; the sums over the tile's pixels (raw moments of s about the tile centre plus the
// three colour terms) are a small matrix product done on the tensor cores per warp (below), added into a
// per-tile shared-memory accumulator, and at the end of a batch one thread per splat turns the moments into
// gradients and issues one vectorised global reduction (2x red.global.add.v4.f32 + 1 scalar) per (tile, splat)
// instance: global atomic traffic drops from 9 * pixels * contributors to 3 * R.
// The cotangent of the fused coverage image (ALPHA, see render_fwd.cu) enters as a fourth channel with colour 1 and
// background 0: it only adds to dL/dalpha, exactly the sum the reference gets from the backward of its second pass.
// The (warp, entry) pairs swept are exactly the ones the forward recorded in hit[R] (pairs in which at least one
// pixel of the warp's 8x4 patch accumulated the entry).
// Accumulation order differs from the reference's (as it does between two runs of the reference), so
// parity here is tolerance-based: 1e-3 relative on every per-splat gradient.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <cstddef>

namespace sfb {


__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core pixel -> splat reduction.
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
// A recursive-halving shuffle tree (round 1's first version: 14 shuffles for 9 values) cost ~65 of the ~140
// instructions of an iteration; this costs ~12 per entry (1 STS, 1.5 LDS, 4 split ops, 1.25 MMA, 0.4 shared atomics).  The moments come out about the tile centre
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

constexpr int NT = 256;          // threads per CTA = pixels of a tile
constexpr int NW = NT / 32;      // warps = 8x4 pixel patches

// BBT = list entries per batch.  256 (default): 67.5 KB of shared memory / 80 registers -> 3 CTAs per SM;
// 128: 54 KB / 64 registers -> 4 CTAs per SM (measured slower in round 2: twice the per-batch overhead and spills).
template <int BBT>
struct SmemBwdMma {
  float4 row[BBT][4];               // staged 64-byte rows (common.cuh); [3] = true conic A, B, C + Gaussian index bits
  float2 stage[NW][MG * 32];        // per warp: [entry row][pixel ^ swizzle] = (sG, w)
  float2 dlp[NW][4][32];            // per warp: (hi, lo) of dL/dC_c per pixel; channel 3 = zeros
  alignas(16) uint8_t list[NW][BBT];   // per warp: compacted entry slots (read 16 at a time)
  alignas(16) float acc[BBT * 10];   // per entry: six moments, three colour sums, [9] = sum of w * dL/ddepth (DEPTH)
  uint64_t bar;
  uint32_t maxc[NW];
  uint32_t tile;
  uint8_t mask[BBT];
};
static_assert(offsetof(SmemBwdMma<128>, list) % 16 == 0 && offsetof(SmemBwdMma<256>, list) % 16 == 0, "16-byte list loads");
static_assert(offsetof(SmemBwdMma<128>, stage) % 16 == 0 && offsetof(SmemBwdMma<128>, bar) % 8 == 0, "smem alignment");

// Tile order: the forward left every tile in one of TILE_BUCKETS cost buckets (deepest contributor of the tile / 32)
// with a unique rank inside its bucket; CTA i takes the i-th tile counting from the most expensive bucket down, so the
// long tiles start first and the kernel's tail consists of cheap ones.
__device__ __forceinline__ int ordered_tile(int i, int T, const uint32_t* __restrict__ bcount,
                                            const uint32_t* __restrict__ btile, uint32_t* s_tile) {
  if (threadIdx.x < 32) {
    const int b = TILE_BUCKETS - 1 - (int)threadIdx.x;          // lane 0 = most expensive bucket
    const uint32_t c = bcount[b];
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)threadIdx.x >= o) inc += t;
    }
    const uint32_t ex = inc - c;
    if ((uint32_t)i >= ex && (uint32_t)i < inc) *s_tile = btile[(size_t)b * T + ((uint32_t)i - ex)];
  }
  __syncthreads();
  return (int)*s_tile;
}

// DEPTH: the cotangent of the depth image enters as one more composited channel ("colour" = the splat's view-space z,
// no background): it adds (z - depth accumulated behind) * dL/dD to dL/dalpha and yields a tenth per-splat sum,
// sum_p w_p dL/dD_p = dL/dz, which rides in the unused fourth column of the colour product.
template <bool ALPHA, int BBT, int MINB, bool TMA, bool PRED, bool DEPTH = false>
__global__ void __launch_bounds__(NT, MINB)
render_backward_mma_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                           const uint32_t* __restrict__ point_list, uint32_t idx_mask,
                           const SplatRec* __restrict__ rec, const __grid_constant__ CUtensorMap rec_map,
                           const float* __restrict__ bg, const float* __restrict__ final_T,
                           const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                           const float* __restrict__ dL_dalpha_img, const float* __restrict__ dL_ddepth_img,
                           const uint8_t* __restrict__ hit,
                           const uint32_t* __restrict__ bcount, const uint32_t* __restrict__ btile,
                           GradRec* __restrict__ grad) {
  constexpr int NA = DEPTH ? 10 : 9;      // sums per list entry
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemBwdMma<BBT>& sm = *reinterpret_cast<SmemBwdMma<BBT>*>(smem_raw);

  if (TMA && threadIdx.x == 0) mbar_init(&sm.bar, 1);
  const int tile = bcount ? ordered_tile((int)blockIdx.x, (int)gridDim.x, bcount, btile, &sm.tile) : (int)blockIdx.x;
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
  float dLpd = 0.f, accd = 0.f;
  if (DEPTH && inside) dLpd = dL_ddepth_img[pix];
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
    if (DEPTH) { split_tf32(dLpd, h, l); sm.dlp[warp][3][lane] = make_float2(__uint_as_float(h), __uint_as_float(l)); }
    else sm.dlp[warp][3][lane] = make_float2(0.f, 0.f);
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
  for (int w = 0; w < NW; w++) hi = max(hi, sm.maxc[w]);

  // mma.sync needs the whole warp converged; fail loudly rather than reduce garbage if it ever is not
  if (__activemask() != 0xffffffffu) __trap();

  uint32_t phase = 0;
  for (int top = (int)hi; top > 0; top -= BBT) {
    // smem slot j holds list position top-1-j (back to front)
    const int n = top < BBT ? top : BBT;
    __syncthreads();
    uint32_t mask = 0u;
    {
      uint32_t id = 0u;
      const bool mine = (int)threadIdx.x < n;
      const uint32_t pos = range.x + (uint32_t)(top - 1 - (int)threadIdx.x);
      if (mine) { id = point_list[pos] & idx_mask; mask = (uint32_t)hit[pos]; }
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
      if (TMA) {
        // every fourth thread gathers its own and its three neighbours' rows (entries past n re-read row 0)
        if (threadIdx.x < (unsigned)BBT) {        // (whole warps: BBT is a multiple of 32)
          const uint32_t i1 = __shfl_down_sync(0xffffffffu, id, 1), i2 = __shfl_down_sync(0xffffffffu, id, 2),
                         i3 = __shfl_down_sync(0xffffffffu, id, 3);
          if (threadIdx.x == 0) mbar_expect_tx(&sm.bar, (uint32_t)((n + 3) / 4) * 256u);
          if ((threadIdx.x & 3) == 0 && mine)
            tma_gather4(&sm.row[threadIdx.x][0], &rec_map, 0, (int)id, (int)i1, (int)i2, (int)i3, &sm.bar);
        }
        mbar_wait(&sm.bar, phase);
        phase ^= 1u;
        if (mine) { a = sm.row[threadIdx.x][0]; b = sm.row[threadIdx.x][1]; }
      } else if (mine) {
        const float4* rp = reinterpret_cast<const float4*>(rec + id);
        a = __ldg(rp); b = __ldg(rp + 1); c = __ldg(rp + 2);
        sm.row[threadIdx.x][2] = c;
      }
      if (mine) {
        if (!TMA) { sm.row[threadIdx.x][0] = a; sm.row[threadIdx.x][1] = b; }
        sm.row[threadIdx.x][3] = make_float4(0.f, 0.f, 0.f, __uint_as_float(id));
        if (TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // rows are overwritten by the TMA unit next batch
      }
    }
    if (threadIdx.x < (unsigned)BBT) sm.mask[threadIdx.x] = (uint8_t)mask;
    for (int k = threadIdx.x; k < NA * BBT; k += NT) sm.acc[k] = 0.f;
    __syncthreads();
    int nsweep = 0;
    {
      const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
      for (int c8 = 0; c8 < BBT / 32; c8++) {
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
        if (PRED) {
          // Predicated body (no branch per entry): a pixel that does not take the entry runs the same instructions with
          // alpha_eff = 0 — T, the running colour and the staged (sG, w) come out unchanged / zero — which costs nothing
          // extra in SIMT, drops the divergence bookkeeping and lets the 16 unrolled entries overlap their loads / exp.
          const float4 q0 = sm.row[j][0];
          const float4 q1 = sm.row[j][1];
          const float dx = q0.x - pixfx, dy = q0.y - pixfy;
          const float kp = render_power(q0.z, q0.w, q1.x, dx, dy);
          const float G = render_exp(kp);
          const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
          const bool c = (j >= jmin) & !((kp > 0.0f) | (alpha < 1.0f / 255.0f));   // the forward's own test
          const float ae = c ? alpha : 0.f;
          const float2 q2 = make_float2(sm.row[j][2].x, sm.row[j][2].y);
          float inv;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - alpha));
          inv = c ? inv : 1.f;
          T *= inv;
          wgt = ae * T;
          const float d0 = q1.w - acc0, d1 = q2.x - acc1, d2 = q2.y - acc2;
          float dL_dalpha = d0 * dLp0;
          dL_dalpha = fmaf(d1, dLp1, dL_dalpha);
          dL_dalpha = fmaf(d2, dLp2, dL_dalpha);
          acc0 = fmaf(ae, d0, acc0);
          acc1 = fmaf(ae, d1, acc1);
          acc2 = fmaf(ae, d2, acc2);
          if (ALPHA) {
            const float da = 1.f - acca;
            dL_dalpha = fmaf(da, dLpa, dL_dalpha);
            acca = fmaf(ae, da, acca);
          }
          if (DEPTH) {
            const float dz = q1.z - accd;
            dL_dalpha = fmaf(dz, dLpd, dL_dalpha);
            accd = fmaf(ae, dz, accd);
          }
          dL_dalpha = fmaf(dL_dalpha, T, -Tf_bg * inv);
          sG = c ? q1.y * dL_dalpha * G : 0.f;
        } else if (j >= jmin) {
          const float4 q0 = sm.row[j][0];
          const float4 q1 = sm.row[j][1];
          const float dx = q0.x - pixfx, dy = q0.y - pixfy;
          const float kp = render_power(q0.z, q0.w, q1.x, dx, dy);
          const float G = render_exp(kp);
          const float alpha = fminf(0.99f, __fmul_rn(q1.y, G));
          if (!((kp > 0.0f) | (alpha < 1.0f / 255.0f))) {        // the forward's own test: same pixels, same entries
            const float2 q2 = make_float2(sm.row[j][2].x, sm.row[j][2].y);
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
            if (DEPTH) {
              const float dz = q1.z - accd;
              dL_dalpha = fmaf(dz, dLpd, dL_dalpha);
              accd = fmaf(alpha, dz, accd);
            }
            dL_dalpha = fmaf(dL_dalpha, T, -Tf_bg * inv);
            sG = q1.y * dL_dalpha * G;
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
          float* a = &sm.acc[(int)sm.list[warp][g0 + gid] * NA];
          atomicAdd(a + 2 * tig, mom ? dS[0] : dW[0]);
          atomicAdd(a + 2 * tig + 1, mom ? dS[1] : dW[1]);
          if (tig == 2) atomicAdd(a + 8, dW[1]);
          if (DEPTH && tig == 2) atomicAdd(a + 9, dW[0]);      // column 4 of the colour product: sum of w * dL/ddepth
        }
        if (gid + 8 < gn) {
          float* a = &sm.acc[(int)sm.list[warp][g0 + gid + 8] * NA];
          atomicAdd(a + 2 * tig, mom ? dS[2] : dW[2]);
          atomicAdd(a + 2 * tig + 1, mom ? dS[3] : dW[3]);
          if (tig == 2) atomicAdd(a + 8, dW[3]);
          if (DEPTH && tig == 2) atomicAdd(a + 9, dW[2]);
        }
      }
      __syncwarp();   // staging rows are rewritten by the next group
    }
    __syncthreads();
    if ((int)threadIdx.x < n) {
      float a[NA];
      bool nz = false;
#pragma unroll
      for (int k = 0; k < NA; k++) { a[k] = sm.acc[threadIdx.x * NA + k]; nz |= (a[k] != 0.f); }
      if (nz) {
        const float4 q0 = sm.row[threadIdx.x][0];
        const float4 q1 = sm.row[threadIdx.x][1];
        const float4 q3 = sm.row[threadIdx.x][3];
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
        float* gp = reinterpret_cast<float*>(grad + __float_as_uint(q3.w));
        red_add_v4(gp, gx, gy, gA, gB);
        red_add_v4(gp + 4, gC, S0 != 0.f ? gop : 0.f, a[6], a[7]);
        atomicAdd(gp + 8, a[8]);
        if (DEPTH) atomicAdd(gp + 9, a[9]);      // GradRec::dz
      }
    }
  }
}


bool make_rec_tensor_map(const SplatRec* rec, size_t P, void* out_map);   // render_fwd.cu

// A/B knobs of the round-2 sessions (profiles/): SFB_BWD_STAGE=tma (TMA row gather instead of three 16-byte loads per
// thread), SFB_BWD_BATCH=128 (128-entry batches, 64 registers, 4 CTAs per SM) or 128x3 (128-entry batches at 80
// registers), SFB_BWD_ORDER=0 (tiles in launch order), SFB_BWD_SWEEP=pred (predicated per-entry body).
static int env_choice(const char* name, const char* alt) {
  const char* e = getenv(name);
  return (e && strcmp(e, alt) == 0) ? 1 : 0;
}

int launch_render_backward(int W, int H, const uint2* ranges, const uint32_t* point_list, uint32_t idx_mask,
                           const SplatRec* rec, size_t P,
                           const float* bg, const float* final_T, const uint32_t* n_contrib,
                           const float* dL_dpixels, const float* dL_dalpha_img, const float* dL_ddepth_img,
                           const uint8_t* hit,
                           const uint32_t* bcount, const uint32_t* btile, GradRec* grad,
                           cudaStream_t s) {
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  static int cfg = -1;
  if (cfg < 0) cfg = env_choice("SFB_BWD_STAGE", "tma") | (env_choice("SFB_BWD_BATCH", "128") << 1) |
                     (env_choice("SFB_BWD_ORDER", "0") << 2) | (env_choice("SFB_BWD_BATCH", "128x3") << 3) |
                     (env_choice("SFB_BWD_SWEEP", "pred") << 4);
  const bool tma = (cfg & 1) != 0, b128 = (cfg & 2) != 0, ordered = !(cfg & 4), b128x3 = (cfg & 8) != 0, pred = (cfg & 16) != 0;
  if (!ordered) { bcount = nullptr; btile = nullptr; }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (tma && !make_rec_tensor_map(rec, P, &map)) {
    set_error("cuTensorMapEncodeTiled failed for the splat record table (TMA staging of the tile lists)");
    return -1;
  }
  // > 48 KB of dynamic shared memory needs the attribute on EVERY device the process uses (it is per device and the
  // call is cheap: it is made with every launch)
#define SFB_RBK(A, B, MB, TM) do { if (pred) SFB_RBP(A, B, MB, TM, true, false); else SFB_RBP(A, B, MB, TM, false, false); } while (0)
#define SFB_RBP(A, B, MB, TM, PR, DP)                                                                               \
  do {                                                                                                              \
    auto kern = render_backward_mma_kernel<A, B, MB, TM, PR, DP>;                                                   \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemBwdMma<B>));            \
    kern<<<gx * gy, NT, sizeof(SmemBwdMma<B>), s>>>(W, H, gx, ranges, point_list, idx_mask, rec, map, bg, final_T,  \
                                                   n_contrib, dL_dpixels, dL_dalpha_img, dL_ddepth_img, hit, bcount, \
                                                   btile, grad);                                                    \
  } while (0)
#define SFB_RBA(A)                                                                                                  \
  do {                                                                                                              \
    if (b128) { if (tma) SFB_RBK(A, 128, 4, true); else SFB_RBK(A, 128, 4, false); }                                \
    else if (b128x3) { SFB_RBK(A, 128, 3, false); }   /* 128-entry batches at 80 registers (3 CTAs / SM) */           \
    else     { if (tma) SFB_RBK(A, 256, 3, true); else SFB_RBK(A, 256, 3, false); }                                 \
  } while (0)
  if (dL_ddepth_img) {     // depth cotangent: the default configuration of the kernel (the A/B arms do not carry it)
    if (dL_dalpha_img) SFB_RBP(true, 256, 3, false, false, true); else SFB_RBP(false, 256, 3, false, false, true);
  } else if (dL_dalpha_img) SFB_RBA(true); else SFB_RBA(false);
#undef SFB_RBA
#undef SFB_RBK
#undef SFB_RBP
  return 0;
}

}  // namespace sfb
