// render_fwd.cu — K6: per-16x16-tile front-to-back alpha compositing (RGB + depth).
// Restates the external rasterizer's forward render: for every pixel
// walk the tile's depth-ordered list; alpha = min(0.99, o * exp(power)); skip alpha < 1/255; stop before
// the contributor that would push T below 1e-4; C += rgb*alpha*T, D += depth*alpha*T; out = C + T*bg.
//
// One CTA per tile, 256 threads = one pixel each; a warp covers an 8x4 pixel patch (not a 16x2 strip)
// so that the pixels of a warp see nearly the same set of contributing splats.  The tile's list is
// consumed in batches of 256: every thread gathers one 48-byte SplatRec (three 16-byte loads out of L2,
// where the whole record array of a 1M-splat scene stays resident) into shared memory, then all threads
// sweep the batch reading the records as warp-wide broadcasts.
#include "common.cuh"
#include <cstdlib>

namespace sfb {

constexpr int RB = 256;  // batch = block size

// CULL: each fetched splat gets an 8-bit mask of the 8x4 patches (= warps) its alpha >= 1/255 footprint box
// can touch (SplatRec::hx/hy, computed conservatively in preprocess); every warp then sweeps only its own
// compacted sub-list.  Skipped (splat, patch) pairs are pairs the reference's own `alpha < 1/255` test would
// reject for all 32 pixels, so the output is bit-identical; the sweep shrinks ~3x on dense scenes.
// ALPHA: also accumulate the coverage image A = sum(alpha * T) — the image the reference obtains from a SECOND
// full rasterizer pass with colours = 1 and bg = 0 (gaussian_renderer/__init__.py:104-115); it shares every
// skip / stop decision with the colour pass, so one extra FADD per contribution replaces that whole pass.
// REC: record, per list entry, which warps (8x4 patches) accumulated it.  The backward sweeps exactly those
// (warp, entry) pairs instead of every pair whose footprint box touches the patch (3.1 M -> 1.9 M per view).
template <bool ALPHA>
__global__ void __launch_bounds__(RB, 6)   // <= 42 registers: the sweep loop needs ~40; the fetch-phase culling math may spill
render_forward_kernel(int W, int H, int grid_x, uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ point_list, uint32_t idx_mask,
                      const SplatRec* __restrict__ rec,
                      const float* __restrict__ bg, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_alpha,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, uint8_t* __restrict__ hit,
                      float4* __restrict__ zero16, uint32_t zero_per_cta, uint32_t zero_total) {
  // one struct = one base register: every access below is base + immediate (+ j * stride)
  struct Smem {
    float4 q0[RB];   // x, y, conA, conB
    float4 q1[RB];   // conC, opacity, depth, r
    float4 q2[RB];   // g, b, -, -   (16-byte stride like q0 / q1: one address register + immediates)
    uint8_t mask[RB];
    uint8_t list[RB / 32][RB];
    uint8_t hitw[RB / 32][RB];   // [warp][entry]: 1 = some pixel of the warp accumulated it
  };
  __shared__ Smem sm;
  float4* const s_q0 = sm.q0;
  float4* const s_q1 = sm.q1;
  float4* const s_q2 = sm.q2;
  uint8_t* const s_mask = sm.mask;
  uint8_t (*const s_list)[RB] = sm.list;

  const int tile = blockIdx.x;
  // Prologue: clear this CTA's slice of the backward's per-splat gradient accumulators (GradRec[P]).  The kernel is
  // issue-bound with an idle memory system, so these fire-and-forget stores replace a 48 B/splat memset in front of
  // the backward render for free.
  if (zero16) {
    const uint32_t z0 = (uint32_t)tile * zero_per_cta;
    const uint32_t z1 = min(z0 + zero_per_cta, zero_total);
    for (uint32_t i = z0 + threadIdx.x; i < z1; i += RB) zero16[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = tile_x * TILE_X + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pixfx = (float)px, pixfy = (float)py;
  const float tx0 = (float)(tile_x * TILE_X), ty0 = (float)(tile_y * TILE_Y);

  uint2 range = ranges[tile];
  if (range.y == 0u && range.x != 0u) {
    // tile without instances: the fused range pass (binning.cu) leaves (0xFFFFFFFF, 0); store the reference's (0, 0)
    range.x = 0u;
    if (threadIdx.x == 0) ranges[tile] = make_uint2(0u, 0u);
  }
  int todo = (int)(range.y - range.x);
  bool done = !inside;

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, Ac = 0.f;
  uint32_t last_contributor = 0;

  for (int base = 0; todo > 0; base += RB, todo -= RB) {
    if (__syncthreads_count(done) == RB) break;
    uint32_t mask = 0u;
    if ((int)threadIdx.x < todo) {
      uint32_t id = point_list[range.x + base + threadIdx.x] & idx_mask;
      const float4* rp = reinterpret_cast<const float4*>(rec + id);
      float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
      s_q0[threadIdx.x] = a;
      s_q1[threadIdx.x] = b;
      s_q2[threadIdx.x] = c;
      mask = refine_patch_mask(patch_mask(a.x, a.y, c.z, c.w, tx0, ty0), a.x, a.y, a.z, a.w, b.x, b.y, c.z, tx0, ty0);
    }
    s_mask[threadIdx.x] = (uint8_t)mask;
    reinterpret_cast<uint2*>(&sm.hitw[0][0])[threadIdx.x] = make_uint2(0u, 0u);   // 8 warps x 256 B
    __syncthreads();
    int n;
    {
      int cnt = 0;
      const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
      for (int c8 = 0; c8 < RB / 32; c8++) {
        const int idx = c8 * 32 + lane;
        const bool hit = (s_mask[idx] >> warp) & 1;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (hit) s_list[warp][cnt + __popc(bal & lt)] = (uint8_t)idx;
        cnt += __popc(bal);
      }
      __syncwarp();
      n = cnt;
    }
    // Sweep.  `done` pixels skip the whole batch; a pixel that saturates leaves the loop (no per-iteration flag
    // bookkeeping: the loop body is the kernel, every instruction in it is paid ~150 M times per view).
    if (!done) {
      int lastj = -1;
      for (int k = 0; k < n; k++) {
        const int j = (int)s_list[warp][k];
        const float4 q0 = s_q0[j];
        const float dx = q0.x - pixfx, dy = q0.y - pixfy;
        const float4 q1 = s_q1[j];
        // -0.5f*(A*dx*dx + C*dy*dy) - B*dx*dy, in the op order nvcc gives that expression
        const float s = __fmaf_rn(__fmul_rn(q0.z, dx), dx, __fmul_rn(__fmul_rn(q1.x, dy), dy));
        const float power = __fmaf_rn(s, -0.5f, -__fmul_rn(__fmul_rn(q0.w, dx), dy));
        if (power > 0.0f) continue;
        const float alpha = fminf(0.99f, __fmul_rn(q1.y, splat_exp(power)));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = __fmul_rn(T, 1.0f - alpha);
        if (test_T < 0.0001f) { done = true; break; }
        const float w = __fmul_rn(alpha, T);
        const float2 q2 = make_float2(s_q2[j].x, s_q2[j].y);
        C0 = __fmaf_rn(q1.w, w, C0);
        C1 = __fmaf_rn(q2.x, w, C1);
        C2 = __fmaf_rn(q2.y, w, C2);
        Dp = __fmaf_rn(q1.z, w, Dp);
        if (ALPHA) Ac = __fmaf_rn(1.0f, w, Ac);
        T = test_T;
        lastj = j;                                     // list order is kept by the compaction: the last one wins
        sm.hitw[warp][j] = 1;                          // same value from every contributing lane: benign
      }
      if (lastj >= 0) last_contributor = (uint32_t)(base + lastj + 1);   // 1-based position in the tile list
    }
    {
      __syncthreads();
      const int nfetch = todo < RB ? todo : RB;
      if ((int)threadIdx.x < nfetch) {
        uint32_t bits = 0;
#pragma unroll
        for (int w = 0; w < RB / 32; w++) bits |= (uint32_t)sm.hitw[w][threadIdx.x] << w;
        hit[range.x + base + threadIdx.x] = (uint8_t)bits;
      }
    }
  }
  if (inside) {
    const size_t pix = (size_t)py * W + px;
    const size_t HW = (size_t)H * W;
    final_T[pix] = T;
    n_contrib[pix] = last_contributor;
    out_color[pix] = __fmaf_rn(T, bg[0], C0);
    out_color[HW + pix] = __fmaf_rn(T, bg[1], C1);
    out_color[2 * HW + pix] = __fmaf_rn(T, bg[2], C2);
    out_depth[pix] = Dp;
    if (ALPHA) out_alpha[pix] = Ac;     // bg = 0 in the reference's alpha pass
  }
}

void launch_render_forward(int W, int H, uint2* ranges, const uint32_t* point_list, uint32_t idx_mask,
                           const SplatRec* rec,
                           const float* bg, float* out_color, float* out_depth, float* out_alpha, float* final_T,
                           uint32_t* n_contrib, uint8_t* hit, GradRec* zero_grad, size_t P, cudaStream_t s) {
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  // GradRec = 3 x 16 bytes; 32-bit word counts cover P < 2^30 (checked by the caller for num_rendered anyway)
  float4* const zero16 = (zero_grad && P > 0 && P < ((size_t)1 << 30)) ? reinterpret_cast<float4*>(zero_grad) : nullptr;
  const uint32_t zero_total = zero16 ? (uint32_t)(P * 3) : 0u;
  const uint32_t zero_per_cta = zero16 ? (zero_total + (uint32_t)(gx * gy) - 1u) / (uint32_t)(gx * gy) : 0u;
  if (out_alpha)
    render_forward_kernel<true><<<gx * gy, RB, 0, s>>>(W, H, gx, ranges, point_list, idx_mask, rec, bg, out_color, out_depth,
                                                       out_alpha, final_T, n_contrib, hit, zero16, zero_per_cta, zero_total);
  else
    render_forward_kernel<false><<<gx * gy, RB, 0, s>>>(W, H, gx, ranges, point_list, idx_mask, rec, bg, out_color, out_depth,
                                                        out_alpha, final_T, n_contrib, hit, zero16, zero_per_cta, zero_total);
}

}  // namespace sfb
