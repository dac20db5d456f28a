// render_fwd.cu — K6: per-16x16-tile front-to-back alpha compositing (RGB + depth).
// Restates the external rasterizer's forward render: for every pixel
// walk the tile's depth-ordered list; alpha = min(0.99, o * exp(power)); skip alpha < 1/255; stop before
// the contributor that would push T below 1e-4; C += rgb*alpha*T, D += depth*alpha*T; out = C + T*bg.
//
// One CTA per tile, 256 threads = one pixel each; a warp covers an 8x4 pixel patch (not a 16x2 strip)
// so that the pixels of a warp see nearly the same set of contributing splats.  The tile's list is
// consumed in batches of 256 entries staged in shared memory as 64-byte rows (common.cuh):
//   * LDG staging (default): every thread loads one record with three 16-byte loads (the record table stays
//     L2-resident) and stores it to shared memory; single buffer.
//   * TMA staging (SFB_FWD_STAGE=tma): the 48-byte SplatRec rows of a batch are gathered by the TMA unit —
//     cp.async.bulk.tensor.2d ... tile::gather4 (SASS: UTMALDG), four list entries per instruction, 64 instructions
//     per batch issued by every fourth thread — into one of two buffers; completion is an mbarrier transaction
//     count; the gather of batch k+1 is issued before batch k is swept.  Measured SLOWER on B200 for this access
//     pattern (0.156 vs 0.148 ms per lego_1m forward, profiles/r02a_ab_matrix.txt): a tile is saturated after ~1.2
//     batches, so there is little to overlap, and 64 four-row gathers of 48-byte rows through the SM's one TMA unit take
//     longer than 768 independent 16-byte loads through the LSU.  Kept as a tested variant (it goes through the parity
//     suite) and as the evidence the A/B rests on.
// The thread that owns a row then computes its culling mask and rewrites the conic pre-multiplied for the sweep;
// all threads sweep the batch reading the rows as warp-wide broadcasts.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>

namespace sfb {

constexpr int RB = 256;  // batch = block size

// Each fetched splat gets an 8-bit mask of the 8x4 patches (= warps) its alpha >= 1/255 footprint can touch
// (SplatRec::hx/hy box computed conservatively in preprocess, then the exact ellipse-vs-patch test of common.cuh);
// every warp sweeps only its own compacted sub-list.  Skipped (splat, patch) pairs are pairs the reference's own
// `alpha < 1/255` test would reject for all 32 pixels, so the output is unchanged; the sweep shrinks ~3x on dense scenes.
// ALPHA: also accumulate the coverage image A = sum(alpha * T) — the image the reference obtains from a SECOND
// full rasterizer pass with colours = 1 and bg = 0 (gaussian_renderer/__init__.py:104-115); it shares every
// skip / stop decision with the colour pass, so one extra FADD per contribution replaces that whole pass.
// The kernel records, per list entry, which warps (8x4 patches) accumulated it (hit[R], one byte): the backward
// sweeps exactly those (warp, entry) pairs (3.1 M -> 1.9 M per view on lego_1m).
template <int NBUF>
struct SmemFwd {
  float4 row[NBUF][RB][4];   // 64-byte rows (common.cuh); 128-byte aligned for the TMA unit
  uint64_t bar[2];
  uint8_t mask[RB];
  uint8_t list[RB / 32][RB];
  uint8_t hitw[RB / 32][RB];   // [warp][entry]: 1 = some pixel of the warp accumulated it
};

template <bool ALPHA, bool TMA>
__global__ void __launch_bounds__(RB, 6)   // <= 42 registers: the sweep loop needs ~40; the fetch-phase culling math may spill
render_forward_kernel(int W, int H, int grid_x, uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ point_list, uint32_t idx_mask,
                      const SplatRec* __restrict__ rec, const __grid_constant__ CUtensorMap rec_map,
                      const float* __restrict__ bg, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_alpha,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, uint8_t* __restrict__ hit,
                      uint32_t* __restrict__ bcount, uint32_t* __restrict__ btile,
                      float4* __restrict__ zero16, uint32_t zero_per_cta, uint32_t zero_total) {
  __shared__ __align__(128) SmemFwd<TMA ? 2 : 1> sm;
  __shared__ uint32_t s_deepest;
  if (threadIdx.x == 0) s_deepest = 0u;

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
  const int total = (int)(range.y - range.x);
  const int nbatch = (total + RB - 1) / RB;
  bool done = !inside;

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, Ac = 0.f;
  uint32_t last_contributor = 0;

  // TMA: gather the rows of batch `it` into buffer `buf`; every fourth thread issues one gather4 for its own and its
  // three neighbours' list entries (entries past the end of the list re-read row 0: harmless, never swept)
  auto issue = [&](const int it, const int buf) {
    const int cnt = min(RB, total - it * RB);
    uint32_t id = 0u;
    if ((int)threadIdx.x < cnt) id = point_list[range.x + (uint32_t)(it * RB) + threadIdx.x] & idx_mask;
    const uint32_t i1 = __shfl_down_sync(0xffffffffu, id, 1), i2 = __shfl_down_sync(0xffffffffu, id, 2),
                   i3 = __shfl_down_sync(0xffffffffu, id, 3);
    if (threadIdx.x == 0) mbar_expect_tx(&sm.bar[buf], (uint32_t)((cnt + 3) / 4) * 256u);
    if ((threadIdx.x & 3) == 0 && (int)threadIdx.x < cnt)
      tma_gather4(&sm.row[buf][threadIdx.x][0], &rec_map, 0, (int)id, (int)i1, (int)i2, (int)i3, &sm.bar[buf]);
  };

  if (TMA) {
    if (threadIdx.x == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); }
    __syncthreads();
    if (nbatch > 0) issue(0, 0);
  }

  for (int it = 0; it < nbatch; it++) {
    const int buf = TMA ? (it & 1) : 0;
    const int base = it * RB;
    const int cnt = min(RB, total - base);
    // every thread has left batch it-1 here: its buffer may be refilled
    const bool all_done = __syncthreads_count(done) == RB;
    if (all_done) {
      if (TMA) mbar_wait(&sm.bar[buf], (uint32_t)(it >> 1) & 1u);   // batch `it` is in flight: it must land before the CTA retires
      break;
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
    float4* const myrow = &sm.row[buf][threadIdx.x][0];
    if (TMA) {
      if (it + 1 < nbatch) issue(it + 1, buf ^ 1);
      mbar_wait(&sm.bar[buf], (uint32_t)(it >> 1) & 1u);
      if ((int)threadIdx.x < cnt) { a = myrow[0]; b = myrow[1]; c = myrow[2]; }
    } else if ((int)threadIdx.x < cnt) {
      const uint32_t id = point_list[range.x + (uint32_t)base + threadIdx.x] & idx_mask;
      const float4* rp = reinterpret_cast<const float4*>(rec + id);
      a = __ldg(rp); b = __ldg(rp + 1); c = __ldg(rp + 2);
      myrow[2] = c;
    }
    uint32_t mask = 0u;
    if ((int)threadIdx.x < cnt) {
      mask = refine_patch_mask(patch_mask(a.x, a.y, c.z, c.w, tx0, ty0), a.x, a.y, a.z, a.w, b.x, b.y, c.z, tx0, ty0);
      if (!TMA) { myrow[0] = a; myrow[1] = b; }
    }
    sm.mask[threadIdx.x] = (uint8_t)mask;
    reinterpret_cast<uint2*>(&sm.hitw[0][0])[threadIdx.x] = make_uint2(0u, 0u);   // 8 warps x 256 B
    __syncthreads();
    int n;
    {
      int k = 0;
      const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
      for (int c8 = 0; c8 < RB / 32; c8++) {
        const int idx = c8 * 32 + lane;
        const bool h = (sm.mask[idx] >> warp) & 1;
        const uint32_t bal = __ballot_sync(0xffffffffu, h);
        if (h) sm.list[warp][k + __popc(bal & lt)] = (uint8_t)idx;
        k += __popc(bal);
      }
      __syncwarp();
      n = k;
    }
    // Sweep.  `done` pixels skip the whole batch; a pixel that saturates leaves the loop (no per-iteration flag
    // bookkeeping: the loop body is the kernel, every instruction in it is paid ~100 M times per view).  A predicated
    // (branch-free) body was measured slower in round 2 (0.152 vs 0.148 ms, profiles/r02a_ab_matrix.txt): a quarter of the
    // swept (warp, entry) pairs have no contributing lane and leave the branching body after half of it.
    if (!done) {
      int lastj = -1;
      const float4* const rows = &sm.row[buf][0][0];
      for (int k = 0; k < n; k++) {
        const int j = (int)sm.list[warp][k];
        const float4 q0 = rows[4 * j];
        const float4 q1 = rows[4 * j + 1];
        const float dx = q0.x - pixfx, dy = q0.y - pixfy;
        const float kp = render_power(q0.z, q0.w, q1.x, dx, dy);
        const float alpha = fminf(0.99f, __fmul_rn(q1.y, render_exp(kp)));
        if ((kp > 0.0f) | (alpha < 1.0f / 255.0f)) continue;     // the reference's two skips (power > 0, alpha < 1/255)
        const float test_T = __fmul_rn(T, 1.0f - alpha);
        if (test_T < 0.0001f) { done = true; break; }
        const float w = __fmul_rn(alpha, T);
        const float4 q2 = rows[4 * j + 2];
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
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      uint32_t bits = 0;
#pragma unroll
      for (int w = 0; w < RB / 32; w++) bits |= (uint32_t)sm.hitw[w][threadIdx.x] << w;
      hit[range.x + (uint32_t)base + threadIdx.x] = (uint8_t)bits;
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
  // Tile order of the backward: cost bucket = deepest contributor of the tile / 32; a unique rank inside the bucket
  // comes from one atomic per tile (render_bwd.cu: ordered_tile).
  {
    const uint32_t m = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0 && m) atomicMax(&s_deepest, m);
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t b = min(s_deepest >> 5, (uint32_t)(TILE_BUCKETS - 1));
      const uint32_t r = atomicAdd(&bcount[b], 1u);
      btile[(size_t)b * gridDim.x + r] = (uint32_t)tile;
    }
  }
}

// ---- tensor map over the SplatRec table: [P] rows of 12 floats, box = 16 floats x 1 row (the 4 floats past the row
// are out of bounds of the 12-wide tensor and arrive as zeros), so that a gathered row fills one 64-byte smem row.
// cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool make_rec_tensor_map(const SplatRec* rec, size_t P, void* out_map) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || P == 0) return false;
  const cuuint64_t dims[2] = {12, (cuuint64_t)P};
  const cuuint64_t strides[1] = {sizeof(SplatRec)};          // bytes between rows (dimension 1)
  const cuuint32_t box[2] = {16, 1};
  const cuuint32_t estr[2] = {1, 1};
  return fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<SplatRec*>(rec), dims,
            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Stand-alone exercise of the staging primitive (sfb_debug_gather_rows): out[i][0..15] = the 64-byte shared-memory row
// the TMA unit delivers for list entry i — table[idx[i]][0..11] followed by four zeros.  One CTA per 256 entries.
__global__ void __launch_bounds__(RB)
gather_rows_probe_kernel(const __grid_constant__ CUtensorMap map, const uint32_t* __restrict__ idx, int n,
                         float* __restrict__ out) {
  __shared__ __align__(128) float4 row[RB][4];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  const int base = blockIdx.x * RB;
  const int cnt = min(RB, n - base);
  uint32_t id = 0u;
  if ((int)threadIdx.x < cnt) id = idx[base + threadIdx.x];
  const uint32_t i1 = __shfl_down_sync(0xffffffffu, id, 1), i2 = __shfl_down_sync(0xffffffffu, id, 2),
                 i3 = __shfl_down_sync(0xffffffffu, id, 3);
  if (threadIdx.x == 0) mbar_expect_tx(&bar, (uint32_t)((cnt + 3) / 4) * 256u);
  if ((threadIdx.x & 3) == 0 && (int)threadIdx.x < cnt)
    tma_gather4(&row[threadIdx.x][0], &map, 0, (int)id, (int)i1, (int)i2, (int)i3, &bar);
  mbar_wait(&bar, 0);
  if ((int)threadIdx.x < cnt) {
    float4* o = reinterpret_cast<float4*>(out + (size_t)(base + threadIdx.x) * 16);
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = row[threadIdx.x][k];
  }
}

int launch_gather_rows_probe(const SplatRec* table, size_t P, const uint32_t* idx, int n, float* out, cudaStream_t s) {
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (!make_rec_tensor_map(table, P, &map)) { set_error("cuTensorMapEncodeTiled failed"); return -1; }
  if (n > 0) gather_rows_probe_kernel<<<(n + RB - 1) / RB, RB, 0, s>>>(map, idx, n, out);
  return 0;
}

static bool fwd_stage_tma() {      // SFB_FWD_STAGE=tma: the TMA row gather instead of three 16-byte loads per thread
  static int v = -1;
  if (v < 0) { const char* e = getenv("SFB_FWD_STAGE"); v = (e && e[0] == 't') ? 1 : 0; }
  return v == 1;
}

int launch_render_forward(int W, int H, uint2* ranges, const uint32_t* point_list, uint32_t idx_mask,
                          const SplatRec* rec,
                          const float* bg, float* out_color, float* out_depth, float* out_alpha, float* final_T,
                          uint32_t* n_contrib, uint8_t* hit, uint32_t* bcount, uint32_t* btile, GradRec* zero_grad,
                          size_t P, cudaStream_t s) {
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  // GradRec = 3 x 16 bytes; 32-bit word counts cover P < 2^30 (checked by the caller for num_rendered anyway)
  float4* const zero16 = (zero_grad && P > 0 && P < ((size_t)1 << 30)) ? reinterpret_cast<float4*>(zero_grad) : nullptr;
  const uint32_t zero_total = zero16 ? (uint32_t)(P * 3) : 0u;
  const uint32_t zero_per_cta = zero16 ? (zero_total + (uint32_t)(gx * gy) - 1u) / (uint32_t)(gx * gy) : 0u;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const bool tma = fwd_stage_tma();
  if (tma && !make_rec_tensor_map(rec, P, &map)) {
    set_error("cuTensorMapEncodeTiled failed for the splat record table (TMA staging of the tile lists)");
    return -1;
  }
#define SFB_RF(A, TM)                                                                                               \
  render_forward_kernel<A, TM><<<gx * gy, RB, 0, s>>>(W, H, gx, ranges, point_list, idx_mask, rec, map, bg, out_color, \
                                                      out_depth, out_alpha, final_T, n_contrib, hit, bcount, btile, \
                                                      zero16, zero_per_cta, zero_total)
  if (tma) { if (out_alpha) SFB_RF(true, true); else SFB_RF(false, true); }
  else     { if (out_alpha) SFB_RF(true, false); else SFB_RF(false, false); }
#undef SFB_RF
  return 0;
}

}  // namespace sfb
