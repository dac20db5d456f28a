// common.cuh — shared types, scratch-buffer layouts and launch prototypes of the sm_100a rasterizer.
// Path and boundary: DESIGN.md §1-2; reference surface: SURVEY.md §8b (the external
// diff_gaussian_rasterization extension imported at gaussian_renderer/__init__.py:14).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

namespace sfb {

constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr int NUM_SMS_B200 = 148;
constexpr int TILE_BUCKETS = 32;     // cost buckets of the backward's tile order (deepest contributor of a tile / 32)

// One record per Gaussian, written by preprocess, gathered (L2-resident) by both render kernels.
// 48 bytes = three 16-byte loads, always exactly two 32-byte sectors.
struct __align__(16) SplatRec {
  float x, y, conA, conB;   // screen mean (pixels), conic A, B
  float conC, opacity, depth, r;
  float g, b, hx, hy;       // hx, hy: conservative half-extents of the alpha >= 1/255 footprint (-1: never)
};
static_assert(sizeof(SplatRec) == 48, "SplatRec must be 48 bytes");

// Bit w of the returned mask is set iff the splat's footprint box can touch the 8x4-pixel patch of warp w
// (patch w: x-half w&1, y-quarter w>>1) of the 16x16 tile whose first pixel is (tx0, ty0).
__device__ __forceinline__ uint32_t patch_mask(float x, float y, float hx, float hy, float tx0, float ty0) {
  if (hx < 0.f) return 0u;
  const float xl = x - hx, xh = x + hx, yl = y - hy, yh = y + hy;
  uint32_t mx = 0u;
  if (xh >= tx0 && xl <= tx0 + 7.f) mx |= 1u;
  if (xh >= tx0 + 8.f && xl <= tx0 + 15.f) mx |= 2u;
  uint32_t m = 0u;
#pragma unroll
  for (int r = 0; r < 4; r++)
    if (yh >= ty0 + 4.f * r && yl <= ty0 + 4.f * r + 3.f) m |= mx << (2 * r);
  return m;
}

// 256-bit global accesses (sm_100a: LDG.E.ENL2.256 / STG.E.ENL2.256): one full 32-byte sector per lane, which
// is what a one-thread-per-row walk over 192-byte SH rows needs — with 128-bit accesses every sector is touched by
// two separate requests (measured on B200: 4.1 TB/s vs 6+ TB/s for a [1M][48] fp32 row copy).  p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// exp() of the compositing kernels: one MUFU.EX2 on power * log2(e) instead of libdevice expf's 10-instruction
// sequence (the sweep loops are issue-bound: -6 % forward, -5 % backward render time).  alpha differs from the expf
// value by <= ~7e-7 relative; measured against the fp64-accumulating oracle the image / depth / gradient errors are the
// same as with expf (lego_100k: RGB 3.0e-7 vs 3.3e-7, depth 1.7e-6 vs 1.4e-6 max abs; gradients 6e-7..8e-7 norm-wise
// either way; profiles/r01w_parity_margin.jsonl) — a factor 6 inside the 1e-5 tolerance.
// -DSFB_EXACT_EXP (the libsplat_b200_exactexp.so build variant) restores expf, the function the reference calls.
__device__ __forceinline__ float splat_exp(float power) {
#ifdef SFB_EXACT_EXP
  return expf(power);
#else
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(power * 1.4426950408889634f));
  return r;
#endif
}

// ---- staged splat rows of the two render kernels ----
// A list entry is staged in shared memory as one 64-byte row of four float4:
//   [0] x, y, conA, conB    [1] conC, opacity, depth, r    [2] g, b, hx, hy    [3] -, -, -, Gaussian index (backward)
// The first 48 bytes are the SplatRec as it sits in HBM (TMA row gather, or three 16-byte loads).
// power = -0.5 (A dx^2 + C dy^2) - B dx dy  is evaluated in the op order nvcc gives the reference's expression
// (round 2 tried a conic pre-multiplied by -log2(e)/2 — 5 instead of 9 flops — and lost the needle-splat parity case:
// with a nearly singular conic the three terms cancel to 1e-4 of their size, so the ROUNDING ORDER is part of the result;
// measured 1.4e-4 image error against 1e-5 allowed).  Forward and backward share the expression.
__device__ __forceinline__ float render_power(float A, float B, float C, float dx, float dy) {
  const float s = __fmaf_rn(__fmul_rn(A, dx), dx, __fmul_rn(__fmul_rn(C, dy), dy));
  return __fmaf_rn(s, -0.5f, -__fmul_rn(__fmul_rn(B, dx), dy));
}
__device__ __forceinline__ float render_exp(float power) { return splat_exp(power); }

// ---- TMA row gather (cp.async.bulk.tensor ... tile::gather4 -> UTMALDG): four rows of a 2-D tensor per instruction ----
// dst: 4 consecutive box rows in shared memory (128-byte aligned); tmap: CUtensorMap over the row table with
// box = {row floats, 1}; col: first column; r0..r3: row indices; completion on the mbarrier (box bytes x 4).
__device__ __forceinline__ void tma_gather4(void* dst_smem, const void* tmap, int col, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"((uint32_t)__cvta_generic_to_shared(bar))
      : "memory");
}

// Non-blocking L2 prefetch of the 128-byte line holding p (no register, no scoreboard).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- TMA (bulk async copy) + mbarrier helpers: 1-D cp.async.bulk between global and shared memory ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0u;
}
// Bounded: a transaction count that never completes (a bad tensor map, a miscounted expect_tx) must not hang the GPU —
// after ~2 s the kernel traps and the caller gets a CUDA error.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
// global -> shared, completion signalled on the mbarrier (bytes: multiple of 16, both sides 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global (bulk group); call bulk_store_fence() after the generic-proxy smem writes, before issuing
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Exact refinement of a footprint-box mask: for every candidate patch, the minimum of the conic form
// q(d) = A dx^2 + 2 B dx dy + C dy^2 over the patch rectangle (a convex quadratic over a box: 0 if the centre is
// inside, else attained on one of the four edges) is compared with tau = 2 ln(255 o): alpha >= 1/255 needs
// q <= tau.  The comparison carries an absolute + relative slack that covers the fp32 rounding of the reference's
// own per-pixel evaluation (4e-6 * the largest term magnitude over the patch), so a pair is dropped only when the
// reference's `alpha < 1/255` test provably rejects all 32 pixels.  Costs ~40 instructions per candidate patch, once
// per fetched splat, and removes most of the ~30 % of box-passing (warp, splat) sweeps that never contribute.
__device__ __forceinline__ uint32_t refine_patch_mask(uint32_t mask, float x, float y, float A, float B, float C,
                                                      float opacity, float hx, float tx0, float ty0) {
  if (mask == 0u || hx > 1.0e37f || !(A > 0.f) || !(C > 0.f)) return mask;   // nothing to do / never-cull splats
  const float tau = 2.f * __logf(255.f * opacity);
  const float nBA = -B / A, nBC = -B / C;
  uint32_t out = 0u;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    if (!((mask >> w) & 1u)) continue;
    const float X0 = tx0 + (float)((w & 1) * 8) - x, X1 = X0 + 7.f;
    const float Y0 = ty0 + (float)((w >> 1) * 4) - y, Y1 = Y0 + 3.f;
    bool keep = (X0 <= 0.f && X1 >= 0.f && Y0 <= 0.f && Y1 >= 0.f);
    if (!keep) {
      float qmin = 3.0e38f;
      {  // vertical edges x = X0, X1
        const float d0 = fminf(fmaxf(nBC * X0, Y0), Y1), d1 = fminf(fmaxf(nBC * X1, Y0), Y1);
        qmin = fminf(qmin, A * X0 * X0 + 2.f * B * X0 * d0 + C * d0 * d0);
        qmin = fminf(qmin, A * X1 * X1 + 2.f * B * X1 * d1 + C * d1 * d1);
      }
      {  // horizontal edges y = Y0, Y1
        const float d0 = fminf(fmaxf(nBA * Y0, X0), X1), d1 = fminf(fmaxf(nBA * Y1, X0), X1);
        qmin = fminf(qmin, A * d0 * d0 + 2.f * B * d0 * Y0 + C * Y0 * Y0);
        qmin = fminf(qmin, A * d1 * d1 + 2.f * B * d1 * Y1 + C * Y1 * Y1);
      }
      const float DX = fmaxf(fabsf(X0), fabsf(X1)), DY = fmaxf(fabsf(Y0), fabsf(Y1));
      const float M = A * DX * DX + C * DY * DY + 2.f * fabsf(B) * DX * DY;
      keep = !(qmin > tau * 1.0001f + 2.0e-3f + 8.0e-6f * M);     // NaNs keep
    }
    if (keep) out |= 1u << w;
  }
  return out;
}

// Per-Gaussian gradient accumulator filled by the backward render (atomics), consumed by the
// per-Gaussian backward.  Same 48-byte shape so one splat's partials share two sectors.
struct __align__(16) GradRec {
  float dx, dy, dA, dB;     // d/d(NDC-scaled mean) x,y ; d/d conic A, B (true derivatives)
  float dC, dop, dr, dg;    // d/d conic C ; d/d opacity ; d/d rgb
  float db, dz, pad1, pad2;   // dz: d/d view-space depth (only with a depth cotangent; stays 0 otherwise)
};
static_assert(sizeof(GradRec) == 48, "GradRec must be 48 bytes");

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Every carved array is followed by a 256-byte red zone.  With debug=True the API fills the red zones with a
// pattern before the kernels run and verifies them afterwards (forward and backward), which catches writes that
// run past an array but stay inside the caller's single allocation — invisible to compute-sanitizer.
constexpr size_t REDZONE_BYTES = 256;
struct RedzoneList { char* ptr[64]; int n = 0; };
inline RedzoneList*& redzone_collector() { static thread_local RedzoneList* c = nullptr; return c; }

template <typename T>
inline T* carve(char*& p, size_t n) {
  size_t off = align_up(reinterpret_cast<size_t>(p), 256);
  T* r = reinterpret_cast<T*>(off);
  char* end = reinterpret_cast<char*>(r + n);
  char* rz = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(end), 16));
  if (RedzoneList* c = redzone_collector()) { if (c->n < 64) c->ptr[c->n++] = rz; }
  p = rz + REDZONE_BYTES;
  return r;
}

// Radix-sort geometry shared by histogram / scatter kernels.
constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 16;                          // items per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;    // items per block
constexpr int SORT_MAX_BINS = 256;
__host__ __device__ inline int sort_blocks(int n) { return (n + SORT_TILE - 1) / SORT_TILE; }
// words of sort scratch for n items: 4 digit histograms + tickets + one look-back state row per pass
// (worst case 4 passes x 256 bins).
__host__ __device__ inline size_t sort_scratch_words(size_t n) {
  // (small inputs use 1024-item tiles: at most 4 * 4 * NUM_SMS_B200 of them)
  const size_t blocks = (size_t)sort_blocks((int)n) + 1;
  const size_t blocks_small = (n + 1023) / 1024 + 1;
  const size_t nb = blocks < (size_t)4 * NUM_SMS_B200 ? blocks_small : 2 * blocks;   // 2x: 2048-item tile variant
  return (size_t)4 * SORT_MAX_BINS + 8 + (size_t)4 * SORT_MAX_BINS * nb + 16;
}

// Depth ranks expanded per block of the instance-emission kernels (binning.cu); also the granularity of the
// block_sums array.  256 keeps >= 4 CTAs per SM in flight even for 500k-splat scenes with ~140 tiles per splat.
constexpr int DUP_GPB = 256;
__host__ __device__ inline size_t dup_blocks(size_t P) { return (P + DUP_GPB - 1) / DUP_GPB; }

// ---- P-sized scratch ("geomBuffer") ----
struct GeomState {
  SplatRec* rec;            // [P]
  uint32_t* tiles_touched;  // [P]
  uint2* rect;              // [P]    packed tile rectangle: x = xmin | ymin<<16, y = xmax | ymax<<16
  uint8_t* clamped;         // [P]    bit c set: channel c was clamped at 0
  uint32_t* depth_key[2];   // [P]    ping-pong keys of the depth sort (0xFFFFFFFF = culled)
  uint32_t* depth_idx[2];   // [P]    ping-pong values (Gaussian index)
  uint32_t* sort_hist;      // [SORT_MAX_BINS * sort_blocks(P)]
  uint32_t* block_sums;     // [dup_blocks(P) + 2]    per-block instance counts in depth order, then scanned
  uint32_t* counters;       // [8]    [0] = num_rendered, [2] = ~(smallest depth key of a visible splat)
  GradRec* grad;            // [P]    backward accumulators

  static GeomState from_chunk(char*& chunk, size_t P) {
    GeomState g;
    g.rec = carve<SplatRec>(chunk, P);
    g.tiles_touched = carve<uint32_t>(chunk, P);
    g.rect = carve<uint2>(chunk, P);
    g.clamped = carve<uint8_t>(chunk, P);
    for (int i = 0; i < 2; i++) g.depth_key[i] = carve<uint32_t>(chunk, P);
    for (int i = 0; i < 2; i++) g.depth_idx[i] = carve<uint32_t>(chunk, P);
    g.sort_hist = carve<uint32_t>(chunk, sort_scratch_words(P));
    g.block_sums = carve<uint32_t>(chunk, dup_blocks(P) + 2);
    g.counters = carve<uint32_t>(chunk, 8);
    g.grad = carve<GradRec>(chunk, P);
    return g;
  }
  static size_t required(size_t P) {
    char* p = nullptr;
    from_chunk(p, P);
    return reinterpret_cast<size_t>(p) + 256;
  }
};

// Instance packing.  An instance is ONE 32-bit word, tile << low_bits | (gaussian index & low mask), and the tile sort
// moves bare keys.  When ceil(log2 P) + ceil(log2 T) <= 32 the whole index fits (high_bits = 0).  Otherwise (e.g. 2 M
// splats at 1080p: 21 + 13 bits) the word keeps the low 32 - tile_bits index bits and the remaining high_bits <= 8 ride
// along as one byte per instance — 5 instead of 8 bytes per instance and pass — until the last sort pass re-assembles
// the full index as its only output word.  Either way the render kernels read a plain index list: word & idx_mask.
struct InstPacking {
  int low_bits;       // index bits inside the word
  int high_bits;      // index bits in the side byte (0: none)
  uint32_t idx_mask;  // mask of the sorted list the render kernels read (high_bits > 0: the list holds merged indices)
  bool ok;            // false: ceil(log2 P) + ceil(log2 T) > 40 (not supported)
};
inline int ceil_log2(size_t n) { int b = 0; while (((size_t)1 << b) < n) b++; return b; }
inline int tile_bits_for(size_t T) { int b = 1; while (((size_t)1 << b) < T) b++; return b; }
// SFB_NO_PACK=1 forces the split layout (two index bits in the side byte) so that the parity tests can drive that
// path with small scenes.
inline bool inst_split_forced() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SFB_NO_PACK"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline InstPacking inst_packing(size_t P, size_t T) {
  InstPacking k;
  const int ib = ceil_log2(P < 2 ? 2 : P), tb = tile_bits_for(T);
  k.ok = true;
  if (ib + tb <= 32 && !(inst_split_forced() && ib > 2)) {
    k.low_bits = ib; k.high_bits = 0;
    k.idx_mask = ib >= 32 ? 0xFFFFFFFFu : ((1u << ib) - 1u);
  } else {
    k.low_bits = ib + tb <= 32 ? ib - 2 : 32 - tb;
    k.high_bits = ib - k.low_bits;
    k.idx_mask = 0xFFFFFFFFu;
    k.ok = k.high_bits <= 8;
  }
  return k;
}

// ---- R-sized scratch ("binningBuffer") ----
struct BinState {
  uint32_t* tile_key[2];    // [R]  ping-pong instance words (tile | index); the final half is the list the render kernels read
  uint8_t* inst_hi[2];      // [R]  ping-pong high index bits (split instances only; nullptr otherwise)
  uint32_t* sort_hist;      // sort scratch
  uint2* ranges;            // [T]
  uint8_t* hit;             // [R]  per sorted instance: bit w = warp (8x4 patch) w accumulated it in the forward

  static BinState from_chunk(char*& chunk, size_t R, size_t T, bool split) {
    BinState b;
    for (int i = 0; i < 2; i++) b.tile_key[i] = carve<uint32_t>(chunk, R);
    for (int i = 0; i < 2; i++) b.inst_hi[i] = split ? carve<uint8_t>(chunk, R + 16) : nullptr;
    b.sort_hist = carve<uint32_t>(chunk, sort_scratch_words(R));
    b.ranges = carve<uint2>(chunk, T);
    b.hit = carve<uint8_t>(chunk, R + 256);
    return b;
  }
  static size_t required(size_t R, size_t T, bool split) {
    char* p = nullptr;
    from_chunk(p, R, T, split);
    return reinterpret_cast<size_t>(p) + 256;
  }
  // the sorted instance list the render kernels walk: entries are `word & idx_mask`
  const uint32_t* point_list(int final_buf) const { return tile_key[final_buf]; }
};

// ---- pixel-sized scratch ("imgBuffer") ----
struct ImgState {
  float* final_T;           // [H*W]
  uint32_t* n_contrib;      // [H*W]
  uint32_t* tile_bcount;    // [TILE_BUCKETS]      tiles per cost bucket (written by the forward render)
  uint32_t* tile_btile;     // [TILE_BUCKETS][T]   tile ids of every bucket, in arrival order
  static ImgState from_chunk(char*& chunk, size_t N, size_t T) {
    ImgState s;
    s.final_T = carve<float>(chunk, N);
    s.n_contrib = carve<uint32_t>(chunk, N);
    s.tile_bcount = carve<uint32_t>(chunk, TILE_BUCKETS);
    s.tile_btile = carve<uint32_t>(chunk, (size_t)TILE_BUCKETS * T);
    return s;
  }
  static size_t required(size_t N, size_t T) {
    char* p = nullptr;
    from_chunk(p, N, T);
    return reinterpret_cast<size_t>(p) + 256;
  }
};

// Number of tile-id bits the tile sort has to cover.
inline int tile_bits(int num_tiles) {
  int b = 1;
  while ((1 << b) < num_tiles) b++;
  return b;
}

// per-kernel profiling brackets (api.cu); no-ops unless sfb_profile_enable(1)
void prof_begin(const char* name, cudaStream_t s);
void prof_end(cudaStream_t s);
// text returned by sfb_last_error() on the calling thread (api.cu)
void set_error(const char* msg);

// ------------------------------------------------------------------ launchers (one per .cu file)
struct FwdParams {
  int P, D, M, W, H;
  const float *bg, *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp;
  const float *viewmatrix, *projmatrix, *campos;
  float scale_modifier, tan_fovx, tan_fovy;
  int prefiltered;
  int wide256;      // SH rows are 32-byte aligned multiples of 32 bytes: use 256-bit loads
  uint32_t* zero_ptr;     // or nullptr: words the kernel's blocks clear as a prologue (the depth sort's scratch)
  uint32_t zero_words;
  uint32_t* nr_host;      // or nullptr: pinned host word; the last block to finish stores num_rendered there
};

// preprocess.cu
void launch_preprocess(const FwdParams& p, const GeomState& g, int* radii, cudaStream_t s);
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s);
void launch_export_geom(int P, const GeomState& g, const float* scales, const float* rotations, float mod,
                        const float* cov3D_precomp, float* means2D, float* depths, float* cov3D,
                        float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched,
                        cudaStream_t s);

// binning.cu
// Stable LSD radix sort of (key, value) pairs on key bits [0, nbits); returns the index (0/1) of the
// ping-pong half that holds the result.  hist must hold sort_scratch_words(n) words.
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t* hist, int n, int nbits, cudaStream_t s,
                     int* launches, const char* const* names /* {hist, scan, scatter} */,
                     const uint32_t* bias_c = nullptr /* device: ~min key; keys are rebased in place */,
                     int first_bit = 0 /* digits start at this bit; vals[0] == nullptr sorts bare keys */,
                     bool scratch_zeroed = false /* the first radix_sort_zero_words(n, nbits) words of hist are already 0 */,
                     uint2* ranges = nullptr /* fused K5: [T] pre-set to (0xFFFFFFFF, 0); filled by the last pass */,
                     int tile_shift = 0 /* tile id = key >> tile_shift */,
                     uint8_t* const* vals8 = nullptr /* one side byte per key (ping-pong), vals == nullptr */,
                     int merge_bits = 0 /* last pass writes (key & ((1 << merge_bits) - 1)) | byte << merge_bits */,
                     const uint32_t** last_total0 = nullptr /* non-null: the last pass may return at once when it is the
                        identity; receives the device word to compare with n (SortedIdx), or nullptr */);
size_t radix_sort_zero_words(int n, int nbits);
// Result of the depth sort as its consumers see it.  The most significant radix pass of a bounded scene is the identity
// (depth keys are rebased to key - min: < 2^24), known on the device only: instead of copying 8 B/Gaussian through
// that pass, the pass returns at once and the consumers read its input.
struct SortedIdx {
  const uint32_t* primary;   // the last pass's output
  const uint32_t* alt;       // the last pass's input
  const uint32_t* total0;    // number of keys with digit 0 in the last pass (nullptr: the pass always runs)
  uint32_t n;
  __device__ __forceinline__ const uint32_t* get() const { return (total0 && *total0 == n) ? alt : primary; }
};
void launch_instance_block_sums(int P, const SortedIdx& sorted_idx, const uint32_t* tiles_touched,
                                uint32_t* block_sums, cudaStream_t s);
void launch_duplicate(int P, int grid_x, const SortedIdx& sorted_idx, const uint32_t* tiles_touched,
                      const uint2* rect, const uint32_t* block_offsets, uint32_t* tile_keys,
                      uint8_t* inst_hi /* nullptr: the whole index fits the word */, int idx_bits,
                      uint32_t* zero_ptr /* or nullptr */, size_t zero_words, uint2* ranges_init /* or nullptr */, int T,
                      uint32_t* bcount_zero /* [TILE_BUCKETS] or nullptr: cleared on the way */, cudaStream_t s);
void launch_export_keys(int R, int T, const uint32_t* words, bool merged /* words are bare indices */, int idx_bits,
                        const uint2* ranges, const SplatRec* rec, uint64_t* out_keys, uint32_t* out_list, cudaStream_t s);

// render_fwd.cu
// returns 0, or -1 with set_error() (tensor-map encoding failed)
int launch_render_forward(int W, int H, uint2* ranges /* empty tiles are normalised to (0, 0) in place */,
                           const uint32_t* point_list, uint32_t idx_mask,
                           const SplatRec* rec,
                           const float* bg, float* out_color, float* out_depth, float* out_alpha, float* final_T,
                           uint32_t* n_contrib, uint8_t* hit /* [R] */,
                           uint32_t* bcount, uint32_t* btile /* backward tile order (ImgState), bcount zeroed */,
                           GradRec* zero_grad /* [P] or nullptr: cleared by the CTAs as a prologue */, size_t P,
                           cudaStream_t s);

int launch_gather_rows_probe(const SplatRec* table, size_t P, const uint32_t* idx, int n, float* out, cudaStream_t s);

// render_bwd.cu
int launch_render_backward(int W, int H, const uint2* ranges, const uint32_t* point_list, uint32_t idx_mask,
                           const SplatRec* rec, size_t P,
                           const float* bg, const float* final_T, const uint32_t* n_contrib,
                           const float* dL_dpixels, const float* dL_dalpha_img,
                           const float* dL_ddepth_img /* [H][W] cotangent of the depth image, or nullptr */, const uint8_t* hit,
                           const uint32_t* bcount /* [TILE_BUCKETS] tiles per cost bucket (forward) */,
                           const uint32_t* btile /* [TILE_BUCKETS][T] tile ids per bucket */, GradRec* grad, cudaStream_t s);

// geom_bwd.cu
constexpr int XCHG_MAX_RANKS = 16;
struct BwdParams {
  int P, D, M, W, H;
  const float *means3D, *shs, *colors_precomp, *scales, *rotations, *cov3D_precomp;
  const float *viewmatrix, *projmatrix, *campos;
  float scale_modifier, tan_fovx, tan_fovy;
  int wide256;      // SH / dL_dsh rows are 32-byte aligned multiples of 32 bytes: use 256-bit loads and stores
  int sh_factored;  // SFB_BWD_SH_FACTORED: dL_dcolors receives the clamp-masked colour gradient, dL_dsh may be nullptr
  const int* radii;
  float *dL_dmeans2D, *dL_dcolors, *dL_dopacity, *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscales, *dL_drot;
  // view-parallel exchange over NVLink (exchange.cu); x_geo == nullptr: off
  float* x_geo;                 // this rank's packed gradient records [P][x_ngeo] (x_ngeo = 12: SH colours, 16: precomputed)
  int x_ngeo, x_mc, x_ndst, x_nranks;
  float* x_gc_dst[XCHG_MAX_RANKS];    // slot `rank` of the colour-gradient table: x_mc ? {multicast address} : one per rank
  float* x_gc_peer[XCHG_MAX_RANKS];   // the same slot through every rank's unicast mapping (ragged tail)
};
void launch_geom_backward(const BwdParams& p, const GeomState& g, cudaStream_t s);

// exchange.cu
// Layout of one rank's symmetric exchange buffer (identical on every rank; include/splat_b200.h: sfb_xchg):
//   [flags: 256 B][packed gradient records: P * ngeo floats][colour-gradient tables: 2 parities x world x P x 3 floats]
//   [flags 256 B][chunk flags: (XCHG_MAX_RANKS + 1) x nch words][packed records, 2 parities][colour tables, 2 parities]
// Chunk flags (fused backward + exchange, geom_bwd.cu): word r * nch + c = rank r has finished the geometry backward of
// chunk c (XCHG_CHUNK splats) in step `epoch`; word XCHG_MAX_RANKS * nch + c = the sums of chunk c have been broadcast.
// sfb_xchg_finish (exchange.cu) uses record parity 0 only.
constexpr int XCHG_CHUNK = 1024;
struct XchgLayout {
  size_t cflag_off, geo_off[2], gc_off[2], gc_slot_floats, bytes;
  int nch;
  static XchgLayout make(size_t P, int world, int ngeo, bool with_gc) {
    XchgLayout l;
    size_t o = 256;
    l.nch = (int)((P + XCHG_CHUNK - 1) / XCHG_CHUNK);
    l.cflag_off = o; o = align_up(o + (size_t)(XCHG_MAX_RANKS + 1) * (size_t)l.nch * 4, 256);
    for (int k = 0; k < 2; k++) { l.geo_off[k] = o; o = align_up(o + P * (size_t)ngeo * 4, 256); }
    l.gc_slot_floats = align_up(P * 3, 64);            // one view's [P][3] slot, padded to 256 bytes (16-byte stores)
    for (int k = 0; k < 2; k++) { l.gc_off[k] = o; if (with_gc) o = align_up(o + (size_t)world * l.gc_slot_floats * 4, 256); }
    l.bytes = o;
    return l;
  }
};
struct XchgDev {                 // device-side view of the exchange for one step
  int rank, world, P, ngeo;
  int nch;                       // chunks of XCHG_CHUNK splats
  uint32_t* flags;               // this rank's flag words
  uint32_t* peer_flags[XCHG_MAX_RANKS];
  uint32_t* cflags;              // this rank's chunk flags
  uint32_t* peer_cflags[XCHG_MAX_RANKS];
  float* geo;                    // this rank's packed records (sums after the exchange)
  float* geo_mc;                 // the same array through the multicast mapping, or nullptr
  float* peer_geo[XCHG_MAX_RANKS];
  const float* gc;               // this step's colour-gradient table: world slots of gc_slot_floats floats ([P][3] each)
  size_t gc_slot_floats;
};
void launch_xchg_finish(const XchgDev& x, int max_ctas, uint32_t epoch, int D, int M, const float* means3D, const float* campos,
                        float* dL_dmeans3D, float* dL_dopacity, float* dL_dscales, float* dL_drot, float* dL_dcolors,
                        float* dL_dsh, cudaStream_t s);
void xchg_tune(int nred_eighths, int depth);
// Fused geometry backward + exchange (geom_bwd.cu): ONE persistent kernel per step and rank.
struct FusedXchg {
  XchgDev x;                     // geo / geo_mc / peer_geo / gc point at THIS step's parity
  uint32_t epoch;
  int V, M;                      // views (= world), SH coefficients per colour channel
  const float* campos_views;     // [V][3] (with shs)
  float *dL_dmeans3D, *dL_dopacity, *dL_dscales, *dL_drot, *dL_dcolors, *dL_dsh;   // the SUMS over the ranks
};
// false: this configuration has no fused kernel (the caller runs launch_geom_backward + launch_xchg_finish instead)
bool launch_geom_exchange_fused(const BwdParams& p, const GeomState& g, const FusedXchg& f, int max_ctas, cudaStream_t s);
// dL_dsh[i] = sum over V views of basis(normalize(means3D[i] - campos[v])) (x) dcolor[v][i]   (view-parallel exchange)
void launch_sh_grad_combine(int P, int V, int D, int M, const float* means3D, const float* campos /* [V][3] */,
                            const float* dcolor /* [V][P][3] */, float* dL_dsh /* [P][M][3] */, bool wide256,
                            cudaStream_t s);

}  // namespace sfb
