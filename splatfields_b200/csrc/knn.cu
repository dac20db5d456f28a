// knn.cu — SURVEY.md §8f-5: replacement of simple_knn's distCUDA2 (un-vendored dependency, pinned at README.md:29:
// gitlab.inria.fr/bkerbl/simple-knn @ 44f7642; only call site scene/gaussian_model.py:105), which initialises the
// Gaussian scales from the mean squared distance of every point to its 3 nearest OTHER points.
//
// simple_knn sorts the points along a 30-bit Morton curve, cuts the order into boxes of 1024 points and lets every
// point test ALL P/1024 boxes, scanning the 1024 points of each box it cannot reject.  Same exact result here, but
// the boxes form an implicit 32-ary hierarchy over the Morton order — leaves of 32 points (= one warp of queries),
// 32 leaves per level-1 node, 32 level-1 nodes per level-2 node — so a query rejects 1024 / 32768 points with one
// box test and scans ~10 leaves instead of ~8 boxes of 1024.  The hierarchy adapts to the point density (unlike a
// uniform grid) and needs no host round trip: bbox reduction, Morton keys, the library's onesweep radix sort, box
// construction and the query all run back to back on the caller's stream.
//
// Exactness: a box is skipped only if its distance to the query exceeds an upper bound of the query's 3rd-neighbour
// distance; box and point distances are evaluated with the same fp32 expression fma(dz,dz, fma(dy,dy, dx*dx)), which
// is monotonic in |dx|, |dy|, |dz|, so box distance <= distance of every point inside, in fp32, always.  The result
// is therefore bit-identical to a brute-force scan with that expression (the oracle's so_knn3_mean_dist2).
#include "../../include/splat_b200.h"
#include "common.cuh"
#include <cfloat>

namespace sfb {

constexpr int KNN_FAN = 32;

struct __align__(16) KnnBox { float4 lo, hi; };

__device__ __forceinline__ uint32_t ordered_u32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_f32(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// mm[0..2] = min xyz, mm[3..5] = max xyz as order-preserving uint32 (pre-set to 0xFFFFFFFF / 0); NaNs are ignored
__global__ void __launch_bounds__(256) knn_bbox_kernel(int P, const float* __restrict__ pts, uint32_t* __restrict__ mm) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float v = pts[3 * (size_t)i + a];
      lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (lo[a] <= hi[a]) { atomicMin(&mm[a], ordered_u32(lo[a])); atomicMax(&mm[3 + a], ordered_u32(hi[a])); }
    }
  }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {   // 10 bits -> every third bit
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

__global__ void __launch_bounds__(256)
knn_morton_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ mm, uint32_t* __restrict__ keys,
                  uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t code = 0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float lo = ordered_f32(mm[a]), hi = ordered_f32(mm[3 + a]);
    const float ext = hi - lo;
    float t = ext > 0.f ? (pts[3 * (size_t)i + a] - lo) / ext * 1023.f : 0.f;
    t = fminf(fmaxf(t, 0.f), 1023.f);                          // (NaN -> 0)
    code |= spread10((uint32_t)t) << a;
  }
  keys[i] = code;
  vals[i] = (uint32_t)i;
}

// sorted points: (x, y, z, original index as bits)
__global__ void __launch_bounds__(256)
knn_gather_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, float4* __restrict__ sp) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P) return;
  const uint32_t i = order[j];
  sp[j] = make_float4(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], __uint_as_float(i));
}

// one warp per box: level 0 reduces 32 sorted points, the levels above reduce 32 child boxes
template <bool LEAF>
__global__ void __launch_bounds__(256)
knn_boxes_kernel(int n_out, int n_in, const float4* __restrict__ sp, const KnnBox* __restrict__ child,
                 KnnBox* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_out) return;
  const int j = w * KNN_FAN + lane;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (j < n_in) {
    if (LEAF) {
      const float4 q = sp[j];
      lo[0] = hi[0] = q.x; lo[1] = hi[1] = q.y; lo[2] = hi[2] = q.z;
      // a NaN coordinate must not poison the box (fminf / fmaxf below drop NaNs only on one side)
#pragma unroll
      for (int a = 0; a < 3; a++) if (lo[a] != lo[a]) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
    } else {
      const KnnBox b = child[j];
      lo[0] = b.lo.x; lo[1] = b.lo.y; lo[2] = b.lo.z; hi[0] = b.hi.x; hi[1] = b.hi.y; hi[2] = b.hi.z;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if (lane == 0) {
    KnnBox b;
    b.lo = make_float4(lo[0], lo[1], lo[2], 0.f);
    b.hi = make_float4(hi[0], hi[1], hi[2], 0.f);
    out[w] = b;
  }
}

__device__ __forceinline__ float dist2_expr(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
// squared distance from p to the box; same expression shape as the point distance (see the header comment).
// An empty box (lo > hi) yields a huge distance and is never visited.
__device__ __forceinline__ float box_dist2(const KnnBox& b, float px, float py, float pz) {
  const float dx = fmaxf(fmaxf(b.lo.x - px, px - b.hi.x), 0.f);
  const float dy = fmaxf(fmaxf(b.lo.y - py, py - b.hi.y), 0.f);
  const float dz = fmaxf(fmaxf(b.lo.z - pz, pz - b.hi.z), 0.f);
  return dist2_expr(dx, dy, dz);
}
__device__ __forceinline__ void knn_insert(float best[3], float d) {   // simple_knn's updateKBest<3>
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
  }
}

// One warp per leaf: its 32 lanes are the 32 queries of that leaf, which are neighbours in space, so they want nearly
// the same boxes.  The warp walks the hierarchy together — a box is opened when ANY lane cannot reject it (ballot) —
// which keeps control flow uniform, turns every box read into a broadcast and every leaf read into one coalesced 512-byte
// load whose points then go round by shuffle.  A lane that did not need a box just sees extra candidates; a candidate
// can never make a best-3 list wrong.
__global__ void __launch_bounds__(128)
knn_query_kernel(int P, const float4* __restrict__ sp, const KnnBox* __restrict__ L0, int n0,
                 const KnnBox* __restrict__ L1, int n1, const KnnBox* __restrict__ L2, int n2,
                 float* __restrict__ mean_dist2) {
  const int leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (leaf >= n0) return;                                    // whole warps only
  const int i = leaf * KNN_FAN + lane;
  const bool live = i < P;
  const float4 p = live ? __ldg(sp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int cnt_own = min(KNN_FAN, P - leaf * KNN_FAN);
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  // upper bound of the 3rd-neighbour distance from the query's own leaf: its points are held by the other lanes
  for (int t = 0; t < cnt_own; t++) {
    const float qx = __shfl_sync(0xffffffffu, p.x, t), qy = __shfl_sync(0xffffffffu, p.y, t),
                qz = __shfl_sync(0xffffffffu, p.z, t);
    if (t != lane) knn_insert(best, dist2_expr(qx - p.x, qy - p.y, qz - p.z));
  }
  const float reject = best[2];
  best[0] = best[1] = best[2] = FLT_MAX;
  for (int a = 0; a < n2; a++) {
    if (!__any_sync(0xffffffffu, live && box_dist2(L2[a], p.x, p.y, p.z) <= fminf(reject, best[2]))) continue;
    for (int b = a * KNN_FAN; b < min(n1, (a + 1) * KNN_FAN); b++) {
      if (!__any_sync(0xffffffffu, live && box_dist2(L1[b], p.x, p.y, p.z) <= fminf(reject, best[2]))) continue;
      for (int c = b * KNN_FAN; c < min(n0, (b + 1) * KNN_FAN); c++) {
        if (!__any_sync(0xffffffffu, live && box_dist2(L0[c], p.x, p.y, p.z) <= fminf(reject, best[2]))) continue;
        const int j0 = c * KNN_FAN, cnt = min(KNN_FAN, P - j0);
        const float4 mine = lane < cnt ? __ldg(sp + j0 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < cnt; t++) {
          const float qx = __shfl_sync(0xffffffffu, mine.x, t), qy = __shfl_sync(0xffffffffu, mine.y, t),
                      qz = __shfl_sync(0xffffffffu, mine.z, t);
          if (j0 + t != i) knn_insert(best, dist2_expr(qx - p.x, qy - p.y, qz - p.z));
        }
      }
    }
  }
  if (live) mean_dist2[__float_as_uint(p.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

struct KnnScratch {
  uint32_t* mm;          // [8]
  uint32_t* keys[2];     // [P] Morton codes (ping-pong)
  uint32_t* vals[2];     // [P] point indices (ping-pong)
  uint32_t* sort;        // sort scratch
  float4* sp;            // [P] sorted points
  KnnBox *L0, *L1, *L2;
  int n0, n1, n2;
  static KnnScratch carve_from(char*& p, size_t P) {
    KnnScratch k;
    k.n0 = (int)((P + KNN_FAN - 1) / KNN_FAN);
    k.n1 = (k.n0 + KNN_FAN - 1) / KNN_FAN;
    k.n2 = (k.n1 + KNN_FAN - 1) / KNN_FAN;
    k.mm = carve<uint32_t>(p, 8);
    for (int i = 0; i < 2; i++) k.keys[i] = carve<uint32_t>(p, P);
    for (int i = 0; i < 2; i++) k.vals[i] = carve<uint32_t>(p, P);
    k.sort = carve<uint32_t>(p, sort_scratch_words(P));
    k.sp = carve<float4>(p, P);
    k.L0 = carve<KnnBox>(p, (size_t)k.n0);
    k.L1 = carve<KnnBox>(p, (size_t)k.n1);
    k.L2 = carve<KnnBox>(p, (size_t)k.n2);
    return k;
  }
};

}  // namespace sfb

extern "C" {

size_t sfb_knn_scratch_bytes(int P) {
  using namespace sfb;
  if (P <= 0) return 256;
  char* p = nullptr;
  KnnScratch::carve_from(p, (size_t)P);
  return (size_t)(p - (char*)nullptr) + 256;
}

int sfb_knn3_mean_dist2(int P, const float* points, float* mean_dist2, void* scratch, void* stream) {
  using namespace sfb;
  if (P < 0 || (P > 0 && (!points || !mean_dist2 || !scratch)))
    return set_error("sfb_knn3_mean_dist2: bad arguments"), SFB_ERR_ARG;
  if ((reinterpret_cast<size_t>(scratch) & 255) != 0)
    return set_error("sfb_knn3_mean_dist2: scratch must be 256-byte aligned"), SFB_ERR_ARG;
  if (P == 0) return SFB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  char* chunk = (char*)scratch;
  KnnScratch k = KnnScratch::carve_from(chunk, (size_t)P);
  cudaMemsetAsync(k.mm, 0xFF, 3 * sizeof(uint32_t), s);
  cudaMemsetAsync(k.mm + 3, 0, 3 * sizeof(uint32_t), s);
  const int pb = (P + 255) / 256;
  prof_begin("knn.bbox", s);
  knn_bbox_kernel<<<min(pb, 8 * NUM_SMS_B200), 256, 0, s>>>(P, points, k.mm);
  prof_end(s);
  prof_begin("knn.morton", s);
  knn_morton_kernel<<<pb, 256, 0, s>>>(P, points, k.mm, k.keys[0], k.vals[0]);
  prof_end(s);
  static const char* const names[3] = {"knn.sort.hist", "knn.sort.scan", "knn.sort.scatter"};
  const int cur = radix_sort_pairs(k.keys, k.vals, k.sort, P, 30, s, nullptr, names);
  prof_begin("knn.boxes", s);
  knn_gather_kernel<<<pb, 256, 0, s>>>(P, points, k.vals[cur], k.sp);
  knn_boxes_kernel<true><<<(k.n0 * 32 + 255) / 256, 256, 0, s>>>(k.n0, P, k.sp, nullptr, k.L0);
  knn_boxes_kernel<false><<<(k.n1 * 32 + 255) / 256, 256, 0, s>>>(k.n1, k.n0, nullptr, k.L0, k.L1);
  knn_boxes_kernel<false><<<(k.n2 * 32 + 255) / 256, 256, 0, s>>>(k.n2, k.n1, nullptr, k.L1, k.L2);
  prof_end(s);
  prof_begin("knn.query", s);
  knn_query_kernel<<<(k.n0 * 32 + 127) / 128, 128, 0, s>>>(P, k.sp, k.L0, k.n0, k.L1, k.n1, k.L2, k.n2, mean_dist2);
  prof_end(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e)), SFB_ERR_CUDA;
  return SFB_OK;
}

}  // extern "C"
