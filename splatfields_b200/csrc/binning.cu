// binning.cu — K2..K5: instance scan, duplicate-with-keys, stable radix sort, tile ranges.
//
// The external rasterizer sorts R 64-bit keys
// (tile << 32 | depth bits) with ~6 radix passes.  Here the same ORDER is produced with far less
// traffic by splitting the key:
//   1. stable sort of the P Gaussians by their 32 depth bits            (P pairs, 4 passes)
//   2. emit the (tile, gaussian) instances in that depth order            (load-balanced, coalesced)
//   3. stable sort of the R instances by tile id only                     (ceil(log2 T) bits, 2 passes)
// A stable sort by tile of a depth-ordered sequence is ordered by (tile, depth, gaussian index) — exactly
// the order the 64-bit stable sort yields (ties in (tile, depth) keep emission order = ascending index).
// The u64 key buffer is only materialised on request (sfb_export_binning) for the bit-exact parity check.
//
// All of this is integer work bound by HBM/L2 traffic; no tensor cores.
#include "common.cuh"

namespace sfb {

// ------------------------------------------------------------------ radix sort, one pass = 3 kernels
// Block b owns items [b*SORT_TILE, (b+1)*SORT_TILE).  Within a block, warp w owns a contiguous
// segment of 32*SORT_IPT items and item i of lane l sits at seg + i*32 + l (coalesced, and the
// sequential order inside the block is (warp, i, lane)).

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift, int bins, int nblocks,
                  uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_hist[SORT_MAX_BINS];
  for (int i = threadIdx.x; i < bins; i += SORT_THREADS) s_hist[i] = 0;
  __syncthreads();
  const uint32_t mask = (uint32_t)bins - 1;
  const int base = blockIdx.x * SORT_TILE;
#pragma unroll 4
  for (int i = 0; i < SORT_IPT; i++) {
    int k = base + i * SORT_THREADS + threadIdx.x;
    if (k < n) atomicAdd(&s_hist[(keys[k] >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += SORT_THREADS) hist[(size_t)i * nblocks + blockIdx.x] = s_hist[i];
}

// Exclusive scan, in place, of gridDim.x independent rows of `n` words (row r at data + r*n), one block
// of 1024 threads per row; the row total goes to total_out[r] (if non-null).
__global__ void __launch_bounds__(1024) scan_exclusive_kernel(uint32_t* __restrict__ data_all, int n,
                                                              uint32_t* __restrict__ total_out) {
  uint32_t* __restrict__ data = data_all + (size_t)blockIdx.x * n;
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {
    int i0 = base + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? data[i0 + k] : 0u;
    uint32_t tsum = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
      uint32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - w;  // exclusive prefix of warp sums
    }
    __syncthreads();
    uint32_t carry = s_carry;
    uint32_t ex = carry + s_warp[warp] + (inc - tsum);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < n) data[i0 + k] = ex;
      ex += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = ex;  // ex == carry + everything in this tile
    __syncthreads();
  }
  if (total_out && threadIdx.x == 0) total_out[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
                     int bins, int nblocks, const uint32_t* __restrict__ hist_scanned,
                     const uint32_t* __restrict__ digit_totals) {
  constexpr int NW = SORT_THREADS / 32;
  __shared__ uint32_t s_cnt[NW][SORT_MAX_BINS];  // per-warp digit counters, later per-warp global bases
  __shared__ uint32_t s_dbase[SORT_MAX_BINS];    // exclusive scan of the digit totals
  __shared__ uint32_t s_wsum[NW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t mask = (uint32_t)bins - 1;
  for (int i = threadIdx.x; i < NW * SORT_MAX_BINS; i += SORT_THREADS) (&s_cnt[0][0])[i] = 0;
  {  // block-wide exclusive scan of digit_totals[0..bins) (bins <= SORT_THREADS)
    uint32_t t = (int)threadIdx.x < bins ? digit_totals[threadIdx.x] : 0u;
    uint32_t inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t wb = 0;
    for (int w = 0; w < warp; w++) wb += s_wsum[w];
    s_dbase[threadIdx.x] = wb + inc - t;
  }
  __syncthreads();

  const int seg = blockIdx.x * SORT_TILE + warp * (32 * SORT_IPT);
  uint32_t key[SORT_IPT], rank[SORT_IPT];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < SORT_IPT; i++) {
    int k = seg + i * 32 + lane;
    bool valid = k < n;
    key[i] = valid ? keys_in[k] : 0xFFFFFFFFu;
    uint32_t d = (key[i] >> shift) & mask;
    // invalid lanes use a digit no valid lane can match (bins <= 256)
    uint32_t md = valid ? d : 0xFFFFu;
    uint32_t peers = __match_any_sync(0xffffffffu, md);
    uint32_t before = __popc(peers & lt_mask);
    uint32_t prev = 0;
    if (valid && before == 0) {  // leader of its digit group
      prev = s_cnt[warp][d];
      s_cnt[warp][d] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
    rank[i] = prev + before;
    __syncwarp();
  }
  __syncthreads();
  // per digit: turn per-warp counts into global bases (exclusive over warps + block's scanned base)
  for (int d = threadIdx.x; d < bins; d += SORT_THREADS) {
    uint32_t run = s_dbase[d] + hist_scanned[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < NW; w++) {
      uint32_t c = s_cnt[w][d];
      s_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_IPT; i++) {
    int k = seg + i * 32 + lane;
    if (k < n) {
      uint32_t d = (key[i] >> shift) & mask;
      uint32_t dst = s_cnt[warp][d] + rank[i];
      keys_out[dst] = key[i];
      vals_out[dst] = vals_in[k];
    }
  }
}

int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t* hist, int n, int nbits, cudaStream_t s,
                     int* launches, const char* const* names) {
  if (n <= 0 || nbits <= 0) return 0;
  const int npass = (nbits + 7) / 8;
  const int nblocks = sort_blocks(n);
  int cur = 0, shift = 0;
  for (int pass = 0; pass < npass; pass++) {
    // spread the bits evenly over the passes (12 -> 6+6, 13 -> 7+6, 32 -> 8x4)
    int bits = (nbits - shift + (npass - pass) - 1) / (npass - pass);
    int bins = 1 << bits;
    prof_begin(names[0], s);
    radix_hist_kernel<<<nblocks, SORT_THREADS, 0, s>>>(keys[cur], n, shift, bins, nblocks, hist);
    prof_end(s);
    uint32_t* totals = hist + (size_t)SORT_MAX_BINS * nblocks;
    prof_begin(names[1], s);
    scan_exclusive_kernel<<<bins, 1024, 0, s>>>(hist, nblocks, totals);  // one row (digit) per block
    prof_end(s);
    prof_begin(names[2], s);
    radix_scatter_kernel<<<nblocks, SORT_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n,
                                                          shift, bins, nblocks, hist, totals);
    prof_end(s);
    if (launches) *launches += 3;
    cur ^= 1;
    shift += bits;
  }
  return cur;
}

// ------------------------------------------------------------------ instance emission in depth order
constexpr int DUP_THREADS = 256;
constexpr int DUP_GPB = 1024;  // Gaussians (depth ranks) per block

__global__ void __launch_bounds__(DUP_THREADS)
instance_block_sums_kernel(int P, const uint32_t* __restrict__ sorted_idx,
                           const uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t s_w[DUP_THREADS / 32];
  uint32_t v = 0;
  for (int j = blockIdx.x * DUP_GPB + threadIdx.x; j < min(P, (blockIdx.x + 1) * DUP_GPB); j += DUP_THREADS)
    v += tiles_touched[sorted_idx[j]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < DUP_THREADS / 32; w++) t += s_w[w];
    block_sums[blockIdx.x] = t;
  }
}

void launch_instance_block_sums(int P, const uint32_t* sorted_idx, const uint32_t* tiles_touched,
                                uint32_t* block_sums, cudaStream_t s) {
  int nb = (P + DUP_GPB - 1) / DUP_GPB;
  prof_begin("instance_block_sums", s);
  instance_block_sums_kernel<<<nb, DUP_THREADS, 0, s>>>(P, sorted_idx, tiles_touched, block_sums);
  prof_end(s);
  prof_begin("instance_block_scan", s);
  scan_exclusive_kernel<<<1, 1024, 0, s>>>(block_sums, nb, nullptr);
  prof_end(s);
}

// Each block expands DUP_GPB depth-ranked Gaussians.  Output slot k of the block is produced by the
// thread that owns k, which finds its source Gaussian by binary search over the block's exclusive
// scan of tile counts: all stores are contiguous and coalesced no matter how skewed the counts are.
__global__ void __launch_bounds__(DUP_THREADS)
duplicate_kernel(int P, int grid_x, const uint32_t* __restrict__ sorted_idx,
                 const uint32_t* __restrict__ tiles_touched, const uint2* __restrict__ rect,
                 const uint32_t* __restrict__ block_offsets, uint32_t* __restrict__ tile_keys,
                 uint32_t* __restrict__ inst_idx) {
  __shared__ uint32_t s_pref[DUP_GPB + 1];
  __shared__ uint32_t s_gidx[DUP_GPB];
  __shared__ uint2 s_rect[DUP_GPB];
  __shared__ uint32_t s_warp[DUP_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * DUP_GPB;
  constexpr int PER = DUP_GPB / DUP_THREADS;  // consecutive ranks per thread
  uint32_t cnt[PER];
  uint32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < PER; k++) {
    int local = threadIdx.x * PER + k;
    int j = j0 + local;
    uint32_t c = 0;
    if (j < P) {
      uint32_t gi = sorted_idx[j];
      c = tiles_touched[gi];
      s_gidx[local] = gi;
      s_rect[local] = rect[gi];
    }
    cnt[k] = c;
    tsum += c;
  }
  uint32_t inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int w = 0; w < DUP_THREADS / 32; w++) { uint32_t c = s_warp[w]; s_warp[w] = run; run += c; }
    s_pref[DUP_GPB] = run;
  }
  __syncthreads();
  uint32_t ex = s_warp[warp] + inc - tsum;
#pragma unroll
  for (int k = 0; k < PER; k++) { s_pref[threadIdx.x * PER + k] = ex; ex += cnt[k]; }
  __syncthreads();
  const uint32_t total = s_pref[DUP_GPB];
  const uint32_t out0 = block_offsets[blockIdx.x];
  for (uint32_t k = threadIdx.x; k < total; k += DUP_THREADS) {
    // largest s with s_pref[s] <= k
    int lo = 0, hi = DUP_GPB;
#pragma unroll
    for (int it = 0; it < 10; it++) {  // log2(DUP_GPB)
      int mid = (lo + hi) >> 1;
      if (s_pref[mid] <= k) lo = mid; else hi = mid;
    }
    uint32_t t = k - s_pref[lo];
    uint2 r = s_rect[lo];
    uint32_t x0 = r.x & 0xFFFFu, y0 = r.x >> 16, x1 = r.y & 0xFFFFu;
    uint32_t w = x1 - x0;
    uint32_t yy = t / w, xx = t - yy * w;
    tile_keys[out0 + k] = (y0 + yy) * (uint32_t)grid_x + (x0 + xx);
    inst_idx[out0 + k] = s_gidx[lo];
  }
}

void launch_duplicate(int P, int grid_x, const uint32_t* sorted_idx, const uint32_t* tiles_touched,
                      const uint2* rect, const uint32_t* block_offsets, uint32_t* tile_keys,
                      uint32_t* inst_idx, cudaStream_t s) {
  int nb = (P + DUP_GPB - 1) / DUP_GPB;
  duplicate_kernel<<<nb, DUP_THREADS, 0, s>>>(P, grid_x, sorted_idx, tiles_touched, rect, block_offsets,
                                              tile_keys, inst_idx);
}

// ------------------------------------------------------------------ tile ranges
__global__ void tile_ranges_kernel(int R, const uint32_t* __restrict__ keys, uint2* __restrict__ ranges) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  uint32_t t = keys[i];
  if (i == 0) ranges[t].x = 0;
  else {
    uint32_t prev = keys[i - 1];
    if (prev != t) { ranges[prev].y = (uint32_t)i; ranges[t].x = (uint32_t)i; }
  }
  if (i == R - 1) ranges[t].y = (uint32_t)R;
}

void launch_tile_ranges(int R, int T, const uint32_t* sorted_tile_keys, uint2* ranges, cudaStream_t s) {
  cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)T, s);
  if (R > 0) tile_ranges_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, sorted_tile_keys, ranges);
}

__global__ void export_keys_kernel(int R, const uint32_t* __restrict__ tile_keys,
                                   const uint32_t* __restrict__ point_list, const SplatRec* __restrict__ rec,
                                   uint64_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  uint32_t d = __float_as_uint(rec[point_list[i]].depth);
  out[i] = ((uint64_t)tile_keys[i] << 32) | d;
}

void launch_export_keys(int R, const uint32_t* tile_keys, const uint32_t* point_list, const SplatRec* rec,
                        uint64_t* out_keys, cudaStream_t s) {
  if (R > 0) export_keys_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, tile_keys, point_list, rec, out_keys);
}

}  // namespace sfb
