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
#include <cstdlib>

namespace sfb {

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

// ------------------------------------------------------------------ onesweep radix sort (default)
// One kernel per digit pass: a block takes a ticket, ranks its 4096 (key, value) pairs stably, publishes
// its per-digit counts and obtains its global offsets by DECOUPLED LOOK-BACK over the predecessors'
// published counts (flag | value in one word: 1 = block aggregate, 2 = inclusive prefix), stages the pairs
// in shared memory in sorted order and writes them out as coalesced per-digit runs.  Keys and values are
// read once and written once per pass; the digit histograms of ALL passes come from one up-front sweep.
constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_INC = 2u << 30, OS_VAL_MASK = (1u << 30) - 1u;
constexpr int OS_MAX_PASSES = 4;
constexpr int OS_LB = 8;   // look-back probes in flight per digit (16 and 32 measured slower: 91 -> 103 -> 118 us depth sort)

// bias_c (optional): pointer to ~kmin, the complement of the smallest valid key (preprocess reduces it with
// atomicMax).  Keys are then rewritten IN PLACE as key - kmin (0xFFFFFFFF = culled -> 0: culled splats emit no
// instances, so their place in the order is irrelevant).  Depth keys of a bounded scene span < 2^24 after the
// shift, which makes the most significant digit pass an identity permutation (see onesweep_pass_kernel).
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_all_kernel(uint32_t* __restrict__ keys, int n, int npass, int4 shifts, int4 nbins,
                      uint32_t* __restrict__ hist_all /* [npass][SORT_MAX_BINS] */,
                      const uint32_t* __restrict__ bias_c) {
  // one private histogram per warp: 8x fewer same-address collisions on the shared-memory atomics
  constexpr int NW = SORT_THREADS / 32;
  __shared__ uint32_t s_h[NW][OS_MAX_PASSES][SORT_MAX_BINS];
  for (int i = threadIdx.x; i < NW * OS_MAX_PASSES * SORT_MAX_BINS; i += SORT_THREADS) (&s_h[0][0][0])[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int sh[4] = {shifts.x, shifts.y, shifts.z, shifts.w};
  const int nb[4] = {nbins.x, nbins.y, nbins.z, nbins.w};
  const int stride = gridDim.x * SORT_THREADS;
  const uint32_t kmin = bias_c ? ~(*bias_c) : 0u;
  // four keys per thread in flight: the loop is a load -> shared-atomic chain, and with one key per round trip the
  // kernel sat at 12-36 % issue-active waiting on the loads (ncu, round 2)
  constexpr int HU = 4;
  for (int k0 = blockIdx.x * SORT_THREADS + threadIdx.x; k0 < n; k0 += HU * stride) {
    uint32_t key[HU];
#pragma unroll
    for (int u = 0; u < HU; u++) { const int k = k0 + u * stride; key[u] = k < n ? keys[k] : 0u; }
#pragma unroll
    for (int u = 0; u < HU; u++) {
      const int k = k0 + u * stride;
      if (k < n) {
        uint32_t kk = key[u];
        if (bias_c) {
          kk = (kk == 0xFFFFFFFFu || kk < kmin) ? 0u : kk - kmin;
          keys[k] = kk;
        }
#pragma unroll
        for (int p = 0; p < OS_MAX_PASSES; p++)
          if (p < npass) atomicAdd(&s_h[warp][p][(kk >> sh[p]) & (uint32_t)(nb[p] - 1)], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * SORT_MAX_BINS; i += SORT_THREADS) {
    uint32_t v = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) v += (&s_h[w][0][0])[i];
    if (v) atomicAdd(&hist_all[i], v);
  }
}

// IPT items per thread: 16 (4096-item tiles) for large inputs, 4 (1024-item tiles) when 4096-item tiles
// would leave most SMs idle and every block a long latency chain.
// VM (value mode): 0 sorts bare 32-bit words (the packed tile|index instances: nothing but keys is staged or moved),
// 2 moves a 32-bit value with every key (the depth sort: depth bits -> Gaussian index), 1 moves ONE BYTE with every key
// (instances whose tile | index does not fit 32 bits: the index bits that do not fit ride along, 5 instead of 8 bytes
// per instance and pass) and, on the last pass (merge_bits > 0), re-assembles the full index
// (key & low mask) | byte << merge_bits  as the only output word: the render kernels read a plain index list.
// NB = digit bits the ranking resolves with ballots (6, 7 or 8 >= log2(bins)).
// For 4096-item tiles the IPT ranking rounds of a warp are split into CH independent chains (CH * bins <= SORT_MAX_BINS;
// chain c = items [c*IPT/CH, (c+1)*IPT/CH) of every lane) with their own digit counters, so the load -> add -> store ->
// shuffle dependency that serialises the rounds is IPT/CH long; the counters of the chains are stitched together by
// the per-digit exclusive prefix that already runs over the warps.
template <int IPT, int VM, int NB>
// Bare-key passes (the tile sort of packed instances) hold no value registers and run at 64 registers / 4 CTAs per SM
// (28-72 bytes of spills; 0.119 -> 0.116 ms at 9 M keys, 0.431 -> 0.415 ms at 37 M); with the side byte of split
// instances the same bound spills 128 bytes and gains nothing, so those and the pair sorts stay at 3.
__global__ void __launch_bounds__(SORT_THREADS, (VM == 0 && IPT == 16) ? 4 : 3)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const void* __restrict__ vals_in_v,
                     uint32_t* __restrict__ keys_out, void* __restrict__ vals_out_v, int merge_bits, int n, int shift, int bins,
                     const uint32_t* __restrict__ digit_totals /* [SORT_MAX_BINS] for this pass */,
                     uint32_t* __restrict__ tile_state /* [nblocks][bins], zeroed */,
                     uint32_t* __restrict__ ticket /* zeroed */,
                     uint2* __restrict__ ranges /* nullptr, or [T] = (0xFFFFFFFF, 0): fused K5, last tile-sort pass */,
                     int tile_shift /* tile id = key >> tile_shift */,
                     int skip_identity /* the consumers pick this pass's INPUT themselves when it is the identity */) {
  constexpr int NW = SORT_THREADS / 32;
  constexpr int OS_TILE = SORT_THREADS * IPT;
  constexpr int CH = IPT == 16 ? (NB == 6 ? 4 : (NB == 7 ? 2 : 1)) : 1;
  __shared__ uint32_t s_cnt[NW][SORT_MAX_BINS];   // per-warp (x chain) digit counts -> exclusive prefixes
  __shared__ uint32_t s_lstart[SORT_MAX_BINS];    // block-local start of each digit run
  __shared__ int32_t s_gofs[SORT_MAX_BINS];       // global position - local position, per digit
  __shared__ uint32_t s_key[OS_TILE];
  __shared__ uint32_t s_val[VM == 2 ? OS_TILE : (VM == 1 ? OS_TILE / 4 : 1)];
  constexpr bool HAS_VALS = VM != 0;
  const uint32_t* const vals_in = reinterpret_cast<const uint32_t*>(vals_in_v);
  const uint8_t* const vals_in8 = reinterpret_cast<const uint8_t*>(vals_in_v);
  uint32_t* const vals_out = reinterpret_cast<uint32_t*>(vals_out_v);
  uint8_t* const vals_out8 = reinterpret_cast<uint8_t*>(vals_out_v);
  uint8_t* const s_val8 = reinterpret_cast<uint8_t*>(s_val);
  const uint32_t merge_mask = (VM == 1 && merge_bits > 0) ? ((1u << merge_bits) - 1u) : 0u;
  __shared__ uint32_t s_wsum[NW];
  __shared__ uint32_t s_ticket;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t mask = (uint32_t)bins - 1;
  if (digit_totals[0] == (uint32_t)n) {
    // every key has digit 0 in this pass: the stable sort is the identity permutation -> plain coalesced copy
    // (or nothing at all when the consumers select the input buffer themselves: SortedIdx in common.cuh)
    if (skip_identity) return;
    const int base = blockIdx.x * OS_TILE;
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const int k = base + i * SORT_THREADS + threadIdx.x;
      if (k < n) {
        const uint32_t kk = keys_in[k];
        if (VM == 1 && merge_mask) keys_out[k] = (kk & merge_mask) | ((uint32_t)vals_in8[k] << merge_bits);
        else {
          keys_out[k] = kk;
          if (VM == 2) vals_out[k] = vals_in[k];
          if (VM == 1) vals_out8[k] = vals_in8[k];
        }
        if (ranges) {   // tile boundaries of an already sorted sequence: compare with the predecessor
          const uint32_t t = kk >> tile_shift;
          const uint32_t tp = k > 0 ? (keys_in[k - 1] >> tile_shift) : 0xFFFFFFFFu;
          if (tp != t) {
            atomicMin(&ranges[t].x, (uint32_t)k);
            if (k > 0) atomicMax(&ranges[tp].y, (uint32_t)k);
          }
          if (k == n - 1) atomicMax(&ranges[t].y, (uint32_t)n);
        }
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < NW * SORT_MAX_BINS; i += SORT_THREADS) (&s_cnt[0][0])[i] = 0;
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  const int blk = (int)s_ticket;

  // ---- stable ranking inside the block (sequential order = warp, item, lane) ----
  const int seg = blk * OS_TILE + warp * (32 * IPT);
  constexpr bool PRELOAD_VALS = HAS_VALS;     // measured: 88 vs 121 us per 9M-item pass without the early value loads
  uint32_t key[IPT], val[PRELOAD_VALS ? IPT : 1], rank[IPT];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < IPT; i++) {   // all loads of the tile in flight before the first use
    const int k = seg + i * 32 + lane;
    key[i] = k < n ? keys_in[k] : 0xFFFFFFFFu;
    if (PRELOAD_VALS) val[i] = k < n ? (VM == 1 ? (uint32_t)vals_in8[k] : vals_in[k]) : 0u;
  }
  // Round i ranks item i of every lane: lanes with equal digits find each other through a peer mask; the lowest
  // of them bumps the warp's private digit counter and broadcasts the old value.  All peer masks of the tile are
  // built first (independent work), the counter updates then run back to back.
  // (A returning shared-memory atomic instead of the load/store pair was measured slower on B200.)
  uint32_t peers_of[IPT];
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    const int k = seg + i * 32 + lane;
    const uint32_t d = (key[i] >> shift) & mask;
    {
      // MATCH.ANY costs ~300 cycles per warp on B200 with random 6-bit digits (its latency grows with the number of
      // distinct values among the lanes and it does not pipeline: ~110 Gkeys/s per pass whatever n; measured in round 1,
      // profiles/r01s_*).  One ballot per digit bit builds the same peer mask from 4 pipelined instructions per bit:
      // m &= ballot(bit) ^ (my bit ? 0 : ~0).
      uint32_t m = __ballot_sync(0xffffffffu, k < n);
#pragma unroll
      for (int b = 0; b < NB; b++) {
        asm("{\n\t.reg .pred p;\n\t.reg .b32 t, v, x;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
            "selp.b32 x, 0, 0xffffffff, p;\n\t"
            "lop3.b32 %0, %0, v, x, 0x60;\n\t}"          // m & (v ^ x)
            : "+r"(m) : "r"(d), "r"(1u << b));
      }
      peers_of[i] = k < n ? m : (1u << lane);
    }
  }
  constexpr int RPC = IPT / CH;   // rounds per chain
#pragma unroll
  for (int r = 0; r < RPC; r++) {
    uint32_t prevv[CH], before[CH];
    bool leader[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) {    // CH independent counter reads in flight
      const int i = c * RPC + r;
      const int k = seg + i * 32 + lane;
      const uint32_t d = (key[i] >> shift) & mask;
      before[c] = __popc(peers_of[i] & lt_mask);
      leader[c] = k < n && before[c] == 0;
      prevv[c] = leader[c] ? s_cnt[warp][c * bins + d] : 0u;
    }
#pragma unroll
    for (int c = 0; c < CH; c++) {
      const int i = c * RPC + r;
      const uint32_t d = (key[i] >> shift) & mask;
      if (leader[c]) s_cnt[warp][c * bins + d] = prevv[c] + __popc(peers_of[i]);
    }
#pragma unroll
    for (int c = 0; c < CH; c++) {
      const int i = c * RPC + r;
      rank[i] = __shfl_sync(0xffffffffu, prevv[c], __ffs(peers_of[i]) - 1) + before[c];
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit (one thread each): block count, publish the aggregate ----
  __shared__ uint32_t s_count[SORT_MAX_BINS];
  __shared__ uint32_t s_excl[SORT_MAX_BINS];
  __shared__ uint32_t s_wsum_t[NW];
  uint32_t my_count = 0, dtotal = 0;
  const int d = threadIdx.x;
  volatile uint32_t* st = tile_state;
  if (d < bins) {
    uint32_t run = 0;    // exclusive prefix in sequence order: warp-major, chain-minor
#pragma unroll
    for (int w = 0; w < NW; w++)
#pragma unroll
      for (int c = 0; c < CH; c++) { const uint32_t v = s_cnt[w][c * bins + d]; s_cnt[w][c * bins + d] = run; run += v; }
    my_count = run;
    s_count[d] = run;
    dtotal = digit_totals[d];
    st[(size_t)blk * bins + d] = (blk == 0 ? OS_FLAG_INC : OS_FLAG_AGG) | my_count;
  }
  // block-wide exclusive scans: of the block's digit counts (local starts) and of the digit totals (bases)
  uint32_t inc_c = my_count, inc_t = dtotal;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, inc_c, o), b2 = __shfl_up_sync(0xffffffffu, inc_t, o);
    if (lane >= o) { inc_c += a; inc_t += b2; }
  }
  if (lane == 31) { s_wsum[warp] = inc_c; s_wsum_t[warp] = inc_t; }
  __syncthreads();
  uint32_t wb_c = 0, wb_t = 0;
  for (int w = 0; w < warp; w++) { wb_c += s_wsum[w]; wb_t += s_wsum_t[w]; }
  const uint32_t lstart = wb_c + inc_c - my_count;     // local start of digit d inside the block
  const uint32_t dbase = wb_t + inc_t - dtotal;        // global start of digit d

  // ---- decoupled look-back, one thread per digit, OS_LB predecessors probed per round trip ----
  // (the probes of one round are independent loads, so a chain of k not-yet-inclusive predecessors costs
  //  k / OS_LB L2 latencies instead of k; partial progress is kept when a predecessor is not published yet)
  if (d < bins) {
    uint32_t excl = 0;
    if (blk > 0) {
      int p = blk - 1;
      while (true) {
        uint32_t v[OS_LB];
#pragma unroll
        for (int i = 0; i < OS_LB; i++) v[i] = (p - i >= 0) ? st[(size_t)(p - i) * bins + d] : (2u << 30);
        uint32_t add = 0;
        int adv = 0;
        bool fin = false;
#pragma unroll
        for (int i = 0; i < OS_LB; i++) {
          if (!fin && adv == i) {
            const uint32_t f = v[i] & ~OS_VAL_MASK;
            if (f != 0u) {
              add += v[i] & OS_VAL_MASK;
              adv++;
              fin = (f == (2u << 30));
            }
          }
        }
        excl += add;
        p -= adv;
        if (fin) break;
      }
      st[(size_t)blk * bins + d] = (2u << 30) | (excl + my_count);
    }
    s_excl[d] = excl;
  }
  __syncthreads();
  if (d < bins) {
    s_lstart[d] = lstart;
    s_gofs[d] = (int32_t)(dbase + s_excl[d]) - (int32_t)lstart;
  }
  __syncthreads();

  // ---- stage in shared memory in sorted order, then write coalesced runs ----
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    const int k = seg + i * 32 + lane;
    if (k < n) {
      const uint32_t dd = (key[i] >> shift) & mask;
      const uint32_t lp = s_lstart[dd] + s_cnt[warp][(i / RPC) * bins + dd] + rank[i];
      s_key[lp] = key[i];
      if (VM == 2) s_val[lp] = val[i];
      if (VM == 1) s_val8[lp] = (uint8_t)val[i];
    }
  }
  __syncthreads();
  const int cnt_blk = min(OS_TILE, n - blk * OS_TILE);
  for (int q = threadIdx.x; q < cnt_blk; q += SORT_THREADS) {
    const uint32_t kk = s_key[q];
    const uint32_t dd = (kk >> shift) & mask;
    const int32_t dst = (int32_t)q + s_gofs[dd];
    if (VM == 1 && merge_mask) keys_out[dst] = (kk & merge_mask) | ((uint32_t)s_val8[q] << merge_bits);
    else {
      keys_out[dst] = kk;
      if (VM == 2) vals_out[dst] = s_val[q];
      if (VM == 1) vals_out8[dst] = s_val8[q];
    }
    if (ranges) {
      // Fused K5 (identifyTileRanges) on the LAST pass of the tile sort: the block's staged items are in final order
      // inside each digit run, and a run is contiguous in the output.  Inside a run a tile boundary is seen by comparing
      // neighbours; what happens at the two ends of a run is unknown here (the neighbour belongs to another block or
      // digit), so the ends contribute with atomicMin / atomicMax: start = min, end = max over all contributions.
      const uint32_t t = kk >> tile_shift;
      uint32_t tp = 0xFFFFFFFFu, tn = 0xFFFFFFFFu;           // tile of the neighbour inside the same run, if any
      if (q > 0) { const uint32_t kp = s_key[q - 1]; if (((kp >> shift) & mask) == dd) tp = kp >> tile_shift; }
      if (q + 1 < cnt_blk) { const uint32_t kn = s_key[q + 1]; if (((kn >> shift) & mask) == dd) tn = kn >> tile_shift; }
      if (tp != t) atomicMin(&ranges[t].x, (uint32_t)dst);
      if (tn != t) atomicMax(&ranges[t].y, (uint32_t)dst + 1u);
    }
  }
}

static bool sort_small_tiles(int n) {
  return sort_blocks(n) < 64;
}

// words of `hist` that have to be zero when radix_sort_pairs starts (the caller may clear them itself, e.g. from a kernel
// that runs anyway, and pass scratch_zeroed = true)
size_t radix_sort_zero_words(int n, int nbits) {
  if (n <= 0 || nbits <= 0) return 0;
  const int npass = (nbits + 7) / 8;
  const int tile = sort_small_tiles(n) ? SORT_THREADS * 4 : SORT_THREADS * 16;
  const int nblocks = (n + tile - 1) / tile;
  size_t state_words = 0;
  int shift = 0;
  for (int pass = 0; pass < npass; pass++) {
    const int bits = (nbits - shift + (npass - pass) - 1) / (npass - pass);
    state_words += (size_t)nblocks << bits;
    shift += bits;
  }
  return (size_t)OS_MAX_PASSES * SORT_MAX_BINS + 8 + state_words;
}

static int radix_sort_pairs_onesweep(uint32_t* keys[2], uint32_t* vals[2], uint32_t* scratch, int n, int nbits,
                                     cudaStream_t s, int* launches, const char* const* names,
                                     const uint32_t* bias_c, int first_bit, bool scratch_zeroed, uint2* ranges,
                                     int tile_shift, uint8_t* const* vals8, int merge_bits,
                                     const uint32_t** last_total0) {
  const int vm = (vals8 && vals8[0]) ? 1 : ((vals != nullptr && vals[0] != nullptr) ? 2 : 0);
  const int npass = (nbits + 7) / 8;
  // 1024-item tiles only for really small inputs: with the ballot ranking 4096-item tiles win from ~0.3 M items
  // (measured: 1 M pairs 87 -> 65 us, 0.5 M 63 -> 54 us, 0.1 M 43 -> 51 us), although they fill < 2 CTAs per SM
  const bool small = sort_small_tiles(n);
  const int tile = small ? SORT_THREADS * 4 : SORT_THREADS * 16;
  const int nblocks = (n + tile - 1) / tile;
  // scratch layout: [hist_all: 4*256][tickets: 8][tile_state: npass * nblocks * bins]
  uint32_t* hist_all = scratch;
  uint32_t* tickets = scratch + OS_MAX_PASSES * SORT_MAX_BINS;
  uint32_t* state0 = tickets + 8;
  int shifts[4] = {0, 0, 0, 0}, nbins[4] = {1, 1, 1, 1};
  int shift = 0;
  size_t state_words = 0;
  for (int pass = 0; pass < npass; pass++) {
    const int bits = (nbits - shift + (npass - pass) - 1) / (npass - pass);
    shifts[pass] = first_bit + shift;
    nbins[pass] = 1 << bits;
    state_words += (size_t)nblocks * nbins[pass];
    shift += bits;
  }
  if (!scratch_zeroed)
    cudaMemsetAsync(scratch, 0, sizeof(uint32_t) * (OS_MAX_PASSES * SORT_MAX_BINS + 8 + state_words), s);
  prof_begin(names[0], s);
  const int hblocks = min(max(1, (n + 2047) / 2048), 4 * NUM_SMS_B200);   // >= 8 keys per thread, all SMs busy from ~0.3 M keys
  radix_hist_all_kernel<<<hblocks, SORT_THREADS, 0, s>>>(keys[0], n, npass, make_int4(shifts[0], shifts[1], shifts[2], shifts[3]),
                                                          make_int4(nbins[0], nbins[1], nbins[2], nbins[3]), hist_all, bias_c);
  prof_end(s);
  if (launches) *launches += 1;
  int cur = 0;
  uint32_t* state = state0;
  for (int pass = 0; pass < npass; pass++) {
    prof_begin(names[2], s);
    void* vin = vm == 2 ? (void*)vals[cur] : (vm == 1 ? (void*)vals8[cur] : nullptr);
    void* vout = vm == 2 ? (void*)vals[cur ^ 1] : (vm == 1 ? (void*)vals8[cur ^ 1] : nullptr);
    const int mb = (vm == 1 && pass == npass - 1) ? merge_bits : 0;
#define SFB_OS2(IPTV, HV, NBV)                                                                                 \
  onesweep_pass_kernel<IPTV, HV, NBV><<<nblocks, SORT_THREADS, 0, s>>>(keys[cur], vin, keys[cur ^ 1], vout, mb, n, \
                                                                        shifts[pass], nbins[pass],                \
                                                                        hist_all + pass * SORT_MAX_BINS, state, tickets + pass, \
                                                                        pass == npass - 1 ? ranges : nullptr, tile_shift, \
                                                                        (pass == npass - 1 && last_total0 && !ranges && vm != 1) ? 1 : 0)
#define SFB_OS(IPTV, HV)                                                                                       \
  do { if (nb == 6) SFB_OS2(IPTV, HV, 6); else if (nb == 7) SFB_OS2(IPTV, HV, 7); else SFB_OS2(IPTV, HV, 8); } while (0)
    const int nb = nbins[pass] <= 64 ? 6 : (nbins[pass] <= 128 ? 7 : 8);
    if (small) { if (vm == 2) SFB_OS(4, 2); else if (vm == 1) SFB_OS(4, 1); else SFB_OS(4, 0); }
    else       { if (vm == 2) SFB_OS(16, 2); else if (vm == 1) SFB_OS(16, 1); else SFB_OS(16, 0); }
#undef SFB_OS2
#undef SFB_OS
    prof_end(s);
    if (launches) *launches += 1;
    state += (size_t)nblocks * nbins[pass];
    cur ^= 1;
  }
  if (last_total0) *last_total0 = (!ranges && vm != 1) ? hist_all + (npass - 1) * SORT_MAX_BINS : nullptr;
  return cur;
}

int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t* hist, int n, int nbits, cudaStream_t s,
                     int* launches, const char* const* names, const uint32_t* bias_c, int first_bit,
                     bool scratch_zeroed, uint2* ranges, int tile_shift, uint8_t* const* vals8, int merge_bits,
                     const uint32_t** last_total0) {
  if (last_total0) *last_total0 = nullptr;
  if (n <= 0 || nbits <= 0) return 0;
  return radix_sort_pairs_onesweep(keys, vals, hist, n, nbits, s, launches, names, bias_c, first_bit, scratch_zeroed,
                                   ranges, tile_shift, vals8, merge_bits, last_total0);
}

// ------------------------------------------------------------------ instance emission in depth order
constexpr int DUP_THREADS = 256;   // DUP_GPB (Gaussians / depth ranks per block) lives in common.cuh: it sizes block_sums

__global__ void __launch_bounds__(DUP_THREADS)
instance_block_sums_kernel(int P, SortedIdx sorted,
                           const uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ block_sums) {
  const uint32_t* __restrict__ sorted_idx = sorted.get();
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

void launch_instance_block_sums(int P, const SortedIdx& sorted_idx, const uint32_t* tiles_touched,
                                uint32_t* block_sums, cudaStream_t s) {
  int nb = (P + DUP_GPB - 1) / DUP_GPB;
  prof_begin("instance_block_sums", s);
  instance_block_sums_kernel<<<nb, DUP_THREADS, 0, s>>>(P, sorted_idx, tiles_touched, block_sums);
  prof_end(s);
  prof_begin("instance_block_scan", s);
  scan_exclusive_kernel<<<1, 1024, 0, s>>>(block_sums, nb, nullptr);
  prof_end(s);
}

// Each block expands DUP_GPB depth-ranked Gaussians into their (tile, gaussian) instances.  Output slot k
// has to know which Gaussian it belongs to.  Instead of a per-slot binary search, the block works in rounds
// of DUP_CHUNK slots: every Gaussian that STARTS inside the round marks its first slot with its (local
// index + 1), the Gaussian spilling in from the previous round marks slot 0, and a block-wide running
// maximum over the slots (indices increase with the slot) turns the marks into a per-slot owner.  All
// global stores are then slot-strided, i.e. contiguous and coalesced no matter how skewed the counts are.
constexpr int DUP_CHUNK = 4096;
constexpr int DUP_SPT = DUP_CHUNK / DUP_THREADS;   // slots per thread in the scan (16)

__global__ void __launch_bounds__(DUP_THREADS)
duplicate_kernel(int P, int grid_x, SortedIdx sorted,
                 const uint32_t* __restrict__ tiles_touched, const uint2* __restrict__ rect,
                 const uint32_t* __restrict__ block_offsets, uint32_t* __restrict__ tile_keys,
                 uint8_t* __restrict__ inst_hi /* nullptr: the whole index fits the word; else index >> idx_bits */,
                 int idx_bits /* tile_keys[k] = tile << idx_bits | (index & ((1 << idx_bits) - 1)) */, uint32_t* __restrict__ zero_ptr, uint32_t zero_words, uint2* __restrict__ ranges_init,
                 int T, uint32_t* __restrict__ bcount_zero) {
  // Prologue: this kernel runs right in front of the tile sort anyway, so its blocks also clear the sort's scratch
  // (digit histograms, tickets, look-back state) and set the tile ranges to "empty" — two memset nodes less per forward.
  if (zero_ptr) {
    const uint32_t per = (zero_words + gridDim.x - 1) / gridDim.x;
    const uint32_t z0 = blockIdx.x * per, z1 = min(z0 + per, zero_words);
    for (uint32_t i = z0 + threadIdx.x; i < z1; i += DUP_THREADS) zero_ptr[i] = 0u;
  }
  if (bcount_zero && blockIdx.x == 0 && threadIdx.x < TILE_BUCKETS) bcount_zero[threadIdx.x] = 0u;
  if (ranges_init) {
    const int per = (T + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t0 = blockIdx.x * per, t1 = min(t0 + per, T);
    for (int i = t0 + threadIdx.x; i < t1; i += DUP_THREADS) ranges_init[i] = make_uint2(0xFFFFFFFFu, 0u);
  }
  const uint32_t* __restrict__ sorted_idx = sorted.get();
  // (Counting the tile sort's digit histograms here, where every key's tile is known, instead of a separate pass over the
  //  R keys was measured in round 2: the shared-memory atomics cost this kernel more than the histogram pass saves —
  //  +12 us against -17 us at 9 M keys, +73 against -64 us at 37 M; profiles/r02f_quick_perf_hist_in_duplicate.jsonl.)
  const bool fast_div = T <= (1 << 16);
  __shared__ uint32_t s_pref[DUP_GPB + 1];
  __shared__ uint32_t s_gidx[DUP_GPB];
  __shared__ uint2 s_rect[DUP_GPB];
  __shared__ uint32_t s_warp[DUP_THREADS / 32];
  __shared__ __align__(16) uint16_t s_own[DUP_CHUNK];
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
  uint32_t start[PER];
  {
    uint32_t ex = s_warp[warp] + inc - tsum;
#pragma unroll
    for (int k = 0; k < PER; k++) { start[k] = ex; s_pref[threadIdx.x * PER + k] = ex; ex += cnt[k]; }
  }
  __syncthreads();
  const uint32_t total = s_pref[DUP_GPB];
  // (the block offsets come from a separate per-block sum + scan: fusing that scan into this kernel with a
  //  decoupled look-back was measured SLOWER on B200 — 57 vs 53 us at 1M splats, 157 vs 132 us on dtu_500k)
  const uint32_t out0 = block_offsets[blockIdx.x];

  for (uint32_t cb = 0; cb < total; cb += DUP_CHUNK) {
    // A: clear the marks (two 16-byte stores per thread)
    reinterpret_cast<uint4*>(s_own)[threadIdx.x * 2] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(s_own)[threadIdx.x * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // B: Gaussians starting in this round mark their first slot
#pragma unroll
    for (int k = 0; k < PER; k++)
      if (cnt[k] > 0 && start[k] >= cb && start[k] < cb + DUP_CHUNK)
        s_own[start[k] - cb] = (uint16_t)(threadIdx.x * PER + k + 1);
    __syncthreads();
    if (threadIdx.x == 0 && s_own[0] == 0) {
      // the Gaussian that spills in: largest s with s_pref[s] <= cb (it has a non-zero count)
      int lo = 0, hi = DUP_GPB;
#pragma unroll
      for (int it = 0; (1 << it) < DUP_GPB; it++) {
        int mid = (lo + hi) >> 1;
        if (s_pref[mid] <= cb) lo = mid; else hi = mid;
      }
      s_own[0] = (uint16_t)(lo + 1);
    }
    __syncthreads();
    // C: running maximum over the slots; thread t owns slots [16t, 16t+16)
    uint32_t w8[DUP_SPT / 2];
    {
      const uint4 a4 = reinterpret_cast<const uint4*>(s_own)[threadIdx.x * 2];
      const uint4 b4 = reinterpret_cast<const uint4*>(s_own)[threadIdx.x * 2 + 1];
      w8[0] = a4.x; w8[1] = a4.y; w8[2] = a4.z; w8[3] = a4.w; w8[4] = b4.x; w8[5] = b4.y; w8[6] = b4.z; w8[7] = b4.w;
    }
    uint32_t m[DUP_SPT];
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < DUP_SPT; q++) {
      const uint32_t v = (q & 1) ? (w8[q >> 1] >> 16) : (w8[q >> 1] & 0xFFFFu);
      run = max(run, v);
      m[q] = run;
    }
    uint32_t incm = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incm, o);
      if (lane >= o) incm = max(incm, t);
    }
    if (lane == 31) s_warp[warp] = incm;
    uint32_t exm = __shfl_up_sync(0xffffffffu, incm, 1);
    if (lane == 0) exm = 0;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < DUP_THREADS / 32; w++)
      if (w < warp) exm = max(exm, s_warp[w]);
#pragma unroll
    for (int q = 0; q < DUP_SPT / 2; q++)
      w8[q] = max(exm, m[2 * q]) | (max(exm, m[2 * q + 1]) << 16);
    reinterpret_cast<uint4*>(s_own)[threadIdx.x * 2] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
    reinterpret_cast<uint4*>(s_own)[threadIdx.x * 2 + 1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
    __syncthreads();
    // D: slot-strided emission
    const uint32_t nslots = min((uint32_t)DUP_CHUNK, total - cb);
    for (uint32_t q = threadIdx.x; q < nslots; q += DUP_THREADS) {
      const uint32_t sidx = (uint32_t)s_own[q] - 1u;
      const uint32_t t = cb + q - s_pref[sidx];
      const uint2 r = s_rect[sidx];
      const uint32_t x0 = r.x & 0xFFFFu, y0 = r.x >> 16, x1 = r.y & 0xFFFFu;
      const uint32_t w = x1 - x0;
      // t / w without the ~20-instruction emulated integer division (this loop runs once per instance): (t + 0.5) / w sits
      // at least 0.5 / w away from an integer and the approximate quotient is off by < 1e-6 relative, so the truncation is
      // exact while t < 2^16 (t < tiles of one splat <= T; larger grids take the integer division)
      const uint32_t yy = fast_div ? (uint32_t)__fdividef((float)t + 0.5f, (float)w) : t / w;
      const uint32_t xx = t - yy * w;
      const uint32_t tile = (y0 + yy) * (uint32_t)grid_x + (x0 + xx);
      const uint32_t gi = s_gidx[sidx];
      if (inst_hi) {
        tile_keys[out0 + cb + q] = (tile << idx_bits) | (gi & ((1u << idx_bits) - 1u));
        inst_hi[out0 + cb + q] = (uint8_t)(gi >> idx_bits);
      } else {
        tile_keys[out0 + cb + q] = (tile << idx_bits) | gi;
      }
    }
    __syncthreads();
  }
}

void launch_duplicate(int P, int grid_x, const SortedIdx& sorted_idx, const uint32_t* tiles_touched,
                      const uint2* rect, const uint32_t* block_offsets, uint32_t* tile_keys,
                      uint8_t* inst_hi, int idx_bits, uint32_t* zero_ptr, size_t zero_words, uint2* ranges_init, int T,
                      uint32_t* bcount_zero, cudaStream_t s) {
  int nb = (P + DUP_GPB - 1) / DUP_GPB;
  duplicate_kernel<<<nb, DUP_THREADS, 0, s>>>(P, grid_x, sorted_idx, tiles_touched, rect, block_offsets,
                                              tile_keys, inst_hi, idx_bits, zero_ptr, (uint32_t)zero_words, ranges_init, T,
                                              bcount_zero);
}

// Inspection (parity tests): the reference's (tile << 32 | depth bits) keys and the index list from the sorted words.
// merged == false: words are tile << idx_bits | index.  merged == true (split instances): the words are bare indices
// and the tile of position i is the one whose [start, end) range holds i — one thread per tile fills its run.
__global__ void export_keys_kernel(int R, const uint32_t* __restrict__ words, int idx_bits,
                                   const SplatRec* __restrict__ rec, uint64_t* __restrict__ out_keys,
                                   uint32_t* __restrict__ out_list) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint32_t w = words[i];
  const uint32_t tile = w >> idx_bits, idx = w & ((1u << idx_bits) - 1u);
  if (out_keys) out_keys[i] = ((uint64_t)tile << 32) | __float_as_uint(rec[idx].depth);
  if (out_list) out_list[i] = idx;
}
__global__ void export_keys_merged_kernel(int T, const uint2* __restrict__ ranges, const uint32_t* __restrict__ words,
                                          const SplatRec* __restrict__ rec, uint64_t* __restrict__ out_keys,
                                          uint32_t* __restrict__ out_list) {
  const int t = blockIdx.x;
  const uint2 r = ranges[t];
  if (r.y <= r.x || r.x == 0xFFFFFFFFu) return;
  for (uint32_t i = r.x + threadIdx.x; i < r.y; i += blockDim.x) {
    const uint32_t idx = words[i];
    if (out_keys) out_keys[i] = ((uint64_t)t << 32) | __float_as_uint(rec[idx].depth);
    if (out_list) out_list[i] = idx;
  }
}

void launch_export_keys(int R, int T, const uint32_t* words, bool merged, int idx_bits, const uint2* ranges,
                        const SplatRec* rec, uint64_t* out_keys, uint32_t* out_list, cudaStream_t s) {
  if (R <= 0) return;
  if (merged) export_keys_merged_kernel<<<T, 256, 0, s>>>(T, ranges, words, rec, out_keys, out_list);
  else export_keys_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, words, idx_bits, rec, out_keys, out_list);
}

}  // namespace sfb
