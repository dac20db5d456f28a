// exchange.cu — view-parallel gradient exchange.
//
// The reference renders the views of one iteration serially and lets autograd add their per-splat gradients up
// (train.py:169-252).  With one view per GPU those sums have to cross NVLink once per step.  Two implementations of
// the same sum live here:
//
//  * sh_grad_combine_kernel — the rebuild step of the NCCL formulation (all-gather of the [P,3] colour gradients +
//    all-reduce of the [P,11] geometry gradients by torch.distributed; host_api exchange="factored").
//  * xchg_finish_kernel — the whole exchange in ONE kernel over symmetric (peer-mapped / NVSwitch-multicast) memory,
//    fed directly by the backward kernel (geom_backward_kernel<PUSH>, which has already written this rank's colour
//    gradients into every rank's table with multimem.st while it was computing):
//        0. cross-rank barrier on flags in the symmetric buffers (every rank's backward has finished),
//        1. this rank's 1/N slice of the packed geometry records is summed over the ranks INSIDE THE SWITCH
//           (multimem.ld_reduce.add.v4.f32) and broadcast back to every rank (multimem.st) — the NVLS all-reduce,
//           two 16-byte instructions per 16 bytes, no staging buffers, no ring;
//           without multicast support the same slice is summed from the peers' unicast mappings in rank order,
//        2. meanwhile the other CTAs rebuild the SH gradient rows  sum_v basis(dir_v) (x) gc_v  from the local table,
//        3. second barrier (every slice has been broadcast), then the records are unpacked into the caller's
//           per-parameter gradient arrays.
//    HBM-bound streaming work + 2 x 48 B/splat over NVLink per rank; nothing here goes through NCCL.
#include "common.cuh"
#include "xchg.cuh"

namespace sfb {


// ---------------------------------------------------------------------------------------------------------------
// Rebuild step of the NCCL formulation.  One thread per Gaussian, 12 + 12*V bytes in, 12*M out: HBM-bound.
template <int D, bool W256>
__global__ void __launch_bounds__(256)
sh_grad_combine_kernel(int P, int V, int M, const float* __restrict__ means3D, const float* __restrict__ campos,
                       const float* __restrict__ dcolor, float* __restrict__ dL_dsh) {
  __shared__ float s_cam[3 * 64];
  for (int k = threadIdx.x; k < 3 * V; k += blockDim.x) s_cam[k] = campos[k];
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  sh_row_rebuild<D, W256>((size_t)idx, (size_t)P * 3, V, M, means3D, s_cam, dcolor, dL_dsh);
}

void launch_sh_grad_combine(int P, int V, int D, int M, const float* means3D, const float* campos,
                            const float* dcolor, float* dL_dsh, bool wide256, cudaStream_t s) {
  if (P <= 0) return;
  const int blocks = (P + 255) / 256;
  const bool w256 = wide256 && (M * 12) % 32 == 0 && (reinterpret_cast<size_t>(dL_dsh) & 31) == 0;
#define SFB_SC(DD)                                                                                              \
  if (w256 && M == (DD + 1) * (DD + 1) && (3 * (DD + 1) * (DD + 1)) % 8 == 0)                                   \
    sh_grad_combine_kernel<DD, true><<<blocks, 256, 0, s>>>(P, V, M, means3D, campos, dcolor, dL_dsh);          \
  else                                                                                                          \
    sh_grad_combine_kernel<DD, false><<<blocks, 256, 0, s>>>(P, V, M, means3D, campos, dcolor, dL_dsh);
  switch (D) {
    case 0: SFB_SC(0) break;
    case 1: SFB_SC(1) break;
    case 2: SFB_SC(2) break;
    default: SFB_SC(3) break;
  }
#undef SFB_SC
}

// ---------------------------------------------------------------------------------------------------------------
// The exchange over symmetric memory.


template <int D, bool MC, bool HAS_SH, bool W256, int RD = 4>
__global__ void __launch_bounds__(256, 2)
xchg_finish_kernel(XchgDev x, uint32_t epoch, int nred_eighths, int V, int M, const float* __restrict__ means3D,
                   const float* __restrict__ campos, float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dopacity,
                   float* __restrict__ dL_dscales, float* __restrict__ dL_drot, float* __restrict__ dL_dcolors,
                   float* __restrict__ dL_dsh) {
  __shared__ float s_cam[3 * XCHG_MAX_RANKS];
  __shared__ uint32_t s_ticket;
  const int N = x.world;
  if (threadIdx.x == 0) xchg_mark(x.flags, 0, true);
  // ---- 0. announce (stream order: this rank's backward has completed) and wait for everybody
  if (blockIdx.x == 0 && threadIdx.x < N) {
    __threadfence_system();
    st_release_sys(x.peer_flags[threadIdx.x] + FLAG_A + x.rank, epoch);
  }
  if (HAS_SH) for (int k = threadIdx.x; k < 3 * V; k += blockDim.x) s_cam[k] = campos[k];
  {
    const bool ok = threadIdx.x < N ? spin_until(x.flags + FLAG_A + threadIdx.x, epoch) : true;
    if (!__syncthreads_and(ok)) {           // a peer never announced: give up (all threads of the CTA together)
      if (threadIdx.x == 0) atomicMax(x.flags + FLAG_ERR, 1u | (epoch << 8));
      return;
    }
  }
  if (threadIdx.x == 0) xchg_mark(x.flags, 1);

  // ---- 1. sum this rank's slice of the packed records over the ranks, broadcast the sums (in place)
  // The slice is split over the first `nred` CTAs; the rest start on the SH rows right away (they only need barrier A).
  const size_t C = (size_t)x.P * (size_t)(x.ngeo / 4);                  // 16-byte chunks of the record array
  const size_t c0 = C * (size_t)x.rank / (size_t)N, c1 = C * (size_t)(x.rank + 1) / (size_t)N;
  const int nred = HAS_SH ? min((int)gridDim.x, max(1, ((int)gridDim.x * nred_eighths) / 8)) : (int)gridDim.x;
  // (Two ranks were tried with direct sums — every rank reads the peer's records, or has them pushed into an inbox by
  // the peer's geometry backward, and adds them itself, 48 instead of 72 B per splat and direction and no second barrier:
  // SM-issued unicast peer loads ran at 140 GB/s and peer stores at 300 GB/s on B200 against ~585 GB/s through the
  // switch's multicast path, so reduce + broadcast through the switch is used for every world size;
  // profiles/r02d_bench_n2_*.)
  if ((int)blockIdx.x < nred) {
    const size_t stride = (size_t)nred * 256;
    for (size_t c = c0 + (size_t)blockIdx.x * 256 + threadIdx.x; c < c1; c += RD * stride) {
      float4 v[RD];
#pragma unroll
      for (int u = 0; u < RD; u++) {       // RD independent round trips through the switch in flight per thread
        const size_t cu = c + (size_t)u * stride;
        if (cu < c1) {
          if (MC) v[u] = mm_ld_reduce_add(x.geo_mc + 4 * cu);
          else {
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < N; r++) {  // rank order: every rank would get the same bits, and only one computes them
              const float4 t = ld_relaxed_sys_v4(x.peer_geo[r] + 4 * cu);
              v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < RD; u++) {
        const size_t cu = c + (size_t)u * stride;
        if (cu < c1) {
          if (MC) mm_st(x.geo_mc + 4 * cu, v[u]);
          else for (int r = 0; r < N; r++) *reinterpret_cast<float4*>(x.peer_geo[r] + 4 * cu) = v[u];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      xchg_mark(x.flags, 2);
      if (atomicAdd(x.flags + FLAG_DONE, 1u) == (uint32_t)nred - 1u) {    // last CTA of the reduction: tell every rank
        x.flags[FLAG_DONE] = 0u;                                           // (reset for the next launch)
        __threadfence_system();
        for (int r = 0; r < N; r++) st_release_sys(x.peer_flags[r] + FLAG_B + x.rank, epoch);
      }
    }
  }

  // ---- 2. SH rows from the local colour-gradient table (work queue of 256-splat chunks over all CTAs)
  if (HAS_SH) {
    const uint32_t nchunks = (uint32_t)((x.P + 255) / 256);
    while (true) {
      __syncthreads();
      if (threadIdx.x == 0) s_ticket = atomicAdd(x.flags + FLAG_TICKET, 1u);
      __syncthreads();
      const uint32_t t = s_ticket;
      if (t >= nchunks) break;
      const size_t i = (size_t)t * 256 + threadIdx.x;
      if (i < (size_t)x.P) sh_row_rebuild<D, W256>(i, x.gc_slot_floats, V, M, means3D, s_cam, x.gc, dL_dsh);
    }
    if (threadIdx.x == 0) xchg_mark(x.flags, 3);
  }

  // ---- 3. every slice has been broadcast: unpack the summed records into the per-parameter arrays
  {
    const bool ok = threadIdx.x < N ? spin_until(x.flags + FLAG_B + threadIdx.x, epoch) : true;
    if (!__syncthreads_and(ok)) {
      if (threadIdx.x == 0) atomicMax(x.flags + FLAG_ERR, 2u | (epoch << 8));
      return;
    }
  }
  if (threadIdx.x == 0) xchg_mark(x.flags, 4);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)x.P; i += (size_t)gridDim.x * 256) {
    const float4* rec = reinterpret_cast<const float4*>(x.geo + i * (size_t)x.ngeo);
    const float4 a = __ldcg(rec), b = __ldcg(rec + 1), c = __ldcg(rec + 2);
    dL_dmeans3D[3 * i] = a.x; dL_dmeans3D[3 * i + 1] = a.y; dL_dmeans3D[3 * i + 2] = a.z;
    dL_dopacity[i] = a.w;
    dL_dscales[3 * i] = b.x; dL_dscales[3 * i + 1] = b.y; dL_dscales[3 * i + 2] = b.z;
    dL_drot[4 * i] = b.w; dL_drot[4 * i + 1] = c.x; dL_drot[4 * i + 2] = c.y; dL_drot[4 * i + 3] = c.z;
    if (!HAS_SH) {
      const float4 d = __ldcg(rec + 3);
      dL_dcolors[3 * i] = d.x; dL_dcolors[3 * i + 1] = d.y; dL_dcolors[3 * i + 2] = d.z;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) xchg_mark(x.flags, 5);
}

// Tuning hooks of the measurement sessions (sfb_xchg_tune): share of the CTAs that start on the slice reduction
// (eighths of the grid) and reduction round trips in flight per thread (4 or 16).
// Default (0 = by world size): a quarter of the CTAs reduces with two ranks, an eighth with more — the reduction is
// bound by the links whatever the number of CTAs on it (123-135 us at N = 2 with 37 ... 296 CTAs), while the SH rows get
// instruction-bound as the views multiply and were the long pole at N = 8 with half of the grid on them (182 us,
// profiles/r02d_bench_n8_nvlink.json).
static int g_xchg_nred_eighths = 0, g_xchg_depth = 4;
void xchg_tune(int nred_eighths, int depth) {
  if (nred_eighths >= 0 && nred_eighths <= 8) g_xchg_nred_eighths = nred_eighths;
  if (depth == 4 || depth == 16) g_xchg_depth = depth;
}

void launch_xchg_finish(const XchgDev& x, int max_ctas, uint32_t epoch, int D, int M, const float* means3D, const float* campos,
                        float* dL_dmeans3D, float* dL_dopacity, float* dL_dscales, float* dL_drot, float* dL_dcolors,
                        float* dL_dsh, cudaStream_t s) {
  const bool has_sh = dL_dsh != nullptr;
  const bool mc = x.geo_mc != nullptr;
  const bool w256 = has_sh && (M * 12) % 32 == 0 && (reinterpret_cast<size_t>(dL_dsh) & 31) == 0 && M == (D + 1) * (D + 1) &&
                    (3 * (D + 1) * (D + 1)) % 8 == 0;
  cudaMemsetAsync(x.flags + FLAG_TICKET, 0, sizeof(uint32_t), s);      // the SH work queue's ticket
  cudaMemsetAsync(x.flags + FLAG_TL, 0, 6 * sizeof(unsigned long long), s);
  // 2 CTAs per SM (launch bound): the whole grid is resident, so CTAs spinning on a flag can never keep the CTAs
  // that produce this rank's own signals off the SMs
  const int grid = max_ctas > 0 ? min(max_ctas, 2 * NUM_SMS_B200) : 2 * NUM_SMS_B200;
  const int nred8 = g_xchg_nred_eighths > 0 ? g_xchg_nred_eighths : (x.world >= 4 ? 1 : 2);
#define SFB_XF(DD, MCV, SHV, WV)                                                                                    \
  do {                                                                                                              \
    if (g_xchg_depth == 16)                                                                                         \
      xchg_finish_kernel<DD, MCV, SHV, WV, 16><<<grid, 256, 0, s>>>(x, epoch, nred8, x.world, M, means3D, campos, \
                                                                    dL_dmeans3D, dL_dopacity, dL_dscales, dL_drot, dL_dcolors, dL_dsh); \
    else                                                                                                            \
      xchg_finish_kernel<DD, MCV, SHV, WV, 4><<<grid, 256, 0, s>>>(x, epoch, nred8, x.world, M, means3D, campos, \
                                                                   dL_dmeans3D, dL_dopacity, dL_dscales, dL_drot, dL_dcolors, dL_dsh); \
  } while (0)
#define SFB_XD(DD)                                                                                                  \
  do {                                                                                                              \
    if (mc) { if (w256) SFB_XF(DD, true, true, true); else SFB_XF(DD, true, true, false); }                         \
    else    { if (w256) SFB_XF(DD, false, true, true); else SFB_XF(DD, false, true, false); }                       \
  } while (0)
  if (!has_sh) { if (mc) SFB_XF(0, true, false, false); else SFB_XF(0, false, false, false); }
  else switch (D) {
    case 0: SFB_XD(0); break;
    case 1: SFB_XD(1); break;
    case 2: SFB_XD(2); break;
    default: SFB_XD(3); break;
  }
#undef SFB_XD
#undef SFB_XF
}

}  // namespace sfb
