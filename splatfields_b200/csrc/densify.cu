// densify.cu — SURVEY.md §8f-2: the per-iteration consumers of the rasterizer's side outputs
// (viewspace_points.grad, radii, visibility_filter), fused into one pass each.
//
//   sfb_densify_stats   scene/gaussian_model.py:427-430 (add_densification_stats) + train.py:280-282 (max_radii2D):
//                       the reference runs ~10 boolean-mask indexing kernels (each with a nonzero() + host sync);
//   sfb_densify_masks   the selection predicates of densify_and_prune / densify_and_clone / densify_and_split
//                       (scene/gaussian_model.py:355-425): grads = accum / denom with NaN -> 0, clone / split / prune masks
//                       and their counts.  The optimizer-state surgery that follows stays with the caller (out of scope).
// Pure streaming work, 28-41 B per Gaussian; bound by HBM.
#include "../../include/splat_b200.h"
#include "common.cuh"

namespace sfb {

__global__ void __launch_bounds__(256)
densify_stats_kernel(int P, const float* __restrict__ g2d /* [P][3] */, const int* __restrict__ radii,
                     const uint8_t* __restrict__ filter /* or nullptr: radii > 0 */,
                     float* __restrict__ accum, float* __restrict__ denom, float* __restrict__ max_radii2D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int r = radii ? radii[i] : 0;
  const bool on = filter ? (filter[i] != 0) : (r > 0);
  if (!on) return;
  const float gx = g2d[3 * (size_t)i], gy = g2d[3 * (size_t)i + 1];
  // torch.norm(grad[:, :2], dim=-1) on CUDA: sqrt(fl(gx*gx) + fl(gy*gy)), products rounded separately (checked bit for
  // bit on a B200 over 1M random rows, scripts/diag_norm.py; torch's CPU kernel contracts to fma(gy, gy, gx*gx) instead)
  accum[i] += sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
  denom[i] += 1.f;
  if (max_radii2D && radii) max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
densify_masks_kernel(int P, const float* __restrict__ accum, const float* __restrict__ denom,
                     const float* __restrict__ scales /* [P][3] */, const float* __restrict__ opacity /* [P] */,
                     const float* __restrict__ max_radii2D /* or nullptr */, int raw, float grad_threshold,
                     float dense_extent /* percent_dense * extent */, float min_opacity, float max_screen_size,
                     float big_ws /* 0.1 * extent */, uint8_t* __restrict__ clone, uint8_t* __restrict__ split,
                     uint8_t* __restrict__ prune, uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_c[3];
  if (threadIdx.x < 3) s_c[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool c = false, s = false, p = false;
  if (i < P) {
    float g = accum[i] / denom[i];
    if (g != g) g = 0.f;                                          // grads[grads.isnan()] = 0.0
    float s0 = scales[3 * (size_t)i], s1 = scales[3 * (size_t)i + 1], s2 = scales[3 * (size_t)i + 2];
    float op = opacity[i];
    if (raw) { s0 = expf(s0); s1 = expf(s1); s2 = expf(s2); op = sigmoidf_(op); }   // gaussian_model.py:53-58
    const float smax = fmaxf(s0, fmaxf(s1, s2));
    const bool hot = fabsf(g) >= grad_threshold;
    c = hot && smax <= dense_extent;                              // :397-400
    s = (g >= grad_threshold) && smax > dense_extent;             // :360-363
    p = op < min_opacity;                                         // :418
    if (max_screen_size > 0.f) {                                  // :419-422 (`if max_screen_size:`)
      p = p || (max_radii2D && max_radii2D[i] > max_screen_size) || smax > big_ws;
    }
    clone[i] = c; split[i] = s; prune[i] = p;
  }
  const uint32_t bc = __ballot_sync(0xffffffffu, c), bs = __ballot_sync(0xffffffffu, s), bp = __ballot_sync(0xffffffffu, p);
  if ((threadIdx.x & 31) == 0) {
    if (bc) atomicAdd(&s_c[0], __popc(bc));
    if (bs) atomicAdd(&s_c[1], __popc(bs));
    if (bp) atomicAdd(&s_c[2], __popc(bp));
  }
  __syncthreads();
  if (counts && threadIdx.x < 3 && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_c[threadIdx.x]);
}

}  // namespace sfb

extern "C" {

int sfb_densify_stats(int P, const float* dL_dmeans2D, const int* radii, const uint8_t* update_filter,
                      float* xyz_gradient_accum, float* denom, float* max_radii2D, void* stream) {
  using namespace sfb;
  if (P < 0 || (P > 0 && (!dL_dmeans2D || !xyz_gradient_accum || !denom || (!radii && !update_filter))))
    return set_error("sfb_densify_stats: bad arguments"), SFB_ERR_ARG;
  if (P == 0) return SFB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("densify.stats", s);
  densify_stats_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, dL_dmeans2D, radii, update_filter, xyz_gradient_accum, denom,
                                                       max_radii2D);
  prof_end(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e)), SFB_ERR_CUDA;
  return SFB_OK;
}

int sfb_densify_masks(int P, const float* xyz_gradient_accum, const float* denom, const float* scales,
                      const float* opacity, const float* max_radii2D, int raw_params, float grad_threshold,
                      float dense_extent, float big_extent, float min_opacity, float max_screen_size,
                      uint8_t* clone_mask, uint8_t* split_mask, uint8_t* prune_mask, uint32_t* counts, void* stream) {
  using namespace sfb;
  if (P < 0 || (P > 0 && (!xyz_gradient_accum || !denom || !scales || !opacity || !clone_mask || !split_mask || !prune_mask)))
    return set_error("sfb_densify_masks: bad arguments"), SFB_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (counts && cudaMemsetAsync(counts, 0, 3 * sizeof(uint32_t), s) != cudaSuccess)
    return set_error("sfb_densify_masks: memset failed"), SFB_ERR_CUDA;
  if (P == 0) return SFB_OK;
  prof_begin("densify.masks", s);
  densify_masks_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, xyz_gradient_accum, denom, scales, opacity, max_radii2D, raw_params,
                                                       grad_threshold, dense_extent, min_opacity, max_screen_size,
                                                       big_extent, clone_mask, split_mask, prune_mask, counts);
  prof_end(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e)), SFB_ERR_CUDA;
  return SFB_OK;
}

}  // extern "C"
