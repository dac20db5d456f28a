// activate.cu — SURVEY.md §8f-3: the step right BEFORE the rasterizer in a training iteration, get_gaussian_dict's
// static branch (train.py:42-50) = the property getters of GaussianModel (scene/gaussian_model.py:64-86):
//     get_scaling  = exp(_scaling)                       (.repeat(1, 3) when use_isotropic)        :53, :64-68
//     get_rotation = normalize(_rotation)  (x / max(||x||_2, 1e-12), torch.nn.functional.normalize) :61, :70-72
//     get_opacity  = sigmoid(_opacity)                                                              :58, :83-85
//     get_features = cat(_features_dc [P,1,3], _features_rest [P,M-1,3], dim=1)                     :78-82
// and, for the dynamic branch, the `ret['scales'] + scaling` epilogue (train.py:73).  The reference runs 4 elementwise
// kernels + a cat (and autograd's 4 backward kernels + a split) — each a full pass over its tensor; here one kernel
// per direction.  Pure streaming work: (8 + 3M) floats in and out per Gaussian, HBM-bound.
//
// Work split: thread t handles the 8 small per-Gaussian values of Gaussian t, and (independently) four consecutive
// floats of the flat [P][M][3] feature tensor — so the wide side of the cat / split always moves as one 16-byte
// access per thread and the narrow rows (12 B and 12(M-1) B, not 16-byte aligned) as coalesced 4-byte accesses.
#include "../../include/splat_b200.h"
#include "common.cuh"

namespace sfb {

struct ActParams {
  int P, M, iso;
  const float *raw_scaling, *raw_rotation, *raw_opacity, *f_dc, *f_rest, *scale_offset;
  float *scales, *rotations, *opacity, *features;
};

__device__ __forceinline__ float act_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) activate_forward_kernel(ActParams p) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (size_t)p.P) {
    float s0, s1, s2;
    if (p.iso) { s0 = s1 = s2 = expf(p.raw_scaling[t]); }
    else { s0 = expf(p.raw_scaling[3 * t]); s1 = expf(p.raw_scaling[3 * t + 1]); s2 = expf(p.raw_scaling[3 * t + 2]); }
    if (p.scale_offset) { s0 += p.scale_offset[3 * t]; s1 += p.scale_offset[3 * t + 1]; s2 += p.scale_offset[3 * t + 2]; }
    p.scales[3 * t] = s0; p.scales[3 * t + 1] = s1; p.scales[3 * t + 2] = s2;
    const float4 q = reinterpret_cast<const float4*>(p.raw_rotation)[t];
    const float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float d = fmaxf(len, 1e-12f);
    reinterpret_cast<float4*>(p.rotations)[t] = make_float4(q.x / d, q.y / d, q.z / d, q.w / d);
    p.opacity[t] = act_sigmoid(p.raw_opacity[t]);
  }
  if (p.features) {
    const size_t row = 3 * (size_t)p.M, total = (size_t)p.P * row, e0 = 4 * t;
    if (e0 < total) {
      float v[4];
      // one division per thread (32-bit whenever the element index fits), then the (Gaussian, column) pair of the
      // next three elements by carry
      size_t i = total <= 0xFFFFFFFFull ? (size_t)((uint32_t)e0 / (uint32_t)row) : e0 / row;
      uint32_t c = (uint32_t)(e0 - i * row);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        v[k] = 0.f;
        if (e0 + k < total) v[k] = c < 3 ? __ldg(p.f_dc + 3 * i + c) : __ldg(p.f_rest + i * (row - 3) + (c - 3));
        if (++c == (uint32_t)row) { c = 0; i++; }
      }
      if (e0 + 4 <= total) reinterpret_cast<float4*>(p.features)[t] = make_float4(v[0], v[1], v[2], v[3]);
      else {
#pragma unroll
        for (int k = 0; k < 4; k++) if (e0 + k < total) p.features[e0 + k] = v[k];
      }
    }
  }
}

struct ActBwdParams {
  int P, M, iso;
  const float *raw_scaling, *raw_rotation, *raw_opacity;
  const float *dL_dscales, *dL_drotations, *dL_dopacity, *dL_dfeatures;
  float *d_raw_scaling, *d_raw_rotation, *d_raw_opacity, *d_f_dc, *d_f_rest;
};

__global__ void __launch_bounds__(256) activate_backward_kernel(ActBwdParams p) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (size_t)p.P) {
    if (p.d_raw_scaling) {     // d exp(x) = exp(x)
      if (p.iso) {
        const float g = p.dL_dscales[3 * t] + p.dL_dscales[3 * t + 1] + p.dL_dscales[3 * t + 2];
        p.d_raw_scaling[t] = g * expf(p.raw_scaling[t]);
      } else {
#pragma unroll
        for (int a = 0; a < 3; a++) p.d_raw_scaling[3 * t + a] = p.dL_dscales[3 * t + a] * expf(p.raw_scaling[3 * t + a]);
      }
    }
    if (p.d_raw_rotation) {    // y = x / max(|x|, eps):  dx = (g - y (y . g)) / |x|   (|x| > eps),  g / eps otherwise
      const float4 q = reinterpret_cast<const float4*>(p.raw_rotation)[t];
      const float4 g = reinterpret_cast<const float4*>(p.dL_drotations)[t];
      const float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      float4 r;
      if (len > 1e-12f) {
        const float il = 1.f / len;
        const float yx = q.x * il, yy = q.y * il, yz = q.z * il, yw = q.w * il;
        const float dot = yx * g.x + yy * g.y + yz * g.z + yw * g.w;
        r = make_float4((g.x - yx * dot) * il, (g.y - yy * dot) * il, (g.z - yz * dot) * il, (g.w - yw * dot) * il);
      } else {
        r = make_float4(g.x / 1e-12f, g.y / 1e-12f, g.z / 1e-12f, g.w / 1e-12f);
      }
      reinterpret_cast<float4*>(p.d_raw_rotation)[t] = r;
    }
    if (p.d_raw_opacity) {     // d sigmoid = o (1 - o)
      const float o = act_sigmoid(p.raw_opacity[t]);
      p.d_raw_opacity[t] = (p.dL_dopacity[t] * (1.f - o)) * o;     // torch's sigmoid_backward order: (g * (1 - y)) * y
    }
  }
  if (p.dL_dfeatures) {
    const size_t row = 3 * (size_t)p.M, total = (size_t)p.P * row, e0 = 4 * t;
    if (e0 < total) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (e0 + 4 <= total) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p.dL_dfeatures) + t);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) if (e0 + k < total) v[k] = p.dL_dfeatures[e0 + k];
      }
      size_t i = total <= 0xFFFFFFFFull ? (size_t)((uint32_t)e0 / (uint32_t)row) : e0 / row;
      uint32_t c = (uint32_t)(e0 - i * row);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (e0 + k < total) {
          if (c < 3) { if (p.d_f_dc) p.d_f_dc[3 * i + c] = v[k]; }
          else if (p.d_f_rest) p.d_f_rest[i * (row - 3) + (c - 3)] = v[k];
        }
        if (++c == (uint32_t)row) { c = 0; i++; }
      }
    }
  }
}

static inline unsigned act_blocks(int P, int M, bool feats) {
  const size_t n = feats ? ((size_t)P * 3 * (size_t)M + 3) / 4 : 0;
  const size_t threads = n > (size_t)P ? n : (size_t)P;
  return (unsigned)((threads + 255) / 256);
}

}  // namespace sfb

extern "C" {

int sfb_activate_forward(int P, int M, int isotropic, const float* raw_scaling, const float* raw_rotation,
                         const float* raw_opacity, const float* f_dc, const float* f_rest, const float* scale_offset,
                         float* scales, float* rotations, float* opacity, float* features, void* stream) {
  using namespace sfb;
  if (P < 0 || M < 0) return set_error("sfb_activate_forward: bad sizes"), SFB_ERR_ARG;
  if (P == 0) return SFB_OK;
  if (!raw_scaling || !raw_rotation || !raw_opacity || !scales || !rotations || !opacity)
    return set_error("sfb_activate_forward: null pointer"), SFB_ERR_ARG;
  if (features && (M < 1 || !f_dc || (M > 1 && !f_rest)))
    return set_error("sfb_activate_forward: features need f_dc (and f_rest when M > 1)"), SFB_ERR_ARG;
  if (((reinterpret_cast<size_t>(raw_rotation) | reinterpret_cast<size_t>(rotations) |
        reinterpret_cast<size_t>(features)) & 15) != 0)
    return set_error("sfb_activate_forward: rotations / features must be 16-byte aligned"), SFB_ERR_ARG;
  ActParams p{P, M, isotropic, raw_scaling, raw_rotation, raw_opacity, f_dc, f_rest, scale_offset,
              scales, rotations, opacity, features};
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("activate.forward", s);
  activate_forward_kernel<<<act_blocks(P, M, features != nullptr), 256, 0, s>>>(p);
  prof_end(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e)), SFB_ERR_CUDA;
  return SFB_OK;
}

int sfb_activate_backward(int P, int M, int isotropic, const float* raw_scaling, const float* raw_rotation,
                          const float* raw_opacity, const float* dL_dscales, const float* dL_drotations,
                          const float* dL_dopacity, const float* dL_dfeatures, float* dL_draw_scaling,
                          float* dL_draw_rotation, float* dL_draw_opacity, float* dL_df_dc, float* dL_df_rest,
                          void* stream) {
  using namespace sfb;
  if (P < 0 || M < 0) return set_error("sfb_activate_backward: bad sizes"), SFB_ERR_ARG;
  if (P == 0) return SFB_OK;
  if ((dL_draw_scaling && (!raw_scaling || !dL_dscales)) || (dL_draw_rotation && (!raw_rotation || !dL_drotations)) ||
      (dL_draw_opacity && (!raw_opacity || !dL_dopacity)) || ((dL_df_dc || dL_df_rest) && (!dL_dfeatures || M < 1)))
    return set_error("sfb_activate_backward: an output is requested without its inputs"), SFB_ERR_ARG;
  if (((reinterpret_cast<size_t>(raw_rotation) | reinterpret_cast<size_t>(dL_drotations) |
        reinterpret_cast<size_t>(dL_draw_rotation) | reinterpret_cast<size_t>(dL_dfeatures)) & 15) != 0)
    return set_error("sfb_activate_backward: rotations / features must be 16-byte aligned"), SFB_ERR_ARG;
  const bool feats = dL_dfeatures && (dL_df_dc || dL_df_rest);
  ActBwdParams p{P, M, isotropic, raw_scaling, raw_rotation, raw_opacity, dL_dscales, dL_drotations, dL_dopacity,
                 feats ? dL_dfeatures : nullptr, dL_draw_scaling, dL_draw_rotation, dL_draw_opacity, dL_df_dc, dL_df_rest};
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("activate.backward", s);
  activate_backward_kernel<<<act_blocks(P, M, feats), 256, 0, s>>>(p);
  prof_end(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e)), SFB_ERR_CUDA;
  return SFB_OK;
}

}  // extern "C"
