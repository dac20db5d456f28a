// xchg.cuh — device helpers shared by the two implementations of the view-parallel gradient exchange over symmetric
// memory (exchange.cu: sfb_xchg_finish; geom_bwd.cu: the fused backward + exchange kernel): system-scope loads / stores,
// multimem (NVSwitch multicast) reductions, the SH row rebuild, flag words of the symmetric buffer.
#pragma once
#include "common.cuh"

namespace sfb {

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;

namespace {

static __constant__ float SHX_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
static __constant__ float SHX_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};

// real SH basis of utils/sh_utils.py:57-112 at direction (x, y, z), the same expressions as geom_backward_kernel
template <int D>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* basis) {
  basis[0] = SH_C0;
  if (D > 0) { basis[1] = -SH_C1 * y; basis[2] = SH_C1 * z; basis[3] = -SH_C1 * x; }
  if (D > 1) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    basis[4] = SHX_C2[0] * xy; basis[5] = SHX_C2[1] * yz; basis[6] = SHX_C2[2] * (2.f * zz - xx - yy);
    basis[7] = SHX_C2[3] * xz; basis[8] = SHX_C2[4] * (xx - yy);
    if (D > 2) {
      basis[9] = SHX_C3[0] * y * (3.f * xx - yy); basis[10] = SHX_C3[1] * xy * z;
      basis[11] = SHX_C3[2] * y * (4.f * zz - xx - yy); basis[12] = SHX_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
      basis[13] = SHX_C3[4] * x * (4.f * zz - xx - yy); basis[14] = SHX_C3[5] * z * (xx - yy);
      basis[15] = SHX_C3[6] * x * (xx - 3.f * yy);
    }
  }
}

// dL_dsh row of Gaussian i: sum over the V views (index order: bit-reproducible across ranks and runs) of
// basis(dir_v) (x) gc_v, one fused multiply-add per term: with 8 views the row costs 8 x 48 of them and the kernel is
// instruction-bound, not HBM-bound (the first view's products are rounded exactly as geom_backward_kernel stores them).
template <int D, bool W256>
__device__ __forceinline__ void sh_row_rebuild(size_t i, size_t view_stride /* floats between the views' [P][3] slots */,
                                               int V, int M, const float* __restrict__ means3D,
                                               const float* s_cam, const float* __restrict__ dcolor,
                                               float* __restrict__ dL_dsh) {
  constexpr int NB = (D + 1) * (D + 1);
  constexpr int NF8 = (3 * NB + 7) / 8;
  const float mx = means3D[3 * i], my = means3D[3 * i + 1], mz = means3D[3 * i + 2];
  float acc[NF8 * 8];
#pragma unroll
  for (int k = 0; k < NF8 * 8; k++) acc[k] = 0.f;
  for (int v = 0; v < V; v++) {
    const float* gp = dcolor + (size_t)v * view_stride + i * 3;
    // (ld.global.cg: in the NVLink exchange this table is written by the PEERS while the kernel is already resident)
    const float g0 = __ldcg(gp), g1 = __ldcg(gp + 1), g2 = __ldcg(gp + 2);
    if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;     // culled in this view (or no gradient reached it)
    const float vx = mx - s_cam[3 * v], vy = my - s_cam[3 * v + 1], vz = mz - s_cam[3 * v + 2];
    const float ilen = rsqrtf(vx * vx + vy * vy + vz * vz);
    float basis[NB];
    sh_basis<D>(vx * ilen, vy * ilen, vz * ilen, basis);
#pragma unroll
    for (int k = 0; k < NB; k++) {
      acc[3 * k] = __fmaf_rn(basis[k], g0, acc[3 * k]);
      acc[3 * k + 1] = __fmaf_rn(basis[k], g1, acc[3 * k + 1]);
      acc[3 * k + 2] = __fmaf_rn(basis[k], g2, acc[3 * k + 2]);
    }
  }
  float* dsh = dL_dsh + i * M * 3;
  if (W256) {      // launcher: M == NB, rows are 32-byte aligned multiples of 32 bytes
#pragma unroll
    for (int k = 0; k < (3 * NB) / 8; k++) stg256(dsh + 8 * k, acc + 8 * k);
  } else {
#pragma unroll
    for (int k = 0; k < 3 * NB; k++) dsh[k] = acc[k];
    for (int k = 3 * NB; k < 3 * M; k++) dsh[k] = 0.f;      // coefficients above the active degree
  }
}

}  // namespace

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until *p >= epoch (flags only grow).  Bounded: a peer that never arrives must not hang the GPU — after ~2 s the
// wait gives up and returns false; the kernel then records the failure in its rank's error word (sfb_xchg_status) and
// retires without touching the outputs.
__device__ __forceinline__ bool spin_until(const uint32_t* p, uint32_t epoch) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 mm_ld_reduce_add(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* mc_addr, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w) : "memory");
}

}  // namespace

// flags inside every rank's symmetric buffer (uint32 words; only ever grow, one epoch pair per step)
//   [FLAG_A + r]  rank r's backward of step `epoch` is complete (its records are in its buffer, its colour gradients in mine)
//   [FLAG_B + r]  rank r has broadcast its slice of the sums of step `epoch`
//   [FLAG_DONE]   local: CTAs of this launch that finished their part of the slice reduction
//   [FLAG_ERR]    local: 0, or (1 | 2: which barrier timed out) | epoch << 8   (sfb_xchg_status)
//   [FLAG_TL .. +12)  local: timeline of the last launch (six 64-bit globaltimer values, see xchg_mark)
constexpr int FLAG_A = 0, FLAG_B = 16, FLAG_DONE = 32, FLAG_TICKET = 33, FLAG_ERR = 34, FLAG_TL = 36;

namespace {
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Timeline of one launch (sfb_xchg_timeline): slot 0 = ~(earliest CTA start), slots 1..5 = latest CTA to pass barrier A,
// finish its part of the slice reduction, finish its SH rows, pass barrier B, finish unpacking.  One 64-bit atomic per
// CTA and phase; cleared by the launcher together with the work-queue ticket.
__device__ __forceinline__ void xchg_mark(uint32_t* flags, int slot, bool invert = false) {
  const unsigned long long t = globaltimer_ns();
  atomicMax(reinterpret_cast<unsigned long long*>(flags + FLAG_TL) + slot, invert ? ~t : t);
}
}  // namespace

}  // namespace sfb
