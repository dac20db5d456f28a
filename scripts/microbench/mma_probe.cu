// Probe: does mma.sync.m16n8k8 tf32 with the hi/lo split reproduce fp32-grade row sums on this GPU?
// One warp, A = [16 x 32] random fp32 staged in shared memory, B = [32 x 8] small exact values.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int MODE>
__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) {
  if (MODE == 0) { hi = __float_as_uint(v) & 0xffffe000u; lo = __float_as_uint(v - __uint_as_float(hi)); }
  else if (MODE == 1) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
  } else { hi = __float_as_uint(v); lo = 0u; }
}

template <int MODE>
__global__ void probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  __shared__ float sA[16 * 32];
  const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
  for (int i = 0; i < 16; i++) sA[i * 32 + (lane ^ ((i & 3) << 2))] = A[i * 32 + lane];
  __syncwarp();
  const int swz = (gid & 3) << 2;
  float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    const int p0 = (ks * 8 + tig) ^ swz, p1 = (ks * 8 + tig + 4) ^ swz;
    float v0 = sA[gid * 32 + p0], v1 = sA[(gid + 8) * 32 + p0], v2 = sA[gid * 32 + p1], v3 = sA[(gid + 8) * 32 + p1];
    uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
    split<MODE>(v0, h0, l0); split<MODE>(v1, h1, l1); split<MODE>(v2, h2, l2); split<MODE>(v3, h3, l3);
    const uint32_t b0 = __float_as_uint(B[(ks * 8 + tig) * 8 + gid]), b1 = __float_as_uint(B[(ks * 8 + tig + 4) * 8 + gid]);
    mma_tf32(d, h0, h1, h2, h3, b0, b1);
    if (MODE != 2) mma_tf32(d, l0, l1, l2, l3, b0, b1);
  }
  D[gid * 8 + 2 * tig] = d[0]; D[gid * 8 + 2 * tig + 1] = d[1];
  D[(gid + 8) * 8 + 2 * tig] = d[2]; D[(gid + 8) * 8 + 2 * tig + 1] = d[3];
}

int main() {
  float hA[16 * 32], hB[32 * 8], hD[16 * 8];
  srand(1);
  for (int i = 0; i < 16 * 32; i++) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int p = 0; p < 32; p++) {
    float X = (p & 7) - 7.5f, Y = (p >> 3) - 7.5f;
    float f[8] = {1.f, X, Y, X * X, X * Y, Y * Y, 0.f, 0.f};
    for (int n = 0; n < 8; n++) hB[p * 8 + n] = f[n];
  }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 5; mode++) {
    // modes 3/4: rows 8..15 hold huge / NaN garbage; rows 0..7 must be unaffected (rows of an MMA are independent)
    if (mode == 3) { for (int i = 8 * 32; i < 16 * 32; i++) hA[i] = 1e30f; cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); }
    if (mode == 4) { for (int i = 8 * 32; i < 16 * 32; i++) hA[i] = nanf(""); cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); }
    if (mode == 3 || mode == 4) probe<0><<<1, 32>>>(dA, dB, dD);
    if (mode == 0) probe<0><<<1, 32>>>(dA, dB, dD);
    if (mode == 1) probe<1><<<1, 32>>>(dA, dB, dD);
    if (mode == 2) probe<2><<<1, 32>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int i = 0; i < (mode >= 3 ? 8 : 16); i++) for (int n = 0; n < 8; n++) {
      double r = 0, a = 0;
      for (int p = 0; p < 32; p++) { r += (double)hA[i * 32 + p] * hB[p * 8 + n]; a += fabs((double)hA[i * 32 + p] * hB[p * 8 + n]); }
      worst = fmax(worst, fabs(hD[i * 8 + n] - r)); scale = fmax(scale, a);
    }
    printf("{\"probe\": \"mma_tf32_split\", \"mode\": %d, \"err\": \"%s\", \"worst_abs\": %.3e, \"sum_abs_terms\": %.3e, \"row0\": [%.6f, %.6f, %.6f]}\n",
           mode, cudaGetErrorString(e), worst, scale, hD[0], hD[1], hD[2]);
  }
  return 0;
}
