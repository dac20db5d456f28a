// Micro-benchmark (not product): how fast can a one-thread-per-row kernel stream [P][48] fp32 rows in and out,
// compared with a flat coalesced copy?  Decides whether the geometry kernels' 60 % of HBM peak is the access
// pattern (thread-per-row 16-byte requests at a 192-byte stride) or something else (occupancy, dependent phases).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void row_copy(const float4* __restrict__ in, float4* __restrict__ out, int P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float4 v[12];
#pragma unroll
  for (int k = 0; k < 12; k++) v[k] = __ldg(in + (size_t)i * 12 + k);
#pragma unroll
  for (int k = 0; k < 12; k++) { v[k].x += 1.f; out[(size_t)i * 12 + k] = v[k]; }
}

// thread-per-row with 256-bit accesses (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a): one full 32-byte sector per lane
__global__ void row_copy_v8(const float* __restrict__ in, float* __restrict__ out, int P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float v[6][8];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const float* p = in + (size_t)i * 48 + 8 * k;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[k][0]), "=f"(v[k][1]), "=f"(v[k][2]), "=f"(v[k][3]), "=f"(v[k][4]), "=f"(v[k][5]), "=f"(v[k][6]), "=f"(v[k][7])
                 : "l"(p));
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float* q = out + (size_t)i * 48 + 8 * k;
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(q), "f"(v[k][0] + 1.f), "f"(v[k][1]), "f"(v[k][2]),
                 "f"(v[k][3]), "f"(v[k][4]), "f"(v[k][5]), "f"(v[k][6]), "f"(v[k][7])
                 : "memory");
  }
}

// same, but 4 lanes cooperate on one row: lane j of a quad moves float4 j, j+4, j+8 (3 x 64-byte segments / row)
__global__ void quad_copy(const float4* __restrict__ in, float4* __restrict__ out, int P) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int i = t >> 2, j = t & 3;
  if (i >= P) return;
  float4 v[3];
#pragma unroll
  for (int k = 0; k < 3; k++) v[k] = __ldg(in + (size_t)i * 12 + j + 4 * k);
#pragma unroll
  for (int k = 0; k < 3; k++) { v[k].x += 1.f; out[(size_t)i * 12 + j + 4 * k] = v[k]; }
}

__global__ void flat_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { float4 v = __ldg(in + i); v.x += 1.f; out[i] = v; }
}

template <typename F>
static float time_it(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; w++) f();
  cudaEventRecord(a);
  for (int w = 0; w < 20; w++) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 20.f;
}

int main() {
  const int P = 1000000;
  const size_t n4 = (size_t)P * 12;
  float4 *in, *out;
  cudaMalloc(&in, n4 * 16); cudaMalloc(&out, n4 * 16);
  cudaMemset(in, 0, n4 * 16);
  const double gb = 2.0 * n4 * 16 / 1e9;
  float t;
  t = time_it([&] { row_copy<<<(P + 255) / 256, 256>>>(in, out, P); });
  printf("{\"kernel\": \"row_copy_256\", \"ms\": %.4f, \"GBps\": %.0f}\n", t, gb / (t * 1e-3));
  t = time_it([&] { row_copy<<<(P + 127) / 128, 128>>>(in, out, P); });
  printf("{\"kernel\": \"row_copy_128\", \"ms\": %.4f, \"GBps\": %.0f}\n", t, gb / (t * 1e-3));
  t = time_it([&] { row_copy_v8<<<(P + 255) / 256, 256>>>((const float*)in, (float*)out, P); });
  printf("{\"kernel\": \"row_copy_v8_256\", \"ms\": %.4f, \"GBps\": %.0f}\n", t, gb / (t * 1e-3));
  t = time_it([&] { row_copy_v8<<<(P + 127) / 128, 128>>>((const float*)in, (float*)out, P); });
  printf("{\"kernel\": \"row_copy_v8_128\", \"ms\": %.4f, \"GBps\": %.0f}\n", t, gb / (t * 1e-3));
  t = time_it([&] { quad_copy<<<(4 * P + 255) / 256, 256>>>(in, out, P); });
  printf("{\"kernel\": \"quad_copy_256\", \"ms\": %.4f, \"GBps\": %.0f}\n", t, gb / (t * 1e-3));
  t = time_it([&] { flat_copy<<<148 * 8, 256>>>(in, out, n4); });
  printf("{\"kernel\": \"flat_copy\", \"ms\": %.4f, \"GBps\": %.0f}\n", t, gb / (t * 1e-3));
  return 0;
}
