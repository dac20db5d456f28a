// Reference point only (NOT part of the product): how fast is CUB's DeviceRadixSort on this GPU for the two
// sorts of the binning stage?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a cub_sort_bench.cu -o cub_sort_bench
#include <cub/cub.cuh>
#include <cstdio>
#include <vector>
#include <random>

static void run(size_t n, int end_bit, uint32_t key_mask, const char* name) {
  std::vector<uint32_t> hk(n), hv(n);
  std::mt19937 rng(1);
  for (size_t i = 0; i < n; i++) { hk[i] = rng() & key_mask; hv[i] = (uint32_t)i; }
  uint32_t *k0, *k1, *v0, *v1;
  cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
  cudaMemcpy(k0, hk.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(v0, hv.data(), n * 4, cudaMemcpyHostToDevice);
  void* tmp = nullptr; size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (int)n, 0, end_bit);
  cudaMalloc(&tmp, tmp_bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; w++) cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (int)n, 0, end_bit);
  cudaEventRecord(a);
  const int iters = 20;
  for (int w = 0; w < iters; w++) cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (int)n, 0, end_bit);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("{\"case\": \"%s\", \"n\": %zu, \"bits\": %d, \"us_per_sort\": %.2f}\n", name, n, end_bit, ms * 1000.f / iters);
  cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
}

int main() {
  run(1000000, 32, 0xFFFFFFFFu, "depth_sort_1M_32bit");
  run(8954935, 12, 0xFFFu, "tile_sort_9M_12bit");
  run(36933939, 13, 0x1FFFu, "tile_sort_37M_13bit");
  return 0;
}
