// Micro-benchmark: shared / global atomic throughput on random addresses (sizes the voxel-partition design).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE, int BINS, int ITERS>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t* gcnt, int gbins) {
  __shared__ uint32_t h[BINS];
  for (int i = threadIdx.x; i < BINS; i += 256) h[i] = 0;
  __syncthreads();
  uint32_t acc = 0, s = mix(blockIdx.x * 256 + threadIdx.x + 1);
#pragma unroll 8
  for (int it = 0; it < ITERS; ++it) {
    s = mix(s + it);
    if (MODE == 0) acc += atomicAdd(&h[s % BINS], 1u);            // returning shared add
    else if (MODE == 1) atomicAdd(&h[s % BINS], 1u);              // non-returning shared add
    else if (MODE == 2) atomicOr(&h[(s >> 5) % BINS], 1u << (s & 31));  // bitmap set
    else if (MODE == 3) acc += atomicAdd(&gcnt[s % gbins], 1u);   // returning global add
    else if (MODE == 4) atomicAdd(&gcnt[s % gbins], 1u);          // non-returning global add
    else if (MODE == 5) acc += h[s % BINS];                       // plain LDS for scale
  }
  __syncthreads();
  if (MODE != 3 && MODE != 4) for (int i = threadIdx.x; i < BINS; i += 256) acc += h[i];
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}
template <int MODE, int BINS>
void run(const char* name, uint32_t* out, uint32_t* g, int gbins) {
  constexpr int ITERS = 256;
  const int blocks = 148 * 24;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, BINS, ITERS><<<blocks, 256>>>(out, g, gbins);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k<MODE, BINS, ITERS><<<blocks, 256>>>(out, g, gbins);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = 5.0 * blocks * 256.0 * ITERS;
  printf("%-44s bins=%5d gbins=%8d  %8.1f G ops/s  (%s)\n", name, BINS, gbins, ops / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  uint32_t *out, *g; cudaMalloc(&out, 148 * 24 * 256 * 4); cudaMalloc(&g, 64 << 20); cudaMemset(g, 0, 64 << 20);
  run<5, 4096>("shared load (scale)", out, g, 1);
  run<0, 4096>("shared atomicAdd returning", out, g, 1);
  run<0, 256>("shared atomicAdd returning", out, g, 1);
  run<1, 4096>("shared atomicAdd no return", out, g, 1);
  run<1, 128>("shared atomicAdd no return", out, g, 1);
  run<2, 4096>("shared atomicOr bitmap", out, g, 1);
  run<3, 32>("global atomicAdd returning", out, g, 3133);
  run<3, 32>("global atomicAdd returning", out, g, 3133 * 256);
  run<4, 32>("global atomicAdd no return", out, g, 3133 * 256);
  run<4, 32>("global atomicAdd no return", out, g, 3133);
  return 0;
}
