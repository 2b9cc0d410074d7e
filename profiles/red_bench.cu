// Micro-benchmark: throughput of no-return global reductions (the grid-gradient scatter of the fused step kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/red_bench.bin profiles/red_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red_v2(float* p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory"); }
__device__ __forceinline__ void red_s(float* p, float a) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory"); }
// mode 0: v2 random entries; 1: scalar random; 2: v2, 8 neighbouring lanes hit the same entry; 3: v2 lane-consecutive entries (coalesced);
// 4: v4 random (two entries per op)
__global__ void __launch_bounds__(256) red_kernel(float* table, uint32_t mask, int iters, int mode) {
  uint32_t s = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t e = (s >> 8) & mask;
    if (mode == 2) e = __shfl_sync(0xffffffffu, e, threadIdx.x & 24);
    if (mode == 3) e = (__shfl_sync(0xffffffffu, e, 0) + (threadIdx.x & 31)) & mask;
    if (mode == 1) red_s(table + 2 * (size_t)e, 1.f);
    else if (mode == 4) red_v4(table + 4 * (size_t)(e >> 1), 1.f, 2.f, 3.f, 4.f);
    else red_v2(table + 2 * (size_t)e, 1.f, 2.f);
  }
}
int main() {
  const uint32_t entries = 1u << 23;                       // 8.4 M entries x 8 B = 67 MB (C2's gradient table: 8.7 M entries)
  float* t; cudaMalloc(&t, (size_t)entries * 8); cudaMemset(t, 0, (size_t)entries * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"v2 random", "scalar random", "v2 random, 8 lanes share an entry", "v2 lane-consecutive", "v4 random"};
  for (int mode = 0; mode < 5; ++mode)
    for (int blocks : {148, 296, 592}) {
      const int iters = 512;
      red_kernel<<<blocks, 256>>>(t, entries - 1, iters, mode);
      cudaEventRecord(e0);
      red_kernel<<<blocks, 256>>>(t, entries - 1, iters, mode);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)blocks * 256 * iters;
      printf("%-36s blocks %4d : %7.1f us, %6.2f G lane-ops/s, %5.2f cyc/lane/SM @1.9GHz\n", names[mode], blocks, ms * 1e3, ops / ms / 1e6,
             ms * 1e-3 * 1.9e9 * 148 / ops);
    }
  return 0;
}
