// Micro-benchmark: round-trip latency of one small tcgen05 GEMM (issue -> commit -> mbarrier wait [-> tcgen05.ld]) as used by
// nof_step_tc.cu. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_latency profiles/tc_latency.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return ok != 0;
}

// mode 0: only thread 0 waits (others idle at the final barrier); mode 1: all threads wait on the mbarrier + __syncthreads
__global__ void __launch_bounds__(256) lat_kernel(int N, int K, int bmn, int swz, int reps, int mode, int do_ld, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;   // fp16 ones
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32 * 1024;
  const uint32_t idesc = (1u << 4) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  uint32_t phase = 0;
  long long t_issue = 0, t_total = 0;
  float sink = 0.f;
  for (int r = 0; r < reps; ++r) {
    const long long t0 = clock64();
    if (tid == 0) {
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad, bd;
        if (swz == 0) {
          ad = umma_desc(a_addr + ks * 256, 128, K * 16, 0);
          bd = bmn ? umma_desc(b_addr + ks * 2 * (N * 16), N * 16, 128, 0) : umma_desc(b_addr + ks * 256, 128, K * 16, 0);
        } else {   // 128B swizzle, K-major, 64-element (128 B) rows: SBO = 1024, k-step = 32 B inside the swizzle atom
          ad = umma_desc(a_addr + ks * 32, 16, 1024, 2);
          bd = umma_desc(b_addr + ks * 32, 16, 1024, 2);
        }
        umma_f16(tmem, ad, bd, idesc, ks > 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    const long long t1 = clock64();
    if (mode == 1 || tid == 0) {
      while (!mbar_try(&bar, phase)) {}
    }
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (do_ld) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tmem + ((uint32_t)((warp & 3) * 32) << 16)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      sink += __uint_as_float(v[0]);
    }
    if (mode == 1) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long t2 = clock64();
    if (r >= 4) { t_issue += t1 - t0; t_total += t2 - t0; }
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) { out[0] = t_issue / (reps - 4); out[1] = t_total / (reps - 4); out[2] = (long long)sink; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  struct Cfg { int N, K, bmn, swz, mode, ld, blocks; const char* name; };
  const Cfg cfgs[] = {
      {64, 64, 0, 0, 0, 0, 1, "N64 K64 Kmaj  noswz t0-wait"},      {64, 64, 0, 0, 1, 1, 1, "N64 K64 Kmaj  noswz all-wait+ld+sync"},
      {64, 16, 0, 0, 0, 0, 1, "N64 K16 Kmaj  noswz t0-wait"},      {16, 64, 0, 0, 0, 0, 1, "N16 K64 Kmaj  noswz t0-wait"},
      {64, 64, 1, 0, 0, 0, 1, "N64 K64 MNmaj noswz t0-wait"},      {32, 64, 1, 0, 0, 0, 1, "N32 K64 MNmaj noswz t0-wait"},
      {64, 64, 0, 1, 0, 0, 1, "N64 K64 Kmaj  sw128 t0-wait"},      {64, 64, 0, 1, 1, 1, 1, "N64 K64 Kmaj  sw128 all-wait+ld+sync"},
      {64, 64, 0, 0, 1, 1, 296, "N64 K64 Kmaj  noswz all-wait+ld+sync, 2 CTA/SM"}, {64, 64, 0, 1, 1, 1, 296, "N64 K64 Kmaj  sw128 all-wait+ld+sync, 2 CTA/SM"},
  };
  for (const Cfg& c : cfgs) {
    lat_kernel<<<c.blocks, 256, 48 * 1024>>>(c.N, c.K, c.bmn, c.swz, 200, c.mode, c.ld, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3] = {0, 0, 0};
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("%-60s issue %5lld cyc   round-trip %6lld cyc   (%s)\n", c.name, h[0], h[1], cudaGetErrorString(e));
  }
  return 0;
}
