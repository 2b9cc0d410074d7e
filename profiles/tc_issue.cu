// Micro-benchmark: what ONE tcgen05.mma costs the issuing thread when several are issued back to back (slope of issue time vs count), for
// the shapes a TMEM-resident wgrad would use (M = 64 or 128, A and B MN-major, no-swizzle core-matrix operands), and what the commit costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_issue profiles/tc_issue.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
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

template <int NMMA>
__global__ void __launch_bounds__(128) issue_kernel(int M, int N, int amn, int bmn, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;   // fp16 ones
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 32 * 1024;
  const uint32_t idesc = (1u << 4) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  // operands as [128 rows x 64 cols] core-matrix buffers (row block = 64*16 B); MN-major view: k-step = 2 row blocks
  const uint64_t ad0 = amn ? umma_desc(a_addr, 64 * 16, 128) : umma_desc(a_addr, 128, 64 * 16);
  const uint64_t bd0 = bmn ? umma_desc(b_addr, 64 * 16, 128) : umma_desc(b_addr, 128, 64 * 16);
  const uint64_t ainc = amn ? (uint64_t)((2 * 64 * 16) >> 4) : (uint64_t)(256 >> 4);
  const uint64_t binc = bmn ? (uint64_t)((2 * 64 * 16) >> 4) : (uint64_t)(256 >> 4);
  uint32_t phase = 0;
  long long t_issue = 0, t_commit = 0, t_total = 0;
  for (int r = 0; r < reps; ++r) {
    long long t0 = 0, t1 = 0, t2 = 0;
    if (tid == 0) {
      t0 = clock64();
#pragma unroll
      for (int i = 0; i < NMMA; ++i) umma_f16(tmem, ad0 + (uint64_t)(i & 3) * ainc, bd0 + (uint64_t)(i & 3) * binc, idesc, i > 0);
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      t2 = clock64();
      while (!mbar_try(&bar, phase)) {}
      const long long t3 = clock64();
      if (r >= 4) { t_issue += t1 - t0; t_commit += t2 - t1; t_total += t3 - t0; }
    }
    phase ^= 1u;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0 && blockIdx.x == 0) { out[0] = t_issue / (reps - 4); out[1] = t_commit / (reps - 4); out[2] = t_total / (reps - 4); }
  // read one accumulator element back (ones x ones: K_total)
  if (warp == 0) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tmem));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (tid == 0 && blockIdx.x == 0) out[3] = (long long)__uint_as_float(v[0]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int NMMA>
void run(int M, int N, int amn, int bmn, int blocks, long long* d) {
  cudaFuncSetAttribute(issue_kernel<NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  issue_kernel<NMMA><<<blocks, 128, 64 * 1024>>>(M, N, amn, bmn, 68, d);
  long long h[4];
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("M=%3d N=%3d A_%s B_%s blocks=%3d  n_mma=%2d: issue %5lld cyc (%5.1f / mma), commit %4lld, issue->done %5lld   D[0][0]=%lld %s\n", M, N, amn ? "MN" : "K ",
         bmn ? "MN" : "K ", blocks, NMMA, h[0], (double)h[0] / NMMA, h[1], h[2], h[3], e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int blocks : {1, 296}) {
    run<1>(128, 64, 0, 0, blocks, d);  run<4>(128, 64, 0, 0, blocks, d);  run<8>(128, 64, 0, 0, blocks, d);  run<16>(128, 64, 0, 0, blocks, d);  run<32>(128, 64, 0, 0, blocks, d);
    run<8>(128, 64, 0, 1, blocks, d);  run<8>(128, 16, 0, 1, blocks, d);
    run<1>(64, 64, 1, 1, blocks, d);   run<8>(64, 64, 1, 1, blocks, d);   run<16>(64, 64, 1, 1, blocks, d);  run<32>(64, 64, 1, 1, blocks, d);
    run<8>(64, 16, 1, 1, blocks, d);   run<8>(64, 72, 1, 1, blocks, d);   run<8>(64, 32, 1, 1, blocks, d);
  }
  return 0;
}
