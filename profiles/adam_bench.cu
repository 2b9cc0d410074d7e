// Micro-benchmark behind the design of nof_adam.cu: streaming variants of the fused Adam pass on table-sized arrays.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/adam_bench profiles/adam_bench.cu && /tmp/adam_bench
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float4 adam4(float4& p, float4& m, float4& v, float4 g, float lr, float b1, float b2, float eps, float sbc2) {
  float* pp = &p.x; float* mp = &m.x; float* vp = &v.x; float* gp = &g.x;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    mp[c] = mp[c] * b1 + gp[c] * (1.f - b1);
    vp[c] = vp[c] * b2 + gp[c] * gp[c] * (1.f - b2);
    pp[c] -= lr * (mp[c] / (sqrtf(vp[c]) / sbc2 + eps));
  }
  return p;
}

// V0: the shape of the first implementation: grid-stride tiles, 4 sequential float4 groups per thread
__global__ void __launch_bounds__(256) v0(float* p, float* g, float* m, float* v, __half* sh, size_t n) {
  const size_t tiles = (n + 4095) / 4096;
  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      size_t i = t * 4096 + ((size_t)j * 256 + threadIdx.x) * 4;
      if (i + 4 > n) continue;
      float4 gg = *(float4*)(g + i);
      *(float4*)(g + i) = make_float4(0, 0, 0, 0);
      float4 pp = *(float4*)(p + i), mm = *(float4*)(m + i), vv = *(float4*)(v + i);
      adam4(pp, mm, vv, gg, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f);
      *(float4*)(p + i) = pp; *(float4*)(m + i) = mm; *(float4*)(v + i) = vv;
      __half2 h0 = __floats2half2_rn(pp.x, pp.y), h1 = __floats2half2_rn(pp.z, pp.w);
      uint2 pk; pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1;
      *(uint2*)(sh + i) = pk;
    }
  }
}

// V1: restrict, all loads of the thread's UNR groups issued before any store
template <int UNR, bool STREAM>
__global__ void __launch_bounds__(256) v1(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                          __half* __restrict__ sh, size_t n) {
  const size_t per_block = (size_t)256 * 4 * UNR;
  const size_t tiles = (n + per_block - 1) / per_block;
  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    float4 gg[UNR], pp[UNR], mm[UNR], vv[UNR];
    size_t idx[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      idx[j] = t * per_block + ((size_t)j * 256 + threadIdx.x) * 4;
      if (idx[j] + 4 <= n) {
        if (STREAM) {
          gg[j] = __ldcs((const float4*)(g + idx[j])); pp[j] = __ldcs((const float4*)(p + idx[j]));
          mm[j] = __ldcs((const float4*)(m + idx[j])); vv[j] = __ldcs((const float4*)(v + idx[j]));
        } else {
          gg[j] = *(const float4*)(g + idx[j]); pp[j] = *(const float4*)(p + idx[j]);
          mm[j] = *(const float4*)(m + idx[j]); vv[j] = *(const float4*)(v + idx[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      if (idx[j] + 4 <= n) {
        adam4(pp[j], mm[j], vv[j], gg[j], 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f);
        __half2 h0 = __floats2half2_rn(pp[j].x, pp[j].y), h1 = __floats2half2_rn(pp[j].z, pp[j].w);
        uint2 pk; pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1;
        if (STREAM) {
          __stcs((float4*)(g + idx[j]), make_float4(0, 0, 0, 0)); __stcs((float4*)(p + idx[j]), pp[j]);
          __stcs((float4*)(m + idx[j]), mm[j]); __stcs((float4*)(v + idx[j]), vv[j]);
          *(uint2*)(sh + idx[j]) = pk;                      // the fp16 shadow is what the next step gathers: keep it cacheable
        } else {
          *(float4*)(g + idx[j]) = make_float4(0, 0, 0, 0); *(float4*)(p + idx[j]) = pp[j];
          *(float4*)(m + idx[j]) = mm[j]; *(float4*)(v + idx[j]) = vv[j];
          *(uint2*)(sh + idx[j]) = pk;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) copy4(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (size_t n : {(size_t)9110880, (size_t)59200000}) {
    float *p, *g, *m, *v; __half* sh; float* big;
    CK(cudaMalloc(&p, n * 4)); CK(cudaMalloc(&g, n * 4)); CK(cudaMalloc(&m, n * 4)); CK(cudaMalloc(&v, n * 4)); CK(cudaMalloc(&sh, n * 2));
    CK(cudaMalloc(&big, (size_t)512 << 20));
    CK(cudaMemset(p, 0, n * 4)); CK(cudaMemset(g, 0, n * 4)); CK(cudaMemset(m, 0, n * 4)); CK(cudaMemset(v, 0, n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch, double bytes) {
      for (int w = 0; w < 3; ++w) launch();
      float best = 1e9, tot = 0;
      for (int r = 0; r < 10; ++r) {
        cudaMemsetAsync(big, 1, (size_t)512 << 20);     // flush L2 between repetitions
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; tot += ms;
      }
      printf("n=%zu %-28s best %.1f us  mean %.1f us  -> %.0f GB/s (best)\n", n, name, best * 1e3, tot * 100, bytes / best / 1e6);
    };
    const double bytes = (double)n * 34;
    for (int bps : {4, 8}) {
      const int blocks = sms * bps;
      char nm[64];
      snprintf(nm, 64, "v0 blocks=%dxSM", bps); run(nm, [&] { v0<<<blocks, 256>>>(p, g, m, v, sh, n); }, bytes);
      snprintf(nm, 64, "v1<2,plain> blocks=%dxSM", bps); run(nm, [&] { v1<2, false><<<blocks, 256>>>(p, g, m, v, sh, n); }, bytes);
      snprintf(nm, 64, "v1<4,plain> blocks=%dxSM", bps); run(nm, [&] { v1<4, false><<<blocks, 256>>>(p, g, m, v, sh, n); }, bytes);
      snprintf(nm, 64, "v1<2,stream> blocks=%dxSM", bps); run(nm, [&] { v1<2, true><<<blocks, 256>>>(p, g, m, v, sh, n); }, bytes);
      snprintf(nm, 64, "v1<4,stream> blocks=%dxSM", bps); run(nm, [&] { v1<4, true><<<blocks, 256>>>(p, g, m, v, sh, n); }, bytes);
    }
    {
      const size_t tiles = (n + 1023) / 1024;
      run("v1<1,stream> one tile/block", [&] { v1<1, true><<<(unsigned)tiles, 256>>>(p, g, m, v, sh, n); }, bytes);
      run("v1<2,stream> one tile/block", [&] { v1<2, true><<<(unsigned)((n + 2047) / 2048), 256>>>(p, g, m, v, sh, n); }, bytes);
    }
    run("copy 4 arrays (8n B each)", [&] { copy4<<<sms * 8, 256>>>((const float4*)p, (float4*)g, n / 4); copy4<<<sms * 8, 256>>>((const float4*)m, (float4*)v, n / 4); }, (double)n * 16);
    cudaFree(p); cudaFree(g); cudaFree(m); cudaFree(v); cudaFree(sh); cudaFree(big);
  }
  return 0;
}
