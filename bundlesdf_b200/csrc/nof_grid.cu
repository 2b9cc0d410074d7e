// Op-level multiresolution hash-grid encoder: forward gather (+ d out / d x), backward scatter, input backward.
// Semantics follow mycuda/torch_ngp_grid_encoder/gridencoder.cu:107-365 of the reference exactly (same index
// function, same fp32 arithmetic and rounding points, same [L,B,C] / [B,L,D,C] layouts) so that results are
// bit-comparable for fp32 forward; the implementation (vector loads of the C features of a corner, one thread
// per (point, level) with the corner loop fully unrolled, launch on the caller's stream) is ours.
#include "nof_common.cuh"

namespace nof {

template <typename T> struct Acc;
// fp32: `results += w * g` is contracted by nvcc into fma(w, g, results) in the reference build.
template <> struct Acc<float> {
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ float madd(float acc, float w, float g) { return __fmaf_rn(w, g, acc); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float zero() { return 0.f; }
};
// fp16 (c10::Half arithmetic in the reference): float*Half -> float, Half += float rounds the addend to fp16 and
// the sum to fp16; Half-Half and Half*Half round to fp16.
template <> struct Acc<__half> {
  static __device__ __forceinline__ __half ld(const __half* p) { return __ldg(p); }
  static __device__ __forceinline__ __half madd(__half acc, float w, __half g) {
    __half t = __float2half_rn(__fmul_rn(w, __half2float(g)));
    return __float2half_rn(__half2float(acc) + __half2float(t));
  }
  static __device__ __forceinline__ __half sub(__half a, __half b) { return __float2half_rn(__half2float(a) - __half2float(b)); }
  static __device__ __forceinline__ __half mul(__half a, __half b) { return __float2half_rn(__half2float(a) * __half2float(b)); }
  static __device__ __forceinline__ __half zero() { return __float2half_rn(0.f); }
};

template <uint32_t D>
__device__ __forceinline__ uint32_t corner_index(uint32_t gridtype, bool align_corners, uint32_t hashmap_size,
                                                 uint32_t resolution, const uint32_t pos[D]) {
  constexpr uint32_t primes[3] = {1u, 2654435761u, 805459861u};
  uint32_t stride = 1, index = 0;
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    if (stride <= hashmap_size) {
      index += pos[d] * stride;
      stride *= align_corners ? resolution : (resolution + 1);
    }
  }
  if (gridtype == 0 && stride > hashmap_size) {
    index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) index ^= pos[d] * primes[d];
  }
  return index % hashmap_size;
}

template <typename T, uint32_t C>
__device__ __forceinline__ void load_feat(const T* __restrict__ p, T out[C]) {
  if constexpr (sizeof(T) * C == 4) {
    uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
    memcpy(out, &v, 4);
  } else if constexpr (sizeof(T) * C == 8) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    memcpy(out, &v, 8);
  } else if constexpr (sizeof(T) * C == 16) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    memcpy(out, &v, 16);
  } else if constexpr (sizeof(T) * C == 32) {
    uint4 v0 = __ldg(reinterpret_cast<const uint4*>(p));
    uint4 v1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    memcpy(out, &v0, 16);
    memcpy(reinterpret_cast<char*>(out) + 16, &v1, 16);
  } else {
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) out[c] = Acc<T>::ld(p + c);
  }
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_fwd_kernel(const float* __restrict__ inputs, const T* __restrict__ grid,
                                                       const int32_t* __restrict__ offsets, T* __restrict__ outputs,
                                                       uint32_t B, uint32_t L, float S, uint32_t H, bool calc_grad,
                                                       T* __restrict__ dy_dx, uint32_t gridtype, bool align_corners) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t level = blockIdx.y;
  T* out = outputs + ((size_t)level * B + b) * C;
  T* dyo = calc_grad ? dy_dx + ((size_t)b * L + level) * D * C : nullptr;

  float x[D];
  bool oob = false;
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    x[d] = __ldg(inputs + (size_t)b * D + d);
    oob |= (x[d] < 0.f) || (x[d] > 1.f);
  }
  if (oob) {                                  // gridencoder.cu:128-152
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) out[c] = Acc<T>::zero();
    if (calc_grad) {
#pragma unroll
      for (uint32_t i = 0; i < D * C; ++i) dyo[i] = Acc<T>::zero();
    }
    return;
  }
  const uint32_t off = (uint32_t)offsets[level];
  const uint32_t hashmap_size = (uint32_t)offsets[level + 1] - off;
  const float scale = level_scale(level, S, H);
  const uint32_t resolution = (uint32_t)ceilf(scale) + 1u;
  const T* tab = grid + (size_t)off * C;

  float frac[D];
  uint32_t pg[D];
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    float p = __fmaf_rn(x[d], scale, align_corners ? 0.0f : 0.5f);
    float fl = floorf(p);
    pg[d] = (uint32_t)fl;
    frac[d] = __fsub_rn(p, (float)pg[d]);
  }

  // gather all 2^D corners first (independent loads in flight), then blend in the reference's order
  T feat[1 << D][C];
#pragma unroll
  for (uint32_t idx = 0; idx < (1u << D); ++idx) {
    uint32_t pl[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) pl[d] = pg[d] + ((idx >> d) & 1u);
    const uint32_t index = corner_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
    load_feat<T, C>(tab + (size_t)index * C, feat[idx]);
  }
  T res[C];
#pragma unroll
  for (uint32_t c = 0; c < C; ++c) res[c] = Acc<T>::zero();
#pragma unroll
  for (uint32_t idx = 0; idx < (1u << D); ++idx) {
    float w = 1.f;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) w = __fmul_rn(w, ((idx >> d) & 1u) ? frac[d] : __fsub_rn(1.f, frac[d]));
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) res[c] = Acc<T>::madd(res[c], w, feat[idx][c]);
  }
#pragma unroll
  for (uint32_t c = 0; c < C; ++c) out[c] = res[c];

  if (calc_grad) {                            // gridencoder.cu:202-245
#pragma unroll
    for (uint32_t gd = 0; gd < D; ++gd) {
      T rg[C];
#pragma unroll
      for (uint32_t c = 0; c < C; ++c) rg[c] = Acc<T>::zero();
#pragma unroll
      for (uint32_t idx = 0; idx < (1u << (D - 1)); ++idx) {
        float w = scale;
        uint32_t base = 0;
#pragma unroll
        for (uint32_t nd = 0; nd < D - 1; ++nd) {
          const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
          if ((idx >> nd) & 1u) {
            w = __fmul_rn(w, frac[d]);
            base |= (1u << d);
          } else {
            w = __fmul_rn(w, __fsub_rn(1.f, frac[d]));
          }
        }
#pragma unroll
        for (uint32_t c = 0; c < C; ++c)
          rg[c] = Acc<T>::madd(rg[c], w, Acc<T>::sub(feat[base | (1u << gd)][c], feat[base][c]));
      }
#pragma unroll
      for (uint32_t c = 0; c < C; ++c) dyo[gd * C + c] = rg[c];
    }
  }
}

// One thread per (point, level); all C channels of a corner go out as one vector reduction where the ISA has it.
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_bwd_kernel(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                       const int32_t* __restrict__ offsets, T* __restrict__ grad_grid,
                                                       uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                                       bool align_corners) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t level = blockIdx.y;
  float x[D];
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    x[d] = __ldg(inputs + (size_t)b * D + d);
    if (x[d] < 0.f || x[d] > 1.f) return;     // gridencoder.cu:275-280
  }
  const uint32_t off = (uint32_t)offsets[level];
  const uint32_t hashmap_size = (uint32_t)offsets[level + 1] - off;
  const float scale = level_scale(level, S, H);
  const uint32_t resolution = (uint32_t)ceilf(scale) + 1u;
  T* gg = grad_grid + (size_t)off * C;
  float frac[D];
  uint32_t pg[D];
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    float p = __fmaf_rn(x[d], scale, align_corners ? 0.0f : 0.5f);
    pg[d] = (uint32_t)floorf(p);
    frac[d] = __fsub_rn(p, (float)pg[d]);
  }
  float g[C];
#pragma unroll
  for (uint32_t c = 0; c < C; ++c) {
    if constexpr (sizeof(T) == 2) g[c] = __half2float(__ldg(reinterpret_cast<const __half*>(grad) + ((size_t)level * B + b) * C + c));
    else g[c] = __ldg(reinterpret_cast<const float*>(grad) + ((size_t)level * B + b) * C + c);
  }
#pragma unroll
  for (uint32_t idx = 0; idx < (1u << D); ++idx) {
    float w = 1.f;
    uint32_t pl[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
      const uint32_t bit = (idx >> d) & 1u;
      w = __fmul_rn(w, bit ? frac[d] : __fsub_rn(1.f, frac[d]));
      pl[d] = pg[d] + bit;
    }
    const uint32_t index = corner_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
    T* dst = gg + (size_t)index * C;
    if constexpr (sizeof(T) == 2) {
      if constexpr (C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) {  // gridencoder.cu:321-327 (__half2 atomics)
          __half2 v = __halves2half2(__float2half_rn(w * g[c]), __float2half_rn(w * g[c + 1]));
          atomicAdd(reinterpret_cast<__half2*>(dst + c), v);
        }
      } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) atomicAdd(reinterpret_cast<__half*>(dst + c), __float2half_rn(w * g[c]));
      }
    } else {
      float* d32 = reinterpret_cast<float*>(dst);
      if constexpr (C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) red_add_v2(d32 + c, w * g[c], w * g[c + 1]);
      } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) red_add(d32 + c, w * g[c]);
      }
    }
  }
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_input_bwd_kernel(const T* __restrict__ grad, const T* __restrict__ dy_dx,
                                                             T* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;   // gridencoder.cu:340-365
  if (t >= B * D) return;
  const uint32_t b = t / D, d = t - b * D;
  const T* dy = dy_dx + (size_t)b * L * D * C;
  T result = Acc<T>::zero();
  for (uint32_t l = 0; l < L; ++l) {
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) {
      T gv = grad[((size_t)l * B + b) * C + c];
      T jv = dy[(l * D + d) * C + c];
      if constexpr (sizeof(T) == 2) {
        result = __float2half_rn(__half2float(result) + __half2float(Acc<T>::mul(gv, jv)));
      } else {
        result = __fmaf_rn(gv, jv, result);
      }
    }
  }
  grad_inputs[t] = result;
}

template <typename T, uint32_t D, uint32_t C>
static int launch_fwd(const float* in, const void* emb, const int32_t* off, void* out, uint32_t B, uint32_t L, float S,
                      uint32_t H, bool cg, void* dy, uint32_t gt, bool ac, cudaStream_t st) {
  if (B == 0) return NOF_OK;
  dim3 grid(div_up(B, 256u), L);
  grid_fwd_kernel<T, D, C><<<grid, 256, 0, st>>>(in, (const T*)emb, off, (T*)out, B, L, S, H, cg, (T*)dy, gt, ac);
  return check_launch("grid_fwd_kernel");
}
template <typename T, uint32_t D, uint32_t C>
static int launch_bwd(const void* grad, const float* in, const int32_t* off, void* gg, uint32_t B, uint32_t L, float S,
                      uint32_t H, bool cg, const void* dy, void* gin, uint32_t gt, bool ac, cudaStream_t st) {
  if (B == 0) return NOF_OK;
  dim3 grid(div_up(B, 256u), L);
  grid_bwd_kernel<T, D, C><<<grid, 256, 0, st>>>((const T*)grad, in, off, (T*)gg, B, L, S, H, gt, ac);
  int rc = check_launch("grid_bwd_kernel");
  if (rc) return rc;
  if (cg) {
    grid_input_bwd_kernel<T, D, C><<<div_up(B * D, 256u), 256, 0, st>>>((const T*)grad, (const T*)dy, (T*)gin, B, L);
    rc = check_launch("grid_input_bwd_kernel");
  }
  return rc;
}

__global__ void level_scales_kernel(float S, uint32_t H, int L, float* out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < L) out[l] = level_scale((uint32_t)l, S, H);
}

#define NOF_DISPATCH_DC(FN, ...)                                                              \
  do {                                                                                        \
    if (dtype == NOF_F32) {                                                                   \
      if (D == 3 && C == 2) return FN<float, 3, 2>(__VA_ARGS__);                              \
      if (D == 3 && C == 1) return FN<float, 3, 1>(__VA_ARGS__);                              \
      if (D == 3 && C == 4) return FN<float, 3, 4>(__VA_ARGS__);                              \
      if (D == 3 && C == 8) return FN<float, 3, 8>(__VA_ARGS__);                              \
      if (D == 2 && C == 2) return FN<float, 2, 2>(__VA_ARGS__);                              \
    } else if (dtype == NOF_F16) {                                                            \
      if (D == 3 && C == 2) return FN<__half, 3, 2>(__VA_ARGS__);                             \
      if (D == 3 && C == 1) return FN<__half, 3, 1>(__VA_ARGS__);                             \
      if (D == 3 && C == 4) return FN<__half, 3, 4>(__VA_ARGS__);                             \
      if (D == 3 && C == 8) return FN<__half, 3, 8>(__VA_ARGS__);                             \
      if (D == 2 && C == 2) return FN<__half, 2, 2>(__VA_ARGS__);                             \
    }                                                                                         \
  } while (0)

}  // namespace nof

using namespace nof;

extern "C" int nof_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                                       uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                       int calc_grad_inputs, void* dy_dx, uint32_t gridtype, int align_corners, int dtype,
                                       nof_stream_t stream) {
  NOF_REQUIRE(inputs && embeddings && offsets && outputs, "nof_grid_encode_forward: null pointer");
  NOF_REQUIRE(!calc_grad_inputs || dy_dx, "nof_grid_encode_forward: calc_grad_inputs needs dy_dx");
  NOF_REQUIRE(L >= 1 && L <= 65535, "nof_grid_encode_forward: L=%u out of range", L);
  NOF_REQUIRE(gridtype <= 1, "nof_grid_encode_forward: gridtype must be 0 (hash) or 1 (tiled)");
  cudaStream_t st = as_stream(stream);
  NOF_DISPATCH_DC(launch_fwd, inputs, embeddings, offsets, outputs, B, L, S, H, calc_grad_inputs != 0, dy_dx, gridtype,
                  align_corners != 0, st);
  set_error("nof_grid_encode_forward: D=%u C=%u dtype=%d not built (D in {2,3}, C in {1,2,4,8})", D, C, dtype);
  return NOF_E_UNSUPPORTED;
}

extern "C" int nof_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                        const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C,
                                        uint32_t L, float S, uint32_t H, int calc_grad_inputs, const void* dy_dx,
                                        void* grad_inputs, uint32_t gridtype, int align_corners, int dtype,
                                        nof_stream_t stream) {
  (void)embeddings;   // the reference passes it but the kernels never read it (gridencoder.cu:250-336)
  NOF_REQUIRE(grad && inputs && offsets && grad_embeddings, "nof_grid_encode_backward: null pointer");
  NOF_REQUIRE(!calc_grad_inputs || (dy_dx && grad_inputs), "nof_grid_encode_backward: calc_grad_inputs needs dy_dx and grad_inputs");
  NOF_REQUIRE(L >= 1 && L <= 65535, "nof_grid_encode_backward: L=%u out of range", L);
  NOF_REQUIRE(gridtype <= 1, "nof_grid_encode_backward: gridtype must be 0 (hash) or 1 (tiled)");
  cudaStream_t st = as_stream(stream);
  NOF_DISPATCH_DC(launch_bwd, grad, inputs, offsets, grad_embeddings, B, L, S, H, calc_grad_inputs != 0, dy_dx, grad_inputs,
                  gridtype, align_corners != 0, st);
  set_error("nof_grid_encode_backward: D=%u C=%u dtype=%d not built", D, C, dtype);
  return NOF_E_UNSUPPORTED;
}

extern "C" int nof_grid_level_scales(float S, uint32_t H, int L, float* scales_out, nof_stream_t stream) {
  NOF_REQUIRE(scales_out && L >= 1 && L <= 65535, "nof_grid_level_scales: bad arguments");
  level_scales_kernel<<<div_up(L, 64), 64, 0, as_stream(stream)>>>(S, H, L, scales_out);
  return check_launch("level_scales_kernel");
}
