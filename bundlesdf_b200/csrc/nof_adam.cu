// Fused dense Adam (exact torch.optim.Adam math, nerf_runner.py:502: betas (0.9,0.999), eps 1e-15, no weight decay)
// + GradScaler semantics (nerf_runner.py:159, 756-761) + optimizer.zero_grad() + the fp16 shadow-table refresh
// (grid.py:50-51 casts the whole fp32 table to fp16 every step) in ONE streaming pass:
//   read g, m, v, p (16 B/param)  ->  write p, m, v, g=0 (16 B/param) + 2 B/param fp16 shadow  = 34 B/param.
// HBM-bound by construction. Shape chosen with profiles/adam_bench.cu on a B200: one 2048-element tile per CTA,
// every thread issues all of its 128-bit streaming loads (ld.global.cs) BEFORE its first store (with stores interleaved
// and no __restrict__ the same pass runs at 1.6 TB/s; this shape sustains 5.8-6.4 TB/s = 88-97 % of the measured copy peak).
#include <algorithm>

#include "nof_common.cuh"

namespace nof {

constexpr int ADAM_MAX_SEGS = 8;
constexpr int ADAM_THREADS = 256;
#ifndef NOF_ADAM_UNR
#define NOF_ADAM_UNR 2
#endif
constexpr int ADAM_UNR = NOF_ADAM_UNR;                      // float4 groups per thread
constexpr int ADAM_TILE = ADAM_THREADS * ADAM_UNR * 4;      // 2048 elements per CTA

struct AdamArgs {
  NofAdamSeg seg[ADAM_MAX_SEGS];
  uint32_t tile_begin[ADAM_MAX_SEGS + 1];
  int n_segs;
  float beta1, beta2, eps;
};

// torch computes the bias corrections in fp64; kept out of line so that the streaming kernel's register count stays at 4 CTAs/SM
#ifdef NOF_ADAM_INLINE_COLD
static __device__ __forceinline__ void bias_corrections_cold(
#else
static __device__ __noinline__ void bias_corrections_cold(
#endif
    float beta1, float beta2, int step, float* out) {
  out[0] = (float)(1.0 / (1.0 - pow((double)beta1, (double)step)));
  out[1] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
}

// GradScaler.update() (growth_factor 2, backoff 0.5, growth_interval 2000) + step counter + RNG tick + flag reset.
__device__ __forceinline__ void adam_finish(int32_t* step_ptr, float* scale_state, int32_t* found_inf, unsigned long long* tick, float beta1,
                                            float beta2) {
  if (tick) *tick += 1ull;
  const bool inf = found_inf && (*found_inf != 0);
  if (!inf && step_ptr) *step_ptr += 1;
  if (inf && step_ptr) {
    step_ptr[5] += 1;                                       // skipped updates so far (GradScaler back-offs)
    if (*found_inf == 2) step_ptr[7] = 1;                   // sticky: the step kernel reported INVALID results (tcgen05 wait timed out)
  }
  if (step_ptr) {                                           // cache the NEXT update's bias corrections (torch computes them in fp64)
    const int next = *step_ptr + 1;
    float bc[2];
    bias_corrections_cold(beta1, beta2, next, bc);
    step_ptr[1] = __float_as_int(bc[0]);
    step_ptr[2] = __float_as_int(bc[1]);
    step_ptr[3] = next;
  }
  if (scale_state) {
    if (inf) {
      scale_state[0] *= 0.5f;
      scale_state[1] = 0.f;
    } else {
      scale_state[1] += 1.f;
      if (scale_state[1] >= 2000.f) {
        scale_state[0] *= 2.0f;
        scale_state[1] = 0.f;
      }
    }
  }
  if (found_inf) *found_inf = 0;
}
// only when the caller passes no step buffer (no place for the completion counter) or there is nothing to update
__global__ void adam_finish_kernel(int32_t* step_ptr, float* scale_state, int32_t* found_inf, unsigned long long* tick, float beta1,
                                   float beta2) {
  adam_finish(step_ptr, scale_state, found_inf, tick, beta1, beta2);
}

#ifndef NOF_ADAM_MINB
#define NOF_ADAM_MINB 4
#endif
__global__ void __launch_bounds__(ADAM_THREADS, NOF_ADAM_MINB) adam_kernel(const AdamArgs a, int32_t* step_ptr, float* scale_state, int32_t* found_inf,
                                                            unsigned long long* tick, int finish_at) {
  const uint32_t tile = blockIdx.x;
  int si = 0;
#pragma unroll
  for (int k = 1; k < ADAM_MAX_SEGS; ++k) si += (k < a.n_segs && tile >= a.tile_begin[k]) ? 1 : 0;
  const NofAdamSeg sg = a.seg[si];
  float* __restrict__ P = sg.param;
  float* __restrict__ G = sg.grad;
  float* __restrict__ M = sg.exp_avg;
  float* __restrict__ V = sg.exp_avg_sq;
  __half* __restrict__ SH = reinterpret_cast<__half*>(sg.shadow_f16);
  const size_t base = (size_t)(tile - a.tile_begin[si]) * ADAM_TILE;

  // ---- issue every load of this thread first
  float4 g[ADAM_UNR], p[ADAM_UNR], m[ADAM_UNR], v[ADAM_UNR];
  size_t idx[ADAM_UNR];
  bool full[ADAM_UNR];
#pragma unroll
  for (int j = 0; j < ADAM_UNR; ++j) {
    idx[j] = base + ((size_t)j * ADAM_THREADS + threadIdx.x) * 4;
    full[j] = idx[j] + 4 <= sg.n;
    if (full[j]) {
      g[j] = __ldcs(reinterpret_cast<const float4*>(G + idx[j]));
      p[j] = __ldcs(reinterpret_cast<const float4*>(P + idx[j]));
      m[j] = __ldcs(reinterpret_cast<const float4*>(M + idx[j]));
      v[j] = __ldcs(reinterpret_cast<const float4*>(V + idx[j]));
    }
  }
  // ---- per-launch scalars: every thread reads them itself (independent loads, L1 hits after the first CTA of the SM) while its
  // data loads are in flight; only a cache miss of the bias corrections (first update, restored counter) takes the FP64 path
  const int step = (step_ptr ? step_ptr[0] : 0) + 1;        // this update's 1-based step
  const bool cached = step_ptr && step_ptr[3] == step;      // bias corrections left by the previous call's adam_finish
  float inv_bc1 = cached ? __int_as_float(step_ptr[1]) : 0.f, sqrt_bc2 = cached ? __int_as_float(step_ptr[2]) : 0.f;
  const float inv_scale = scale_state ? 1.0f / scale_state[0] : 1.0f;
  const bool skip = found_inf && (*found_inf != 0);
  const float lr = sg.lr_ptr ? *sg.lr_ptr : sg.lr;
  if (!cached) {                                            // CTA-uniform
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) bias_corrections_cold(a.beta1, a.beta2, step, s_bc);
    __syncthreads();
    inv_bc1 = s_bc[0];
    sqrt_bc2 = s_bc[1];
  }
  const float step_size = lr * inv_bc1;
  const float b1 = a.beta1, b2 = a.beta2, eps = a.eps;

#pragma unroll
  for (int j = 0; j < ADAM_UNR; ++j) {
    if (full[j]) {
      __stcs(reinterpret_cast<float4*>(G + idx[j]), make_float4(0.f, 0.f, 0.f, 0.f));
      if (skip) continue;
      float* gp = &g[j].x; float* pp = &p[j].x; float* mp = &m[j].x; float* vp = &v[j].x;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float gg = gp[c] * inv_scale;
        mp[c] = mp[c] * b1 + gg * (1.f - b1);
        vp[c] = vp[c] * b2 + (gg * gg) * (1.f - b2);
        const float denom = sqrtf(vp[c]) / sqrt_bc2 + eps;
        pp[c] = pp[c] - step_size * (mp[c] / denom);
      }
      __stcs(reinterpret_cast<float4*>(P + idx[j]), p[j]);
      __stcs(reinterpret_cast<float4*>(M + idx[j]), m[j]);
      __stcs(reinterpret_cast<float4*>(V + idx[j]), v[j]);
      if (SH) {                                            // the fp16 shadow is what the next step gathers: keep it cacheable
        __half2 h0 = __floats2half2_rn(p[j].x, p[j].y), h1 = __floats2half2_rn(p[j].z, p[j].w);
        uint2 pk;
        memcpy(&pk.x, &h0, 4);
        memcpy(&pk.y, &h1, 4);
        *reinterpret_cast<uint2*>(SH + idx[j]) = pk;
      }
    } else if (idx[j] < sg.n) {                            // ragged tail of a segment (< 4 elements)
      for (size_t e = idx[j]; e < sg.n; ++e) {
        const float gg = G[e] * inv_scale;
        G[e] = 0.f;
        if (skip) continue;
        const float mm = M[e] * b1 + gg * (1.f - b1);
        const float vv = V[e] * b2 + (gg * gg) * (1.f - b2);
        const float denom = sqrtf(vv) / sqrt_bc2 + eps;
        const float pn = P[e] - step_size * (mm / denom);
        P[e] = pn; M[e] = mm; V[e] = vv;
        if (SH) SH[e] = __float2half_rn(pn);
      }
    }
  }
  // ---- the last CTA to get here does the scalar bookkeeping (every CTA has read step/scale/found_inf by then): saves a launch.
  // finish_at = the number of CTAs that have to pass: this launch's grid (nof_adam_step) or the sum over the launches that share one step's
  // update (nof_adam_update_shared: e.g. the table segment on one stream and the small segments on another, in either order).
  if (step_ptr && finish_at) {
    __syncthreads();                                        // every thread of this CTA has consumed the scalars (its stores depend on them)
    if (threadIdx.x == 0) {                                 // no fence: the bookkeeping touches nothing the other CTAs write
      if (atomicAdd(step_ptr + 4, 1) == finish_at - 1) {
        step_ptr[4] = 0;
        adam_finish(step_ptr, scale_state, found_inf, tick, a.beta1, a.beta2);
      }
    }
  }
}

}  // namespace nof

using namespace nof;

static int adam_launch(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, int32_t* step, float* scale_state,
                       int32_t* found_inf, uint64_t* tick, bool finish, const char* fn, nof_stream_t stream, int shared_tiles = 0) {
  NOF_REQUIRE(segs && n_segs >= 1 && n_segs <= ADAM_MAX_SEGS, "%s: n_segs=%d (1..%d)", fn, n_segs, ADAM_MAX_SEGS);
  AdamArgs a;
  a.n_segs = n_segs;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  uint64_t tiles = 0;
  for (int i = 0; i < ADAM_MAX_SEGS; ++i) {
    a.tile_begin[i] = (uint32_t)tiles;
    if (i < n_segs) {
      a.seg[i] = segs[i];
      NOF_REQUIRE(segs[i].param && segs[i].grad && segs[i].exp_avg && segs[i].exp_avg_sq, "%s: null pointer in segment %d", fn, i);
      NOF_REQUIRE(((uintptr_t)segs[i].param | (uintptr_t)segs[i].grad | (uintptr_t)segs[i].exp_avg | (uintptr_t)segs[i].exp_avg_sq) % 16 == 0,
                  "%s: segment %d not 16-byte aligned", fn, i);
      NOF_REQUIRE(!segs[i].shadow_f16 || (uintptr_t)segs[i].shadow_f16 % 8 == 0, "%s: shadow of segment %d not 8-byte aligned", fn, i);
      tiles += div_up<uint64_t>(segs[i].n, ADAM_TILE);
    } else {
      a.seg[i] = NofAdamSeg{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0.f, nullptr};
    }
  }
  a.tile_begin[ADAM_MAX_SEGS] = (uint32_t)tiles;
  NOF_REQUIRE(tiles < 0x7fffffffull, "%s: too many elements", fn);
  cudaStream_t st = as_stream(stream);
  unsigned long long* tk = reinterpret_cast<unsigned long long*>(tick);
  if (tiles > 0) {
    adam_kernel<<<(uint32_t)tiles, ADAM_THREADS, 0, st>>>(a, step, scale_state, found_inf, tk, shared_tiles > 0 ? shared_tiles : (finish ? (int)tiles : 0));
    int rc = check_launch("adam_kernel");
    if (rc) return rc;
    if (step || !finish) return NOF_OK;                      // bookkeeping done by the kernel's last CTA / not asked for
  }
  if (!finish) return NOF_OK;
  adam_finish_kernel<<<1, 1, 0, st>>>(step, scale_state, found_inf, tk, beta1, beta2);
  return check_launch("adam_finish_kernel");
}

extern "C" int nof_adam_step(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, int32_t* step,
                             float* scale_state, int32_t* found_inf, uint64_t* tick, nof_stream_t stream) {
  return adam_launch(segs, n_segs, beta1, beta2, eps, step, scale_state, found_inf, tick, true, "nof_adam_step", stream);
}

extern "C" int nof_adam_update(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, const int32_t* step,
                               const float* scale_state, const int32_t* found_inf, nof_stream_t stream) {
  return adam_launch(segs, n_segs, beta1, beta2, eps, const_cast<int32_t*>(step), const_cast<float*>(scale_state),
                     const_cast<int32_t*>(found_inf), nullptr, false, "nof_adam_update", stream);
}

extern "C" int nof_adam_finish(int32_t* step, float* scale_state, int32_t* found_inf, uint64_t* tick, float beta1, float beta2,
                               nof_stream_t stream) {
  adam_finish_kernel<<<1, 1, 0, as_stream(stream)>>>(step, scale_state, found_inf, reinterpret_cast<unsigned long long*>(tick), beta1, beta2);
  return check_launch("adam_finish_kernel");
}

extern "C" int nof_adam_tile_count(const NofAdamSeg* segs, int n_segs) {
  NOF_REQUIRE(segs && n_segs >= 1 && n_segs <= ADAM_MAX_SEGS, "nof_adam_tile_count: n_segs=%d (1..%d)", n_segs, ADAM_MAX_SEGS);
  uint64_t tiles = 0;
  for (int i = 0; i < n_segs; ++i) tiles += div_up<uint64_t>(segs[i].n, ADAM_TILE);
  NOF_REQUIRE(tiles < 0x7fffffffull, "nof_adam_tile_count: too many elements");
  return (int)tiles;
}

extern "C" int nof_adam_update_shared(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, int32_t* step, float* scale_state,
                                      int32_t* found_inf, uint64_t* tick, int total_tiles, nof_stream_t stream) {
  NOF_REQUIRE(step != nullptr, "nof_adam_update_shared: needs the step buffer (its completion counter)");
  NOF_REQUIRE(total_tiles > 0, "nof_adam_update_shared: total_tiles=%d", total_tiles);
  return adam_launch(segs, n_segs, beta1, beta2, eps, step, scale_state, found_inf, tick, false, "nof_adam_update_shared", stream, total_tiles);
}
