// Fused dense Adam (exact torch.optim.Adam math, nerf_runner.py:502: betas (0.9,0.999), eps 1e-15, no weight decay)
// + GradScaler semantics (nerf_runner.py:159, 756-761) + optimizer.zero_grad() + the fp16 shadow-table refresh
// (grid.py:50-51 casts the whole fp32 table to fp16 every step) in ONE streaming pass:
//   read g, m, v, p (16 B/param)  ->  write p, m, v, g=0 (16 B/param) + 2 B/param fp16 shadow.
// HBM-bound by construction: 128-bit loads/stores, grid sized to a multiple of the SM count.
#include <algorithm>

#include "nof_common.cuh"

namespace nof {

constexpr int ADAM_MAX_SEGS = 8;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC_PER_THREAD = 4;                      // 4 x float4 per thread per tile
constexpr int ADAM_TILE = ADAM_THREADS * ADAM_VEC_PER_THREAD * 4;   // elements per block tile

struct AdamArgs {
  NofAdamSeg seg[ADAM_MAX_SEGS];
  uint32_t tile_begin[ADAM_MAX_SEGS + 1];
  int n_segs;
  float beta1, beta2, eps;
};

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(AdamArgs a, const int32_t* __restrict__ step_ptr,
                                                            const float* __restrict__ scale_state,
                                                            const int32_t* __restrict__ found_inf, uint32_t total_tiles) {
  __shared__ float s_bc[3];
  if (threadIdx.x == 0) {
    const int step = (step_ptr ? *step_ptr : 0) + 1;        // this update's 1-based step
    const double bc1 = 1.0 - pow((double)a.beta1, (double)step);
    const double bc2 = 1.0 - pow((double)a.beta2, (double)step);
    s_bc[0] = (float)(1.0 / bc1);
    s_bc[1] = (float)sqrt(bc2);
    s_bc[2] = scale_state ? 1.0f / scale_state[0] : 1.0f;
  }
  __syncthreads();
  const bool skip = found_inf && (*found_inf != 0);
  const float inv_bc1 = s_bc[0], sqrt_bc2 = s_bc[1], inv_scale = s_bc[2];
  const float b1 = a.beta1, b2 = a.beta2, eps = a.eps;

  for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int si = 0;
#pragma unroll
    for (int k = 1; k < ADAM_MAX_SEGS; ++k) si += (k < a.n_segs && tile >= a.tile_begin[k]) ? 1 : 0;
    const NofAdamSeg sg = a.seg[si];
    const size_t base = (size_t)(tile - a.tile_begin[si]) * ADAM_TILE;
    const float step_size = (sg.lr_ptr ? __ldg(sg.lr_ptr) : sg.lr) * inv_bc1;
#pragma unroll
    for (int j = 0; j < ADAM_VEC_PER_THREAD; ++j) {
      const size_t i = base + ((size_t)j * ADAM_THREADS + threadIdx.x) * 4;
      if (i >= sg.n) continue;
      if (i + 4 <= sg.n) {
        float4 g = *reinterpret_cast<const float4*>(sg.grad + i);
        *reinterpret_cast<float4*>(sg.grad + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (skip) continue;
        float4 p = *reinterpret_cast<const float4*>(sg.param + i);
        float4 m = *reinterpret_cast<const float4*>(sg.exp_avg + i);
        float4 v = *reinterpret_cast<const float4*>(sg.exp_avg_sq + i);
        float* gp = &g.x; float* pp = &p.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float gg = gp[c] * inv_scale;
          mp[c] = mp[c] * b1 + gg * (1.f - b1);
          vp[c] = vp[c] * b2 + (gg * gg) * (1.f - b2);
          const float denom = sqrtf(vp[c]) / sqrt_bc2 + eps;
          pp[c] = pp[c] - step_size * (mp[c] / denom);
        }
        *reinterpret_cast<float4*>(sg.param + i) = p;
        *reinterpret_cast<float4*>(sg.exp_avg + i) = m;
        *reinterpret_cast<float4*>(sg.exp_avg_sq + i) = v;
        if (sg.shadow_f16) {
          __half2 h0 = __floats2half2_rn(p.x, p.y), h1 = __floats2half2_rn(p.z, p.w);
          uint2 pk;
          memcpy(&pk.x, &h0, 4);
          memcpy(&pk.y, &h1, 4);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(sg.shadow_f16) + i) = pk;
        }
      } else {
        for (size_t e = i; e < sg.n; ++e) {
          const float gg = sg.grad[e] * inv_scale;
          sg.grad[e] = 0.f;
          if (skip) continue;
          const float m = sg.exp_avg[e] * b1 + gg * (1.f - b1);
          const float v = sg.exp_avg_sq[e] * b2 + (gg * gg) * (1.f - b2);
          const float denom = sqrtf(v) / sqrt_bc2 + eps;
          const float p = sg.param[e] - step_size * (m / denom);
          sg.param[e] = p; sg.exp_avg[e] = m; sg.exp_avg_sq[e] = v;
          if (sg.shadow_f16) reinterpret_cast<__half*>(sg.shadow_f16)[e] = __float2half_rn(p);
        }
      }
    }
  }
}

// GradScaler.update() (growth_factor 2, backoff 0.5, growth_interval 2000) + step counter + flag reset.
__global__ void adam_finish_kernel(int32_t* step_ptr, float* scale_state, int32_t* found_inf, unsigned long long* tick) {
  if (tick) *tick += 1ull;
  const bool inf = found_inf && (*found_inf != 0);
  if (!inf && step_ptr) *step_ptr += 1;
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

}  // namespace nof

using namespace nof;

extern "C" int nof_adam_step(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, int32_t* step,
                             float* scale_state, int32_t* found_inf, uint64_t* tick, nof_stream_t stream) {
  NOF_REQUIRE(segs && n_segs >= 1 && n_segs <= ADAM_MAX_SEGS, "nof_adam_step: n_segs=%d (1..%d)", n_segs, ADAM_MAX_SEGS);
  AdamArgs a;
  a.n_segs = n_segs;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  uint64_t tiles = 0;
  for (int i = 0; i < ADAM_MAX_SEGS; ++i) {
    a.tile_begin[i] = (uint32_t)tiles;
    if (i < n_segs) {
      a.seg[i] = segs[i];
      NOF_REQUIRE(segs[i].param && segs[i].grad && segs[i].exp_avg && segs[i].exp_avg_sq, "nof_adam_step: null pointer in segment %d", i);
      NOF_REQUIRE(((uintptr_t)segs[i].param | (uintptr_t)segs[i].grad | (uintptr_t)segs[i].exp_avg | (uintptr_t)segs[i].exp_avg_sq) % 16 == 0,
                  "nof_adam_step: segment %d not 16-byte aligned", i);
      tiles += div_up<uint64_t>(segs[i].n, ADAM_TILE);
    } else {
      a.seg[i] = NofAdamSeg{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0.f, nullptr};
    }
  }
  a.tile_begin[ADAM_MAX_SEGS] = (uint32_t)tiles;
  NOF_REQUIRE(tiles < 0xffffffffull, "nof_adam_step: too many elements");
  cudaStream_t st = as_stream(stream);
  if (tiles > 0) {
    int sms = 148;
    nof_device_info(&sms, nullptr);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)sms * 8);
    adam_kernel<<<blocks, ADAM_THREADS, 0, st>>>(a, step, scale_state, found_inf, (uint32_t)tiles);
    int rc = check_launch("adam_kernel");
    if (rc) return rc;
  }
  adam_finish_kernel<<<1, 1, 0, st>>>(step, scale_state, found_inf, reinterpret_cast<unsigned long long*>(tick));
  return check_launch("adam_finish_kernel");
}
