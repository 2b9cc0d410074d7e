// Per-frame pose correction: PoseArray.get_matrices (nerf_helpers.py:143-154; pytorch3d se3_exp_map restated:
// clamp(|w|^2, 1e-4), R = I + sin(a)/a K + (1-cos a)/a^2 K^2, V = I + (1-cos a)/a^2 K + (a - sin a)/a^3 K^2)
// followed by tf = dT @ c2w (nerf_runner.py:1051-1053). One thread per frame; the backward evaluates the same
// expression tree on forward-mode dual numbers (6 tangents = the 6 pose parameters) and contracts with grad_tf, so it
// is the exact derivative of the forward code — replacing ~60 tiny autograd kernels per step with two launches.
#include <algorithm>

#include "nof_common.cuh"

namespace nof {

template <int NT>
struct Dual {
  float v;
  float d[NT];
};
template <int NT> __device__ __forceinline__ Dual<NT> dconst(float c) {
  Dual<NT> r; r.v = c;
#pragma unroll
  for (int i = 0; i < NT; ++i) r.d[i] = 0.f;
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> operator+(const Dual<NT>& a, const Dual<NT>& b) {
  Dual<NT> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> operator-(const Dual<NT>& a, const Dual<NT>& b) {
  Dual<NT> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> operator*(const Dual<NT>& a, const Dual<NT>& b) {
  Dual<NT> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> operator*(const Dual<NT>& a, float c) {
  Dual<NT> r; r.v = a.v * c;
#pragma unroll
  for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> operator/(const Dual<NT>& a, const Dual<NT>& b) {
  Dual<NT> r; const float ib = 1.f / b.v; r.v = a.v * ib;
#pragma unroll
  for (int i = 0; i < NT; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> dchain(const Dual<NT>& a, float f, float df) {
  Dual<NT> r; r.v = f;
#pragma unroll
  for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] * df;
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> dsqrt(const Dual<NT>& a) { float s = sqrtf(a.v); return dchain(a, s, 0.5f / s); }
template <int NT> __device__ __forceinline__ Dual<NT> dsin(const Dual<NT>& a) { return dchain(a, sinf(a.v), cosf(a.v)); }
template <int NT> __device__ __forceinline__ Dual<NT> dcos(const Dual<NT>& a) { return dchain(a, cosf(a.v), -sinf(a.v)); }
template <int NT> __device__ __forceinline__ Dual<NT> dtanh(const Dual<NT>& a) { float t = tanhf(a.v); return dchain(a, t, 1.f - t * t); }
template <int NT> __device__ __forceinline__ Dual<NT> dclamp_min(const Dual<NT>& a, float lo) {
  return a.v < lo ? dconst<NT>(lo) : a;        // torch.clamp backward: zero gradient where clamped
}

// Scalar-type-generic evaluation of the 3x4 correction dT(data) (rows of [R | V t]).
struct FOps {
  using T = float;
  static __device__ __forceinline__ T c(float x) { return x; }
  static __device__ __forceinline__ T sqrt_(T x) { return sqrtf(x); }
  static __device__ __forceinline__ T sin_(T x) { return sinf(x); }
  static __device__ __forceinline__ T cos_(T x) { return cosf(x); }
  static __device__ __forceinline__ T tanh_(T x) { return tanhf(x); }
  static __device__ __forceinline__ T clamp_min(T x, float lo) { return fmaxf(x, lo); }
};
struct DOps {
  using T = Dual<6>;
  static __device__ __forceinline__ T c(float x) { return dconst<6>(x); }
  static __device__ __forceinline__ T sqrt_(const T& x) { return dsqrt(x); }
  static __device__ __forceinline__ T sin_(const T& x) { return dsin(x); }
  static __device__ __forceinline__ T cos_(const T& x) { return dcos(x); }
  static __device__ __forceinline__ T tanh_(const T& x) { return dtanh(x); }
  static __device__ __forceinline__ T clamp_min(const T& x, float lo) { return dclamp_min(x, lo); }
};

template <typename O>
__device__ __forceinline__ void delta_pose(const typename O::T data[6], float max_trans, float max_rot_deg,
                                           typename O::T M[12]) {
  using T = typename O::T;
  T th[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) th[i] = O::tanh_(data[i]);
  T v[3], w[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = th[i] * max_trans;                                          // nerf_helpers.py:148
    w[i] = th[3 + i] * max_rot_deg * (1.0f / 180.0f) * 3.14159265358979323846f;   // :149
  }
  T nr = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  T ang = O::sqrt_(O::clamp_min(nr, 1e-4f));
  T one = O::c(1.f);
  T inv = one / ang;
  T s = O::sin_(ang), co = O::cos_(ang);
  T fac1 = inv * s;
  T fac2 = inv * inv * (one - co);
  T fac3 = (ang - s) / (ang * ang * ang);
  // K = hat(w); K2 = K@K
  T K[9] = {O::c(0.f), O::c(0.f) - w[2], w[1], w[2], O::c(0.f), O::c(0.f) - w[0], O::c(0.f) - w[1], w[0], O::c(0.f)};
  T K2[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j] + K[i * 3 + 2] * K[2 * 3 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T t = O::c(0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T eye = O::c(i == j ? 1.f : 0.f);
      M[i * 4 + j] = fac1 * K[i * 3 + j] + fac2 * K2[i * 3 + j] + eye;
      T Vij = eye + fac2 * K[i * 3 + j] + fac3 * K2[i * 3 + j];
      t = t + Vij * v[j];
    }
    M[i * 4 + 3] = t;
  }
}

__device__ __forceinline__ void pose_forward_frame(const float* __restrict__ pose_data, const float* __restrict__ c2w, float* __restrict__ tf,
                                                   int f, float max_trans, float max_rot_deg) {
  const float* A = c2w + (size_t)f * 16;
  float M[12];
  if (pose_data == nullptr || f == 0) {        // frame 0 is pinned to identity (nerf_helpers.py:151-153)
#pragma unroll
    for (int i = 0; i < 12; ++i) tf[(size_t)f * 12 + i] = A[i];
    return;
  }
  float d[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) d[i] = pose_data[(size_t)f * 6 + i];
  delta_pose<FOps>(d, max_trans, max_rot_deg, M);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = M[i * 4 + 0] * A[0 * 4 + j] + M[i * 4 + 1] * A[1 * 4 + j] + M[i * 4 + 2] * A[2 * 4 + j] + M[i * 4 + 3] * A[3 * 4 + j];
      tf[(size_t)f * 12 + i * 4 + j] = acc;
    }
}

__global__ void pose_forward_kernel(const float* __restrict__ pose_data, const float* __restrict__ c2w, float* __restrict__ tf,
                                    int F, float max_trans, float max_rot_deg) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) pose_forward_frame(pose_data, c2w, tf, f, max_trans, max_rot_deg);
}

// Everything a train step does before the ray march, in ONE launch (each of the pieces is launch-latency-sized): the batch gather
// from the device-resident ray pool at a DEVICE-resident cursor (so that a CUDA graph can hold many consecutive steps), the pose
// correction of all frames, and the bump of the sampler's RNG tick and of the cursor. The counters are advanced by the last CTA to
// finish (completion ticket), i.e. after every CTA has read them.
__global__ void __launch_bounds__(256) step_prologue_kernel(const NofPrologue p) {
  __shared__ long long s_base;
  if (threadIdx.x == 0) s_base = (p.pool && p.cursor) ? *p.cursor : 0;
  __syncthreads();
  const long long base = s_base;
  const int gid = blockIdx.x * 256 + threadIdx.x, gsz = gridDim.x * 256;
  if (p.pool) {
    const int total = p.N * p.ray_dim;
    for (int i = gid; i < total; i += gsz) {
      const int r = i / p.ray_dim, c = i - r * p.ray_dim;
      long long k = base + r;
      if (k >= p.n_ids) k -= p.n_ids;                       // never taken when the host keeps whole batches inside an epoch
      p.batch[i] = __ldg(p.pool + (size_t)p.ids[k] * p.ray_dim + c);
    }
  }
  for (int f = gid; f < p.F; f += gsz) pose_forward_frame(p.pose_data, p.c2w, p.tf, f, p.max_trans, p.max_rot_deg);
  if (p.done) {
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(p.done, 1) == (int)gridDim.x - 1) {
      *p.done = 0;
      if (p.pool && p.cursor) *p.cursor = base + p.N;
      if (p.tick) *p.tick += 1ull;
      if (p.gstep) {
        const long long g = *p.gstep;
        if (p.trunc_table && p.trunc_out && p.trunc_len > 0) *p.trunc_out = p.trunc_table[g < 0 ? 0 : (g >= p.trunc_len ? p.trunc_len - 1 : g)];
        *p.gstep = g + 1;
      }
    }
  }
}

__global__ void pose_backward_kernel(const float* __restrict__ pose_data, const float* __restrict__ c2w,
                                     const float* __restrict__ grad_tf, float* __restrict__ grad_pose, int F, float max_trans,
                                     float max_rot_deg, const float* __restrict__ loss_scale) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F || f == 0) return;
  const float* A = c2w + (size_t)f * 16;
  const float* G = grad_tf + (size_t)f * 12;
  // dL/dM = G @ A[:, :]^T restricted to the 3x4 block:  tf = M(3x4) @ A(4x4)
  float gM[12];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += G[i * 4 + j] * A[k * 4 + j];
      gM[i * 4 + k] = acc;
    }
  Dual<6> d[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    d[i] = dconst<6>(pose_data[(size_t)f * 6 + i]);
    d[i].d[i] = 1.f;
  }
  Dual<6> M[12];
  delta_pose<DOps>(d, max_trans, max_rot_deg, M);
  const float sc = loss_scale ? 1.0f / (*loss_scale) : 1.0f;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) acc += gM[i] * M[i].d[k];
    grad_pose[(size_t)f * 6 + k] += acc * sc;
  }
}

}  // namespace nof

using namespace nof;

extern "C" int nof_pose_forward(const float* pose_data, const float* c2w, float* tf, int F, float max_trans, float max_rot_deg,
                                nof_stream_t stream) {
  NOF_REQUIRE(c2w && tf, "nof_pose_forward: null pointer");
  NOF_REQUIRE(F >= 0, "nof_pose_forward: F=%d", F);
  if (F == 0) return NOF_OK;
  pose_forward_kernel<<<div_up(F, 64), 64, 0, as_stream(stream)>>>(pose_data, c2w, tf, F, max_trans, max_rot_deg);
  return check_launch("pose_forward_kernel");
}

extern "C" int nof_step_prologue(const NofPrologue* p, nof_stream_t stream) {
  NOF_REQUIRE(p && p->c2w && p->tf && p->F >= 1, "nof_step_prologue: null pose pointers or F < 1");
  NOF_REQUIRE(!p->pool || (p->ids && p->batch && p->N >= 1 && p->ray_dim >= 1 && p->n_ids >= p->N), "nof_step_prologue: bad gather arguments");
  NOF_REQUIRE(p->done || !(p->cursor || p->tick || p->gstep), "nof_step_prologue: counters need the `done` ticket");
  NOF_REQUIRE(!p->trunc_out || (p->trunc_table && p->gstep && p->trunc_len >= 1), "nof_step_prologue: trunc_out needs trunc_table, trunc_len and gstep");
  const int work = std::max(p->pool ? p->N * p->ray_dim : 0, p->F);
  step_prologue_kernel<<<std::max(1, std::min(div_up(work, 256), 128)), 256, 0, as_stream(stream)>>>(*p);
  return check_launch("step_prologue_kernel");
}

extern "C" int nof_pose_backward(const float* pose_data, const float* c2w, const float* grad_tf, float* grad_pose, int F,
                                 float max_trans, float max_rot_deg, const float* loss_scale_or_null, nof_stream_t stream) {
  NOF_REQUIRE(pose_data && c2w && grad_tf && grad_pose, "nof_pose_backward: null pointer");
  NOF_REQUIRE(F >= 0, "nof_pose_backward: F=%d", F);
  if (F == 0) return NOF_OK;
  pose_backward_kernel<<<div_up(F, 64), 64, 0, as_stream(stream)>>>(pose_data, c2w, grad_tf, grad_pose, F, max_trans, max_rot_deg,
                                                                     loss_scale_or_null);
  return check_launch("pose_backward_kernel");
}
