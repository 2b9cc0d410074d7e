// Ray-side kernels of the hot path:
//   * postprocess_kernel        — 1:1 with common.cu:129-149 (pack kaolin nuggets into a padded [N,I,2])
//   * interval_walk_kernel      — 1:1 with common.cu:41-105 (cumulative occupied length -> z), never spins
//   * ray_march_kernel          — fused replacement of OctreeManager.ray_trace (Utils.py:443-475, kaolin trace +
//                                 postprocess) + sample_rays_uniform_occupied_voxels (nerf_runner.py:979-1011)
//                                 + around-depth samples (nerf_runner.py:1063-1081): one warp per ray, DDA through a
//                                 dense (2^level)^3 occupancy bitmask held in shared memory, intervals kept in
//                                 shared memory, samples produced by all 32 lanes. No host sync, no [N,I,2] round
//                                 trip through HBM unless the caller asks for the tap.
// All float arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) so it is reproducible
// operation-for-operation by the numpy oracle.
#include "nof_common.cuh"

namespace nof {

// ------------------------------------------------------------------------------------------------
__global__ void postprocess_kernel(const int64_t* __restrict__ ray_index, const float* __restrict__ depth_in_out,
                                   const int64_t* __restrict__ unique_ids, const int64_t* __restrict__ start_poss, int M,
                                   int U, int max_inter, int n_rays, float* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  const int64_t i_ray = unique_ids[u];
  if (i_ray < 0 || i_ray >= n_rays) return;
  float* dst = out + (size_t)i_ray * max_inter * 2;
  int k = 0;
  for (int64_t i = start_poss[u]; i < M; ++i) {
    if (ray_index[i] != i_ray) break;
    const float a = depth_in_out[2 * i], b = depth_in_out[2 * i + 1];
    if (a == 0.f || b == 0.f) break;
    if (a > b) continue;
    if ((double)fabsf(__fsub_rn(b, a)) < 1e-4) continue;
    if (k >= max_inter) break;                 // the reference would write out of bounds here
    dst[2 * k] = a;
    dst[2 * k + 1] = b;
    ++k;
  }
}

// ------------------------------------------------------------------------------------------------
// Walk one sample through an interval list held anywhere addressable. Returns z; sets *err on inconsistency.
template <typename IoPtr>
__device__ __forceinline__ float walk_intervals(IoPtr io, int I, float rem, bool* err) {
  const float eps = 1e-4f;
  int k = 0;
  while (true) {
    if (k >= I) {
      if (!(rem <= eps)) *err = true;
      return io[2 * (I - 1) + 1];
    }
    const float a = io[2 * k];
    if (a == 0.f) {
      if (!(rem <= eps && k >= 1)) *err = true;
      return k >= 1 ? io[2 * (k - 1) + 1] : 0.f;
    }
    const float len = __fsub_rn(io[2 * k + 1], a);
    if (rem <= len) return __fadd_rn(a, rem);
    rem = __fsub_rn(rem, len);
    ++k;
  }
}

// x = sample (fastest, coalesced), y = ray — the reference maps x to rays (uncoalesced, common.cu:116-122).
__global__ void interval_walk_kernel(const float* __restrict__ z_in_out, const float* __restrict__ z_sampled,
                                     float* __restrict__ z_vals, int N, int I, int S, int32_t* err_flag) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (s >= S || r >= N) return;
  const float* io = z_in_out + (size_t)r * I * 2;
  if (io[0] == 0.f) return;                    // common.cu:54
  bool err = false;
  const float z = walk_intervals(io, I, z_sampled[(size_t)r * S + s], &err);
  z_vals[(size_t)r * S + s] = z;
  if (err && err_flag) atomicExch(err_flag, 1);
}

// ------------------------------------------------------------------------------------------------
constexpr int MARCH_WARPS = 8;
constexpr int WALK_FLOATS = 3 * 64 + 8;     // per warp: event times of the three axes (up to 64 crossings each, level <= 6) + 8 words of keep bits

// torch.linspace(0,1,S) as its CUDA kernel evaluates it (step=(end-start)/(steps-1); lower half start+step*i, upper half
// end-step*(steps-i-1), each contracted to one FMA); the division is hoisted out of the per-sample code.
struct Lin { float step; int S; };
__device__ __forceinline__ float lin_at(const Lin& L, int i) {
  if (L.S == 1) return 0.f;
  return (i < L.S / 2) ? __fmul_rn(L.step, (float)i) : __fmaf_rn(-L.step, (float)(L.S - i - 1), 1.0f);
}
// sample_rays_uniform (nerf_runner.py:67-87) for one (ray, sample): stratified value in [near, far].
__device__ __forceinline__ float strat_one(const Lin& L, int i, float nearv, float farv, bool perturb, float u) {
  auto zlin = [&](int j) {
    const float t = lin_at(L, j);
    return __fadd_rn(__fmul_rn(nearv, __fsub_rn(1.f, t)), __fmul_rn(farv, t));
  };
  float z = zlin(i);
  if (perturb) {
    const float lower = (i == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, zlin(i - 1)));
    const float upper = (i == L.S - 1) ? z : __fmul_rn(0.5f, __fadd_rn(zlin(i + 1), z));
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
    z = fminf(fmaxf(z, nearv), farv);
  }
  return z;
}

// One warp per ray. Phase 1 (the walk through the occupancy grid) is warp-parallel over the plane-crossing events (see the comment in the
// kernel; round 1 walked the cells sequentially on lane 0: 3.4 k warp instructions per ray, half of them that dependent chain). The
// reference's `(double)|dt| < 1e-4` is the float test `|dt| <= 1e-4f` (the float nearest to 1e-4 lies below it, the next one above).
// Phase 2 (samples): all 32 lanes, a lane owns 4 consecutive samples (one Philox call, one 16-byte store).
__global__ void __launch_bounds__(MARCH_WARPS * 32) ray_march_kernel(NofMarchCfg cfg, const float* __restrict__ rays,
                                                                     const float* __restrict__ tf,
                                                                     const uint32_t* __restrict__ occ_bits,
                                                                     const float* __restrict__ t_rand,
                                                                     float* __restrict__ z_vals,
                                                                     float* __restrict__ intervals_out,
                                                                     int32_t* err_flag) {
  extern __shared__ uint32_t smem_u32[];
  const int n = 1 << cfg.level;
  const int occ_words = (n * n * n + 31) / 32;
  const int I = cfg.I_max;
  uint32_t* s_occ = smem_u32;
  float* s_io_t = reinterpret_cast<float*>(smem_u32 + occ_words);       // [MARCH_WARPS][I][2] travel-time intervals per ray
  float* s_io_z = s_io_t + (size_t)MARCH_WARPS * I * 2;                 // [MARCH_WARPS][I][2] z intervals (scaled, clipped)
  int* s_count = reinterpret_cast<int*>(s_io_z + (size_t)MARCH_WARPS * I * 2);   // [MARCH_WARPS] intervals, [MARCH_WARPS] overflow
  float* s_walk = reinterpret_cast<float*>(s_count + 2 * MARCH_WARPS);            // [MARCH_WARPS][WALK_FLOATS] event times + keep bits
  for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s_occ[i] = occ_bits[i];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = cfg.S_occ + cfg.S_depth;
  if (cfg.offset_ptr) cfg.offset += *cfg.offset_ptr;
  const Lin lin_occ = {cfg.S_occ > 1 ? __fdiv_rn(1.0f, (float)(cfg.S_occ - 1)) : 0.f, cfg.S_occ};
  const Lin lin_d = {cfg.S_depth > 1 ? __fdiv_rn(1.0f, (float)(cfg.S_depth - 1)) : 0.f, cfg.S_depth};

  __syncthreads();                                     // bitmask staged
  for (int r = blockIdx.x * MARCH_WARPS + warp; r < cfg.N; r += gridDim.x * MARCH_WARPS) {
    // ================= phase 1: the ray's walk through the occupancy grid (replaces kaolin unbatched_raytrace; packing rule of
    // common.cu:137-148), WARP-PARALLEL. The exit time of a cell through the k-th plane of axis a is a closed form of k and non-decreasing
    // in k, so the sequential voxel walk is the merge of three sorted lists ordered by (T, axis) — `t < t_out` with the axes tried in order
    // 0,1,2 is exactly that tie rule (oracle.ray_trace_intervals_merge, bit-identical to the sequential oracle.ray_trace_intervals).
    // Lane = event: its rank m in the merged order comes from two binary searches, step m's cell from the per-axis event counts before
    // it, t_in from the latest of the three predecessors; only the packing rule needs two warp reductions and a bitmask prefix count.
    {
      const float* row = rays + (size_t)r * cfg.ray_dim;
      const float dx = row[0], dy = row[1], dz = row[2];
      const float* T = tf + (size_t)((int)row[8]) * 12;
      // unit camera-frame direction, world origin and world direction (nerf_runner.py:1045-1057)
      const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      const float ux = __fdiv_rn(dx, nrm), uy = __fdiv_rn(dy, nrm), uz = __fdiv_rn(dz, nrm);
      float o[3], d[3], inv[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        o[a] = T[a * 4 + 3];
        d[a] = __fadd_rn(__fadd_rn(__fmul_rn(T[a * 4 + 0], ux), __fmul_rn(T[a * 4 + 1], uy)), __fmul_rn(T[a * 4 + 2], uz));
        inv[a] = __fdiv_rn(1.0f, d[a]);
      }
      float* io_t = s_io_t + (size_t)warp * I * 2;
      float* sT = s_walk + (size_t)warp * WALK_FLOATS;          // [3][64] event times, then 8 words of keep bits
      uint32_t* sKeep = reinterpret_cast<uint32_t*>(sT + 3 * 64);
      const float INF = __int_as_float(0x7f800000);
      float t0 = 0.f, t1 = INF;
      bool hit = true;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (d[a] == 0.f) {
          if (o[a] < -1.f || o[a] > 1.f) hit = false;
        } else {
          const float ta = __fmul_rn(__fsub_rn(-1.f, o[a]), inv[a]);
          const float tb = __fmul_rn(__fsub_rn(1.f, o[a]), inv[a]);
          t0 = fmaxf(t0, fminf(ta, tb));
          t1 = fminf(t1, fmaxf(ta, tb));
        }
      }
      int count = 0;
      bool overflow = false;
      if (lane < 8) sKeep[lane] = 0u;
      if (hit && t0 < t1) {                                     // warp-uniform (every lane computed the same ray)
        const float cell = __fdiv_rn(2.0f, (float)n);
        int ix0[3], step[3], K[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float p = __fadd_rn(o[a], __fmul_rn(t0, d[a]));
          const int c = (int)floorf(__fdiv_rn(__fadd_rn(p, 1.0f), cell));
          ix0[a] = min(max(c, 0), n - 1);
          step[a] = d[a] > 0.f ? 1 : (d[a] < 0.f ? -1 : 0);
          K[a] = step[a] > 0 ? n - ix0[a] : (step[a] < 0 ? ix0[a] + 1 : 0);     // crossings until the walk leaves the grid on this axis
        }
        // event times T_a[k] (padded with +inf): the walk's plane_t of cell ix0_a + k step_a
        for (int e = lane; e < 3 * 64; e += 32) {
          const int a = e >> 6, k = e & 63;
          float t = INF;
          if (k < K[a]) {
            const int ixa = ix0[a] + k * step[a];
            const float plane = __fsub_rn(__fmul_rn((float)(ixa + (step[a] > 0 ? 1 : 0)), cell), 1.0f);
            t = __fmul_rn(__fsub_rn(plane, o[a]), inv[a]);
          }
          sT[e] = t;
        }
        __syncwarp();
        // number of events of list `b` (K_b long) ordered before time t: strictly smaller, or also equal when `incl`
        auto count_before = [&](int bb, float t, bool incl) {
          const float* L = sT + bb * 64;
          int lo = 0, hi = K[bb];
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const float v = L[mid];
            if (v < t || (incl && v == t)) lo = mid + 1; else hi = mid;
          }
          return lo;
        };
        const int n_ev = 3 * n;                                 // event slots: axis a = e / n, crossing k = e % n
        int m_end = 3 * n + 2, first_zero = 0x7fffffff;
        // pass 1: every lane evaluates its events (at most ceil(3n/32) of them); keeps them in registers for pass 2
        constexpr int MAX_EV = 6;                               // 3 * 64 / 32
        int ev_m[MAX_EV];
        float ev_in[MAX_EV], ev_out[MAX_EV];
        bool ev_occ[MAX_EV];
#pragma unroll
        for (int j = 0; j < MAX_EV; ++j) {
          ev_m[j] = -1; ev_in[j] = 0.f; ev_out[j] = 0.f; ev_occ[j] = false;
          const int e = lane + 32 * j;
          if (e < n_ev) {
            const int a = e / n, k = e - a * n;
            if (k < K[a]) {
              const int b1 = a == 0 ? 1 : 0, b2 = a == 2 ? 1 : 2;            // the other two axes, b1 < b2
              const float t = sT[a * 64 + k];
              const int c1 = count_before(b1, t, b1 < a), c2 = count_before(b2, t, b2 < a);
              const int m = k + c1 + c2;
              float tprev = -INF;
              if (k > 0) tprev = sT[a * 64 + k - 1];
              if (c1 > 0) tprev = fmaxf(tprev, sT[b1 * 64 + c1 - 1]);
              if (c2 > 0) tprev = fmaxf(tprev, sT[b2 * 64 + c2 - 1]);
              const float t_out = fminf(t, t1);
              const float t_in = m == 0 ? t0 : fminf(tprev, t1);
              int cix[3];
              cix[a] = ix0[a] + k * step[a];
              cix[b1] = ix0[b1] + c1 * step[b1];
              cix[b2] = ix0[b2] + c2 * step[b2];
              const int cid = (min(max(cix[0], 0), n - 1) * n + min(max(cix[1], 0), n - 1)) * n + min(max(cix[2], 0), n - 1);
              ev_m[j] = m; ev_in[j] = t_in; ev_out[j] = t_out;
              ev_occ[j] = (s_occ[cid >> 5] >> (cid & 31)) & 1u;
              if (k == K[a] - 1 || t_out >= t1) m_end = min(m_end, m);        // the walk ends with the step that leaves the grid / reaches t1
            }
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m_end = min(m_end, __shfl_xor_sync(0xffffffffu, m_end, off));
#pragma unroll
        for (int j = 0; j < MAX_EV; ++j) {
          ev_occ[j] = ev_occ[j] && ev_m[j] >= 0 && ev_m[j] <= m_end;
          if (ev_occ[j] && (ev_in[j] == 0.f || ev_out[j] == 0.f)) first_zero = min(first_zero, ev_m[j]);   // the packing rule's `break`
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) first_zero = min(first_zero, __shfl_xor_sync(0xffffffffu, first_zero, off));
        // pass 2: keep flags as a bitmask over the step index, then ordered compaction by prefix popcount
#pragma unroll
        for (int j = 0; j < MAX_EV; ++j) {
          ev_occ[j] = ev_occ[j] && ev_m[j] < first_zero && !(ev_in[j] > ev_out[j]) && !(fabsf(__fsub_rn(ev_out[j], ev_in[j])) <= 1e-4f);
          if (ev_occ[j]) atomicOr(&sKeep[ev_m[j] >> 5], 1u << (ev_m[j] & 31));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < MAX_EV; ++j) {
          if (ev_occ[j]) {
            const int m = ev_m[j];
            int pos = __popc(sKeep[m >> 5] & ((1u << (m & 31)) - 1u));
            for (int wq = 0; wq < (m >> 5); ++wq) pos += __popc(sKeep[wq]);
            if (pos < I) { io_t[2 * pos] = ev_in[j]; io_t[2 * pos + 1] = ev_out[j]; }
          }
        }
        int total = 0;
        for (int wq = 0; wq < 8; ++wq) total += __popc(sKeep[wq]);
        overflow = total > I;
        count = min(total, I);
      }
      if (lane == 0) {
        s_count[warp] = count;
        s_count[MARCH_WARPS + warp] = overflow ? 1 : 0;
      }
    }
    __syncwarp();
    // ================= phase 2: samples
    const float* row = rays + (size_t)r * cfg.ray_dim;
    const float dx = row[0], dy = row[1], dz = row[2];
    const float depth = row[6];
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float uz = __fdiv_rn(dz, nrm);
    float* io_t = s_io_t + (size_t)warp * I * 2;
    float* io_z = s_io_z + (size_t)warp * I * 2;
    const int count = s_count[warp];
    const bool overflow = s_count[MARCH_WARPS + warp] != 0;
    // ---- z intervals: t -> z (nerf_runner.py:987-990), clip to depth+trunc for valid-depth rays (:992-999)
    const float absz = fabsf(uz);
    const bool valid_depth = (depth >= cfg.near_sc) && (depth <= cfg.far_sc);
    const float trunc = cfg.trunc_ptr ? __ldg(cfg.trunc_ptr) : cfg.trunc;     // annealed truncation: a device scalar (NofPrologue.trunc_out)
    const float zmax = __fadd_rn(depth, trunc);
    for (int k = lane; k < I; k += 32) {
      float a = 0.f, b = 0.f;
      if (k < count) {
        a = __fmul_rn(io_t[2 * k], absz);
        b = __fmul_rn(io_t[2 * k + 1], absz);
      }
      if (intervals_out) {
        intervals_out[((size_t)r * I + k) * 2] = k < count ? io_t[2 * k] : 0.f;
        intervals_out[((size_t)r * I + k) * 2 + 1] = k < count ? io_t[2 * k + 1] : 0.f;
      }
      io_t[2 * k] = a;                          // keep the unclipped z intervals for the invalid-depth branch
      io_t[2 * k + 1] = b;
      if (valid_depth && a > 0.f && b > 0.f) {
        a = fminf(fmaxf(a, 0.f), zmax);
        b = fminf(fmaxf(b, 0.f), zmax);
      }
      io_z[2 * k] = a;
      io_z[2 * k + 1] = b;
    }
    __syncwarp();
    // total occupied length, summed front to back (every lane redundantly); entries beyond `count` are exact zeros
    float total = 0.f;
    for (int k = 0; k < count; ++k) total = __fadd_rn(total, __fsub_rn(io_z[2 * k + 1], io_z[2 * k]));
    float total2 = 0.f;
    if (cfg.S_depth > 0 && !valid_depth)
      for (int k = 0; k < count; ++k) total2 = __fadd_rn(total2, __fsub_rn(io_t[2 * k + 1], io_t[2 * k]));
    const float nd = __fsub_rn(depth, trunc);
    const float fd = __fadd_rn(depth, __fmul_rn(trunc, cfg.neg_trunc_ratio));
    bool err = false;
    float* zrow = z_vals + (size_t)r * S;
    const bool has_any = io_z[0] != 0.f;       // common.cu:54: rays without intervals keep z = 0
    const bool has_any_t = io_t[0] != 0.f;
    const bool vec_ok = (S & 3) == 0 && (reinterpret_cast<uintptr_t>(z_vals) & 15u) == 0;
    for (int s0 = 4 * lane; s0 < S; s0 += 128) {
      float u4[4] = {0.f, 0.f, 0.f, 0.f};
      if (cfg.perturb) {
        if (t_rand) {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (s0 + j < S) u4[j] = t_rand[(size_t)r * S + s0 + j];
        } else {
          const uint4 rnd = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)(s0 >> 2), (uint32_t)cfg.offset, (uint32_t)(cfg.offset >> 32)),
                                          make_uint2((uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32)));
          u4[0] = u32_to_unit(rnd.x); u4[1] = u32_to_unit(rnd.y); u4[2] = u32_to_unit(rnd.z); u4[3] = u32_to_unit(rnd.w);
        }
      }
      float z4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = s0 + j;
        if (s >= S) break;
        if (s < cfg.S_occ) {                            // stratified over the occupied length (nerf_runner.py:979-1011)
          const float zc = strat_one(lin_occ, s, 0.f, total, cfg.perturb != 0, u4[j]);
          z4[j] = has_any ? walk_intervals(io_z, I, zc, &err) : 0.f;
        } else {                                        // around-depth samples (nerf_runner.py:1063-1081)
          const int jj = s - cfg.S_occ;
          if (valid_depth) {
            z4[j] = strat_one(lin_d, jj, nd, fd, cfg.perturb != 0, u4[j]);
          } else {                                      // second occupied-voxel sampling, unclipped (:1074-1076)
            const float zc = strat_one(lin_d, jj, 0.f, total2, cfg.perturb != 0, u4[j]);
            z4[j] = has_any_t ? walk_intervals(io_t, I, zc, &err) : 0.f;
          }
        }
      }
      if (vec_ok) {
        *reinterpret_cast<float4*>(zrow + s0) = make_float4(z4[0], z4[1], z4[2], z4[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (s0 + j < S) zrow[s0 + j] = z4[j];
      }
    }
    if ((err || (lane == 0 && overflow)) && err_flag) atomicExch(err_flag, 1);
    __syncwarp();
  }
}

__global__ void gather_rays_kernel(const float* __restrict__ pool, const int64_t* __restrict__ ids, float* __restrict__ batch,
                                   int N, int ray_dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * ray_dim) return;
  const int r = i / ray_dim, c = i - r * ray_dim;
  batch[i] = __ldg(pool + (size_t)ids[r] * ray_dim + c);
}

}  // namespace nof

using namespace nof;

extern "C" int nof_sample_rays_uniform_occupied_voxels(const float* z_in_out, const float* z_sampled, float* z_vals, int N,
                                                       int I, int S, int32_t* err_flag, nof_stream_t stream) {
  NOF_REQUIRE(z_in_out && z_sampled && z_vals, "nof_sample_rays_uniform_occupied_voxels: null pointer");
  NOF_REQUIRE(N >= 0 && I >= 1 && S >= 0, "nof_sample_rays_uniform_occupied_voxels: bad sizes N=%d I=%d S=%d", N, I, S);
  if (N == 0 || S == 0) return NOF_OK;
  NOF_REQUIRE(N <= 65535 * 1024, "nof_sample_rays_uniform_occupied_voxels: N too large");
  // grid.y is limited to 65535: fold rays beyond that into multiple launches
  for (int r0 = 0; r0 < N; r0 += 65535) {
    const int n = min(65535, N - r0);
    dim3 grid(div_up(S, 128), n);
    interval_walk_kernel<<<grid, 128, 0, as_stream(stream)>>>(z_in_out + (size_t)r0 * I * 2, z_sampled + (size_t)r0 * S,
                                                               z_vals + (size_t)r0 * S, n, I, S, err_flag);
  }
  return check_launch("interval_walk_kernel");
}

extern "C" int nof_postprocess_octree_ray_tracing(const int64_t* ray_index, const float* depth_in_out,
                                                  const int64_t* unique_ids, const int64_t* start_poss, int M, int U,
                                                  int max_intersections, int N_rays, float* out, nof_stream_t stream) {
  NOF_REQUIRE(out, "nof_postprocess_octree_ray_tracing: null output");
  NOF_REQUIRE(max_intersections >= 1 && N_rays >= 0 && M >= 0 && U >= 0, "nof_postprocess_octree_ray_tracing: bad sizes");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(out, 0, (size_t)N_rays * max_intersections * 2 * sizeof(float), st);
  if (U == 0 || M == 0) return check_launch("postprocess memset");
  NOF_REQUIRE(ray_index && depth_in_out && unique_ids && start_poss, "nof_postprocess_octree_ray_tracing: null pointer");
  postprocess_kernel<<<div_up(U, 256), 256, 0, st>>>(ray_index, depth_in_out, unique_ids, start_poss, M, U, max_intersections,
                                                     N_rays, out);
  return check_launch("postprocess_kernel");
}

extern "C" int nof_ray_march(const NofMarchCfg* cfg, const float* rays, const float* tf, const uint32_t* occ_bits,
                             const float* t_rand, float* z_vals, float* intervals_out, int32_t* err_flag,
                             nof_stream_t stream) {
  NOF_REQUIRE(cfg && rays && tf && occ_bits && z_vals, "nof_ray_march: null pointer");
  NOF_REQUIRE(cfg->level >= 0 && cfg->level <= 6, "nof_ray_march: level=%d unsupported (0..6)", cfg->level);
  NOF_REQUIRE(cfg->I_max >= 1 && cfg->I_max <= 256, "nof_ray_march: I_max=%d out of range", cfg->I_max);
  NOF_REQUIRE(cfg->ray_dim >= 9, "nof_ray_march: ray_dim=%d too small", cfg->ray_dim);
  NOF_REQUIRE(cfg->S_occ >= 1 && cfg->S_depth >= 0, "nof_ray_march: bad sample counts");
  if (cfg->N == 0) return NOF_OK;
  const int n = 1 << cfg->level;
  const size_t smem = (size_t)((n * n * n + 31) / 32) * 4 + (size_t)MARCH_WARPS * cfg->I_max * 4 * sizeof(float) + 2 * MARCH_WARPS * sizeof(int) +
                      (size_t)MARCH_WARPS * WALK_FLOATS * sizeof(float);
  if (smem > 48 * 1024)      // per-device attribute: set whenever it is needed, not once per process
    cudaFuncSetAttribute(ray_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  NOF_REQUIRE(smem <= 100 * 1024, "nof_ray_march: shared memory %zu too large", smem);
  int sms = 148;
  nof_device_info(&sms, nullptr);
  const int blocks = min(div_up(cfg->N, MARCH_WARPS), sms * 4);
  ray_march_kernel<<<blocks, MARCH_WARPS * 32, smem, as_stream(stream)>>>(*cfg, rays, tf, occ_bits, t_rand, z_vals,
                                                                          intervals_out, err_flag);
  return check_launch("ray_march_kernel");
}

extern "C" int nof_gather_rays(const float* pool, const int64_t* ids, float* batch, int N, int ray_dim, nof_stream_t stream) {
  NOF_REQUIRE(pool && ids && batch, "nof_gather_rays: null pointer");
  NOF_REQUIRE(N >= 0 && ray_dim >= 1, "nof_gather_rays: bad sizes");
  if (N == 0) return NOF_OK;
  gather_rays_kernel<<<div_up(N * ray_dim, 256), 256, 0, as_stream(stream)>>>(pool, ids, batch, N, ray_dim);
  return check_launch("gather_rays_kernel");
}
