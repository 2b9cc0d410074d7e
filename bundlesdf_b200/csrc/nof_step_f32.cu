// Fused forward + loss + backward of one train step, fp32 policy (cfg amp: false): fp32 table gathered with 8-byte
// vector loads, fp32 activations in shared memory, the five dense layers on the FP32 pipe (thread = point, weights
// broadcast from shared memory with 128-bit loads), cooperative register-tiled wgrad. Same structure as the AMP
// kernel, except that a ray group is walked in sub-tiles of 128 points: when a group has more than one sub-tile the
// forward is evaluated twice (once for compositing, once — recomputed — right before its backward), so the CTA size
// and shared-memory footprint do not depend on the number of samples per ray.
#include "nof_step_common.cuh"

namespace nof {

constexpr int FT = 128;          // threads per CTA = points per sub-tile
constexpr int LDF64 = 65;        // odd strides: thread t reads row t -> bank (t*ld + k) % 32 is conflict free
constexpr int MAX_GRP_PTS = 1024;

struct F32Plan {
  int w1, w2, w3, w4, w5, bias, x0, x1, xc, x3, x4, d5, d2, pts, rays, lv, total;
  int ld0, ldc, kp3;
};
__host__ __device__ inline F32Plan make_f32_plan(int E, int V, int grp_pts) {
  F32Plan s;
  s.kp3 = (V + 15 + 3) / 4 * 4;
  s.ld0 = E + 1;
  s.ldc = s.kp3 + 1;
  int o = 0;
  auto take = [&](int floats) { int r = o; o += (floats + 31) / 32 * 32; return r; };
  s.w1 = take(64 * E);
  s.w2 = take(16 * 64);
  s.w3 = take(64 * s.kp3);
  s.w4 = take(64 * 64);
  s.w5 = take(4 * 64);
  s.bias = take(216);
  s.x0 = take(FT * s.ld0);
  s.x1 = take(FT * LDF64);
  s.xc = take(FT * s.ldc);
  s.x3 = take(FT * LDF64);
  s.x4 = take(FT * LDF64);
  s.d5 = take(FT * 5);
  s.d2 = take(FT * 17);
  s.pts = take(grp_pts * 6);
  s.rays = take(MAX_R * (int)sizeof(RayS) / 4 + 8);
  s.lv = take((int)sizeof(LevelS) / 4 + 8);
  s.total = o * 4;
  return s;
}

// y[o] = act(b[o] + sum_k x[k] W[o][k]) for this thread's row; W rows 16-byte aligned, K % 4 == 0.
template <int K, bool RELU>
__device__ __forceinline__ void row_fwd(const float* xrow, const float* W, int ldw, const float* b, int N, float* yrow) {
  float x[K];
#pragma unroll
  for (int k = 0; k < K; ++k) x[k] = xrow[k];
  for (int o = 0; o < N; ++o) {
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)o * ldw);
    float acc = b[o];
#pragma unroll
    for (int k = 0; k < K / 4; ++k) {
      const float4 w = w4[k];
      acc = fmaf(x[4 * k + 0], w.x, acc);
      acc = fmaf(x[4 * k + 1], w.y, acc);
      acc = fmaf(x[4 * k + 2], w.z, acc);
      acc = fmaf(x[4 * k + 3], w.w, acc);
    }
    yrow[o] = RELU ? fmaxf(acc, 0.f) : acc;
  }
}
// dx[k] = sum_o dy[o] W[o][k]
template <int K>
__device__ __forceinline__ void row_dgrad(const float* dyrow, int N, const float* W, int ldw, float dx[K]) {
#pragma unroll
  for (int k = 0; k < K; ++k) dx[k] = 0.f;
  for (int o = 0; o < N; ++o) {
    const float dy = dyrow[o];
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)o * ldw);
#pragma unroll
    for (int k = 0; k < K / 4; ++k) {
      const float4 w = w4[k];
      dx[4 * k + 0] = fmaf(dy, w.x, dx[4 * k + 0]);
      dx[4 * k + 1] = fmaf(dy, w.y, dx[4 * k + 1]);
      dx[4 * k + 2] = fmaf(dy, w.z, dx[4 * k + 2]);
      dx[4 * k + 3] = fmaf(dy, w.w, dx[4 * k + 3]);
    }
  }
}
// cooperative wgrad: this thread owns CNT consecutive k of row o: acc[j] += sum_p dY[p][o] * X[p][k0+j]
template <int CNT>
__device__ __forceinline__ void coop_wgrad(const float* dY, int ldy, const float* X, int ldx, int o, int k0, int cnt, float* acc) {
  for (int p = 0; p < FT; ++p) {
    const float dy = dY[(size_t)p * ldy + o];
    const float* xr = X + (size_t)p * ldx + k0;
#pragma unroll
    for (int j = 0; j < CNT; ++j)
      if (j < cnt) acc[j] = fmaf(dy, xr[j], acc[j]);
  }
}
__device__ __forceinline__ float coop_bias(const float* dY, int ldy, int o) {
  float s = 0.f;
  for (int p = 0; p < FT; ++p) s += dY[(size_t)p * ldy + o];
  return s;
}

template <int E_>
__global__ void __launch_bounds__(FT, 1) step_f32_kernel(const StepArgs a) {
  constexpr int E = E_;
  extern __shared__ __align__(128) float smf[];
  const int V = a.V;
  const int grp_pts = a.R * a.Sp;
  const F32Plan sp = make_f32_plan(E, V, grp_pts);
  float* sW1 = smf + sp.w1; float* sW2 = smf + sp.w2; float* sW3 = smf + sp.w3; float* sW4 = smf + sp.w4; float* sW5 = smf + sp.w5;
  float* sB = smf + sp.bias;
  float* X0 = smf + sp.x0; float* X1 = smf + sp.x1; float* XC = smf + sp.xc; float* X3 = smf + sp.x3; float* X4 = smf + sp.x4;
  float* D5 = smf + sp.d5; float* D2 = smf + sp.d2;
  float* sPts = smf + sp.pts;                  // [grp_pts][6]: out4, z, w_raw
  RayS* sRay = reinterpret_cast<RayS*>(smf + sp.rays);
  LevelS& lv = *reinterpret_cast<LevelS*>(smf + sp.lv);
  const int tid = threadIdx.x, lane = tid & 31;
  const int L = a.p.L, K3 = V + 15, KP3 = sp.kp3, LD0 = sp.ld0, LDC = sp.ldc;
  const float scale_ls = a.p.loss_scale ? *a.p.loss_scale : 1.0f;

  // ---- stage parameters (fp32, padded rows)
  {
    const float* P = a.p.mlp;
    for (int i = tid; i < 64 * E; i += FT) sW1[i] = P[a.po[0] + i];
    for (int i = tid; i < 16 * 64; i += FT) sW2[i] = P[a.po[2] + i];
    for (int i = tid; i < 64 * KP3; i += FT) { const int o = i / KP3, k = i % KP3; sW3[i] = k < K3 ? P[a.po[4] + o * K3 + k] : 0.f; }
    for (int i = tid; i < 64 * 64; i += FT) sW4[i] = P[a.po[6] + i];
    for (int i = tid; i < 4 * 64; i += FT) sW5[i] = i < 3 * 64 ? P[a.po[8] + i] : 0.f;
    for (int i = tid; i < 64; i += FT) { sB[i] = P[a.po[1] + i]; sB[80 + i] = P[a.po[5] + i]; sB[144 + i] = P[a.po[7] + i]; }
    for (int i = tid; i < 16; i += FT) sB[64 + i] = P[a.po[3] + i];
    for (int i = tid; i < 8; i += FT) sB[208 + i] = i < 3 ? P[a.po[9] + i] : 0.f;
  }
  init_levels(lv, a);
  __syncthreads();

  // persistent wgrad ownership
  constexpr int CNT1 = E / 2;                  // 64*E/128
  float gW1[CNT1], gW2[8], gW3[16], gW4[32], gW5[2], gB[2];
#pragma unroll
  for (int j = 0; j < CNT1; ++j) gW1[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) gW2[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) gW3[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) gW4[j] = 0.f;
  gW5[0] = gW5[1] = 0.f; gB[0] = gB[1] = 0.f;
  const int cnt3 = KP3 / 2;                    // 64*KP3/128 (KP3 multiple of 4 -> integer)
  float loss_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float n_valid_s = 0.f, n_valid_r = 0.f;

  const int Sp = a.Sp, R = a.R, S = a.p.S;
  const int n_sub = grp_pts / FT;
  float2* Jslot = reinterpret_cast<float2*>(a.jws) + (size_t)blockIdx.x * (MAX_L * 3) * FT;

  for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
    if (tid < R) setup_ray(sRay[tid], a, grp * R + tid);
    __syncthreads();

    // forward of one sub-tile: gather + MLP for point `pid` (this thread); results in X* rows [tid] and out4
    auto forward = [&](int pid, bool want_j, float out4[4], float u[3], float pc[3], bool& valid, bool& active, float& z) {
      const int rl = pid / Sp, sidx = pid - rl * Sp;
      const RayS& rs = sRay[rl];
      active = rs.active && sidx < S;
      z = active ? a.p.z_vals[(size_t)rs.ray * S + sidx] : 0.f;
      float x[3];
      world_point(rs, z, pc, x);
      valid = active && fabsf(x[0]) <= 1.f && fabsf(x[1]) <= 1.f && fabsf(x[2]) <= 1.f;
#pragma unroll
      for (int d = 0; d < 3; ++d) u[d] = (x[d] + 1.0f) * 0.5f;
      float* xc = XC + (size_t)tid * LDC;
      for (int j = 0; j < KP3; ++j) xc[j] = j < V ? rs.views[j] : 0.f;
      float* x0 = X0 + (size_t)tid * LD0;
      if (valid) {
        for (int l = 0; l < L; ++l) {
          float enc[2], J[3][2];
          if (want_j) {
            gather_level<false, true>(a.p.table_f32, lv, l, u, enc, J);
#pragma unroll
            for (int d = 0; d < 3; ++d) Jslot[(size_t)(l * 3 + d) * FT + tid] = make_float2(J[d][0], J[d][1]);
          } else {
            gather_level<false, false>(a.p.table_f32, lv, l, u, enc, J);
          }
          x0[2 * l] = enc[0]; x0[2 * l + 1] = enc[1];
        }
      } else {
        for (int j = 0; j < E; ++j) x0[j] = 0.f;
      }
      float* x1 = X1 + (size_t)tid * LDF64;
      row_fwd<E, true>(x0, sW1, E, sB, 64, x1);
      float h2[16];
      row_fwd<64, false>(x1, sW2, 64, sB + 64, 16, h2);
      out4[3] = h2[0];
#pragma unroll
      for (int j = 0; j < 15; ++j) xc[V + j] = h2[1 + j];
      float* x3 = X3 + (size_t)tid * LDF64;
      // K = KP3 is runtime (24, 28 or 32): dispatch the three padded sizes
      if (KP3 == 24) row_fwd<24, true>(xc, sW3, KP3, sB + 80, 64, x3);
      else if (KP3 == 28) row_fwd<28, true>(xc, sW3, KP3, sB + 80, 64, x3);
      else row_fwd<32, true>(xc, sW3, KP3, sB + 80, 64, x3);
      float* x4 = X4 + (size_t)tid * LDF64;
      row_fwd<64, true>(x3, sW4, 64, sB + 144, 64, x4);
      float o3[4];
      row_fwd<64, false>(x4, sW5, 64, sB + 208, 3, o3);
      out4[0] = o3[0]; out4[1] = o3[1]; out4[2] = o3[2];
    };

    // ============ pass 1: forward of every sub-tile, compositing sums
    float keep_u[3], keep_pc[3], keep_z = 0.f;
    bool keep_valid = false, keep_active = false;
    for (int sub = 0; sub < n_sub; ++sub) {
      const int pid = sub * FT + tid;
      float out4[4];
      forward(pid, a.p.need_pose_grad && n_sub == 1, out4, keep_u, keep_pc, keep_valid, keep_active, keep_z);
      const int rl = pid / Sp;
      const float w_raw = keep_active ? raw_weight(a, keep_z, sRay[rl].depth) : 0.f;
      float* pt = sPts + (size_t)pid * 6;
      pt[0] = out4[0]; pt[1] = out4[1]; pt[2] = out4[2]; pt[3] = out4[3]; pt[4] = keep_z; pt[5] = w_raw;
      const float ws = warp_sum(w_raw);
      const unsigned anyv = __ballot_sync(0xffffffffu, keep_valid);
      if (lane == 0) {
        if (ws != 0.f) atomicAdd(&sRay[rl].sumw, ws);
        if (anyv) atomicOr(&sRay[rl].anyvalid, 1);
      }
    }
    __syncthreads();
    // normalised weights need the complete per-ray sum -> second sweep for rgb_map
    for (int sub = 0; sub < n_sub; ++sub) {
      const int pid = sub * FT + tid;
      const int rl = pid / Sp, sidx = pid - rl * Sp;
      const RayS& rs = sRay[rl];
      const float* pt = sPts + (size_t)pid * 6;
      // validity must be re-derived for sub-tiles other than the last one
      bool valid = false;
      if (rs.active && sidx < S) {
        float pc[3], x[3];
        world_point(rs, pt[4], pc, x);
        valid = fabsf(x[0]) <= 1.f && fabsf(x[1]) <= 1.f && fabsf(x[2]) <= 1.f;
      }
      const float w = valid ? pt[5] / (rs.sumw + 1e-10f) : 0.f;
      float pr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) pr[c] = warp_sum(w * sigmoidf_(pt[c]));
      if (lane == 0 && (pr[0] != 0.f || pr[1] != 0.f || pr[2] != 0.f)) {
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(&sRay[rl].rgb[c], pr[c]);
      }
    }
    __syncthreads();

    // ============ pass 2: per sub-tile (recomputed forward if needed) seeds + backward + scatter
    for (int sub = 0; sub < n_sub; ++sub) {
      const int pid = sub * FT + tid;
      const int rl = pid / Sp, sidx = pid - rl * Sp;
      RayS& rs = sRay[rl];
      float out4[4], u[3], pc[3], z;
      bool valid, active;
      if (n_sub > 1) {
        __syncthreads();                        // previous sub-tile's cooperative wgrad is done with the X buffers
        forward(pid, a.p.need_pose_grad != 0, out4, u, pc, valid, active, z);
      } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) { u[d] = keep_u[d]; pc[d] = keep_pc[d]; }
        valid = keep_valid; active = keep_active; z = keep_z;
        const float* pt = sPts + (size_t)pid * 6;
        out4[0] = pt[0]; out4[1] = pt[1]; out4[2] = pt[2]; out4[3] = pt[3];
      }
      const float w_raw = sPts[(size_t)pid * 6 + 5];
      const float w = valid ? w_raw / (rs.sumw + 1e-10f) : 0.f;
      const float ray_w = rs.ray_w_base * (rs.anyvalid ? 1.f : 0.f);
      float d_out[4];
      loss_seeds(a, rs, out4, z, w, valid, active ? ray_w : 0.f, d_out, loss_acc);
      if (!active) { d_out[0] = d_out[1] = d_out[2] = d_out[3] = 0.f; }
      if (valid) n_valid_s += 1.f;
      if (sidx == 0 && rs.active) {
        float e = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float dd = rs.rgb[c] - rs.gt[c]; e += dd * dd; }
        loss_acc[1] += a.p.rgb_weight * e * ray_w * a.inv_N3;
        if (rs.anyvalid && rs.ray_w_base != 0.f) n_valid_r += 1.f;
        if (a.p.rgb_map) {
#pragma unroll
          for (int c = 0; c < 3; ++c) a.p.rgb_map[(size_t)rs.ray * 3 + c] = rs.rgb[c];
        }
      }
      if (active) {
        const size_t pi = (size_t)rs.ray * S + sidx;
        if (a.p.raw) *reinterpret_cast<float4*>(a.p.raw + pi * 4) = make_float4(out4[0], out4[1], out4[2], out4[3]);
        if (a.p.valid_samples) a.p.valid_samples[pi] = valid ? 1 : 0;
        if (a.p.weights) a.p.weights[pi] = w;
      }
      // ---- own-row backward chain (thread = point); dY written IN PLACE over the activations after the block-wide
      //      barrier that ends each cooperative wgrad.
      float* d5 = D5 + (size_t)tid * 5;
      d5[0] = d_out[0] * scale_ls; d5[1] = d_out[1] * scale_ls; d5[2] = d_out[2] * scale_ls; d5[3] = 0.f;
      const float dsdf = d_out[3] * scale_ls;
      float* x4 = X4 + (size_t)tid * LDF64; float* x3 = X3 + (size_t)tid * LDF64; float* x1 = X1 + (size_t)tid * LDF64;
      float* xc = XC + (size_t)tid * LDC; float* x0 = X0 + (size_t)tid * LD0; float* d2 = D2 + (size_t)tid * 17;
      __syncthreads();
      // layer 5
      if (tid < 96) coop_wgrad<2>(D5, 5, X4, LDF64, tid >> 5, (tid & 31) * 2, 2, gW5);
      if (tid >= 96 && tid < 99) gB[0] += coop_bias(D5, 5, tid - 96);                  // b5
      {
        float dx[64];
        row_dgrad<64>(d5, 3, sW5, 64, dx);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 64; ++k) x4[k] = x4[k] > 0.f ? dx[k] : 0.f;
      }
      __syncthreads();
      // layer 4
      coop_wgrad<32>(X4, LDF64, X3, LDF64, tid >> 1, (tid & 1) * 32, 32, gW4);
      if (tid >= 64) gB[1] += coop_bias(X4, LDF64, tid - 64);                          // b4 (threads 64..127)
      {
        float dx[64];
        row_dgrad<64>(x4, 64, sW4, 64, dx);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 64; ++k) x3[k] = x3[k] > 0.f ? dx[k] : 0.f;
      }
      __syncthreads();
      // layer 3
      coop_wgrad<16>(X3, LDF64, XC, LDC, tid >> 1, (tid & 1) * cnt3, cnt3, gW3);
      float gb3 = 0.f;
      if (tid < 64) gb3 = coop_bias(X3, LDF64, tid);                                   // b3 (threads 0..63)
      {
        float dx[32];
        if (KP3 == 24) row_dgrad<24>(x3, 64, sW3, KP3, dx);
        else if (KP3 == 28) row_dgrad<28>(x3, 64, sW3, KP3, dx);
        else row_dgrad<32>(x3, 64, sW3, KP3, dx);
        // dviews: reduce over the warp's 32 samples (one ray), one shared atomic per column per warp
        for (int j = 0; j < V; ++j) {
          const float v = warp_sum(dx[j]);
          if (lane == 0 && v != 0.f) atomicAdd(&rs.dviews[j], v);
        }
        d2[0] = dsdf;
#pragma unroll
        for (int j = 0; j < 15; ++j) d2[1 + j] = dx[V + j];     // runtime V: dx is indexed dynamically (local memory, small)
      }
      __syncthreads();
      // layer 2
      coop_wgrad<8>(D2, 17, X1, LDF64, tid >> 3, (tid & 7) * 8, 8, gW2);
      float gb2 = 0.f;
      if (tid >= 64 && tid < 80) gb2 = coop_bias(D2, 17, tid - 64);                    // b2
      {
        float dx[64];
        row_dgrad<64>(d2, 16, sW2, 64, dx);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 64; ++k) x1[k] = x1[k] > 0.f ? dx[k] : 0.f;
      }
      __syncthreads();
      // layer 1
      coop_wgrad<CNT1>(X1, LDF64, X0, LD0, tid >> 1, (tid & 1) * CNT1, CNT1, gW1);
      float gb1 = 0.f;
      if (tid < 64) gb1 = coop_bias(X1, LDF64, tid);                                   // b1
      // bias accumulators: fold the per-sub-tile sums into two persistent registers per thread
      //   gB[0]: threads 0..63 -> b1+..., see flush for the exact mapping
      {
        float dE[E];
        row_dgrad<E>(x1, 64, sW1, E, dE);
        float gx[3] = {0.f, 0.f, 0.f};
        if (valid) {
          for (int l = 0; l < L; ++l) {
            const float g0 = dE[2 * l], g1 = dE[2 * l + 1];
            if (g0 != 0.f || g1 != 0.f) scatter_level(a.p.grad_table, lv, l, u, g0, g1);
            if (a.p.need_pose_grad) {
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                const float2 j = Jslot[(size_t)(l * 3 + d) * FT + tid];
                gx[d] = fmaf(g0, j.x, fmaf(g1, j.y, gx[d]));
              }
            }
          }
        }
        if (a.p.need_pose_grad) {
          float gtf[12];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float gi = 0.5f * gx[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) gtf[i * 4 + j] = gi * pc[j];
            gtf[i * 4 + 3] = gi;
          }
#pragma unroll
          for (int i = 0; i < 12; ++i) gtf[i] = warp_sum(gtf[i]);
          if (lane == 0 && rs.active && rs.frame != 0) {
#pragma unroll
            for (int i = 0; i < 12; ++i)
              if (gtf[i] != 0.f) red_add(a.p.grad_tf + (size_t)rs.frame * 12 + i, gtf[i]);
          }
        }
      }
      // the three 64-wide biases and b2 are flushed per sub-tile (cheap: <= 208 atomics per CTA per sub-tile)
      if (tid < 64) {
        if (gb1 != 0.f) red_add(a.p.grad_mlp + a.po[1] + tid, gb1);
        if (gb3 != 0.f) red_add(a.p.grad_mlp + a.po[5] + tid, gb3);
      } else if (tid < 80) {
        if (gb2 != 0.f) red_add(a.p.grad_mlp + a.po[3] + (tid - 64), gb2);
      }
    }
    __syncthreads();
    if (tid < R && sRay[tid].active) {
      RayS& r2 = sRay[tid];
      if (a.p.grad_feat) {
        for (int j = 0; j < a.p.ff; ++j)
          if (r2.dviews[j] != 0.f) red_add(a.p.grad_feat + (size_t)r2.frame * a.p.ff + j, r2.dviews[j]);
      }
      if (a.p.need_pose_grad && r2.frame != 0) {
        float gd[3];
        sh3_backward(r2.dw, r2.dviews + a.p.ff, gd);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float v = gd[i] * r2.u[j];
            if (v != 0.f) red_add(a.p.grad_tf + (size_t)r2.frame * 12 + i * 4 + j, v);
          }
      }
    }
    __syncthreads();
  }

  // ============ flush
  {
    float* G = a.p.grad_mlp;
    {  // W1: row o = tid>>1, k0 = (tid&1)*CNT1
      const int o = tid >> 1, k0 = (tid & 1) * CNT1;
#pragma unroll
      for (int j = 0; j < CNT1; ++j) if (gW1[j] != 0.f) red_add(G + a.po[0] + o * E + k0 + j, gW1[j]);
    }
    {  // W2: o = tid>>3, k0 = (tid&7)*8
      const int o = tid >> 3, k0 = (tid & 7) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (gW2[j] != 0.f) red_add(G + a.po[2] + o * 64 + k0 + j, gW2[j]);
    }
    {  // W3: o = tid>>1, k0 = (tid&1)*cnt3, only real columns k < K3
      const int o = tid >> 1, k0 = (tid & 1) * cnt3;
#pragma unroll
      for (int j = 0; j < 16; ++j) if (j < cnt3 && k0 + j < K3 && gW3[j] != 0.f) red_add(G + a.po[4] + o * K3 + k0 + j, gW3[j]);
    }
    {  // W4
      const int o = tid >> 1, k0 = (tid & 1) * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) if (gW4[j] != 0.f) red_add(G + a.po[6] + o * 64 + k0 + j, gW4[j]);
    }
    if (tid < 96) {
      const int o = tid >> 5, k0 = (tid & 31) * 2;
      if (gW5[0] != 0.f) red_add(G + a.po[8] + o * 64 + k0, gW5[0]);
      if (gW5[1] != 0.f) red_add(G + a.po[8] + o * 64 + k0 + 1, gW5[1]);
    }
    if (tid >= 96 && tid < 99 && gB[0] != 0.f) red_add(G + a.po[9] + (tid - 96), gB[0]);
    if (tid >= 64 && gB[1] != 0.f) red_add(G + a.po[7] + (tid - 64), gB[1]);
  }
  {
    loss_acc[0] = loss_acc[1] + loss_acc[2] + loss_acc[3] + loss_acc[4];
    float vals[7] = {loss_acc[0], loss_acc[1], loss_acc[2], loss_acc[3], loss_acc[4], n_valid_s, n_valid_r};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float v = warp_sum(vals[i]);
      if (lane == 0 && v != 0.f) red_add(a.p.losses + i, v);
    }
  }
}

size_t step_f32_smem(int E, int V, int grp_pts) { return (size_t)make_f32_plan(E, V, grp_pts).total; }

template <int E>
static int launch_f32(const StepArgs& a, int blocks, cudaStream_t st) {
  const size_t smem = step_f32_smem(E, a.V, a.R * a.Sp);
  cudaFuncSetAttribute(step_f32_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per-device attribute
  step_f32_kernel<E><<<blocks, FT, smem, st>>>(a);
  return check_launch("step_f32_kernel");
}

int step_f32_dispatch(const StepArgs& a, int blocks, cudaStream_t st) {
  if (a.E == 32) return launch_f32<32>(a, blocks, st);
  if (a.E == 8) return launch_f32<8>(a, blocks, st);
  if (a.E == 16) return launch_f32<16>(a, blocks, st);
  set_error("nof_step_fused(fp32): L*C=%d not built (8, 16 or 32)", a.E);
  return NOF_E_UNSUPPORTED;
}

}  // namespace nof
