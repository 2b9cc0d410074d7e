// Shared device code of the fused train-step kernels (fp16 tensor-core policy in nof_step_amp.cu, fp32 policy in
// nof_step_f32.cu): per-ray setup, point generation, hash-grid gather with d enc/d x, compositing weights,
// loss seeds (dL/draw), grid-gradient scatter and the pose Jacobian reductions.
//
// Thread mapping: a CTA owns R whole rays; thread = one sample point (Sp = S rounded up to 32 lanes per ray, so a
// warp never straddles two rays). Because a CTA sees complete rays, forward, loss and backward run back to back in
// one kernel and no per-point intermediate ever leaves the SM except the d enc/d x block J, which goes to a per-CTA
// scratch slot that is reused every tile and therefore lives in L2.
#pragma once
#include "nof_common.cuh"

namespace nof {

constexpr int KC = 32;          // padded width of the colour-net input [views(ff+9) | geo(15) | 0...]
constexpr int MAX_L = 16;
constexpr int MAX_R = 4;        // rays per CTA
constexpr int MAX_V = 17;       // ff <= 8

struct StepArgs {
  NofStep p;
  int E, V, KE;                 // enc width L*C, view width ff+9, enc width padded to 16
  int po[10];                   // element offsets of W1 b1 W2 b2 W3 b3 W4 b4 W5 b5 in the packed block
  int R, Sp, n_groups;          // rays per CTA, lanes per ray, number of ray groups
  float inv_N3, inv_NS, inv_NS3;
  void* jws;                    // workspace: per-CTA Jacobian scratch
  void* wpack;                  // workspace: fp16 MLP operands pre-packed for the tcgen05 kernel (kWPackBytes)
  int count_only;               // eikonal: first pass of the tile kernel — forward up to the sdf, count the samples with sdf < 1, nothing else
};
constexpr size_t kWPackBytes = 32 * 1024;

struct RayS {                   // per-ray shared state
  float dir[3], u[3];           // camera-frame dir (non-unit) and its unit vector
  float dw[3];                  // world-frame unit view dir R*u
  float gt[3];
  float depth, ray_w_base;      // first_frame_weight or 1, times (type==0)
  float tf[12];
  int frame, ray, active;
  float sumw;                   // sum of raw compositing weights over the ray's samples
  float rgb[3];                 // composited colour
  int anyvalid;
  float views[MAX_V];           // [frame feature | SH(dw)] fp32
  float dviews[KC];             // sum over samples of dL/d(colour-net input)[0..V)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// SH degree 3 (nerf_helpers.py:72-85) and its Jacobian contraction: g_dir += J^T g_sh
__device__ __forceinline__ void sh3(const float d[3], float* out) {
  const float x = d[0], y = d[1], z = d[2];
  out[0] = 0.28209479177387814f;
  out[1] = -0.4886025119029199f * y;
  out[2] = 0.4886025119029199f * z;
  out[3] = -0.4886025119029199f * x;
  out[4] = 1.0925484305920792f * (x * y);
  out[5] = -1.0925484305920792f * (y * z);
  out[6] = 0.31539156525252005f * (2.0f * (z * z) - x * x - y * y);
  out[7] = -1.0925484305920792f * (x * z);
  out[8] = 0.5462742152960396f * (x * x - y * y);
}
__device__ __forceinline__ void sh3_backward(const float d[3], const float* g, float gd[3]) {
  const float x = d[0], y = d[1], z = d[2];
  const float C1 = 0.4886025119029199f, A = 1.0925484305920792f, Bc = 0.31539156525252005f, Cc = 0.5462742152960396f;
  gd[0] = -C1 * g[3] + A * y * g[4] - 2.f * Bc * x * g[6] - A * z * g[7] + 2.f * Cc * x * g[8];
  gd[1] = -C1 * g[1] + A * x * g[4] - A * z * g[5] - 2.f * Bc * y * g[6] - 2.f * Cc * y * g[8];
  gd[2] = C1 * g[2] - A * y * g[5] + 4.f * Bc * z * g[6] - A * x * g[7];
}

// Per-level constants staged in shared memory once per CTA.
struct LevelS {
  float scale[MAX_L];
  uint32_t res1[MAX_L];         // resolution + 1
  uint32_t hsize[MAX_L];
  uint32_t off[MAX_L];
  uint32_t dense[MAX_L];
  uint32_t hmask[MAX_L];        // hsize-1 when hsize is a power of two (always for hashed levels built by grid.py), else 0
};

__device__ __forceinline__ void init_levels(LevelS& lv, const StepArgs& a) {
  if (threadIdx.x < a.p.L) {
    LevelGeom g = level_geom3(threadIdx.x, a.p.S_log2, a.p.H, a.p.offsets);
    lv.scale[threadIdx.x] = g.scale;
    lv.res1[threadIdx.x] = g.resolution + 1u;
    lv.hsize[threadIdx.x] = g.hashmap_size;
    lv.off[threadIdx.x] = g.offset;
    lv.dense[threadIdx.x] = g.dense;
    lv.hmask[threadIdx.x] = (g.hashmap_size & (g.hashmap_size - 1)) == 0 ? g.hashmap_size - 1 : 0u;
  }
}

// `%` for table sizes that are not powers of two: never the case for grid.py tables, so keep the division sequence out of line.
static __device__ __noinline__ uint32_t umod_cold(uint32_t a, uint32_t b) { return a % b; }

// Table entries of the 8 corners of one cell (gridencoder.cu:62-84): corner c = (c&1, (c>>1)&1, (c>>2)&1). The level is
// warp-uniform, so the dense / hashed split is a real branch; both forms are incremental (uint32 wrap-around arithmetic is
// exact: (y+1)*p == y*p + p mod 2^32), ~20 integer instructions per cell instead of ~45.
__device__ __forceinline__ void corner_indices(const LevelS& lv, int l, const uint32_t pg[3], uint32_t idx[8]) {
  const uint32_t r1 = lv.res1[l];
  if (lv.dense[l]) {                                              // x + y*r1 + z*r1^2, < hsize by construction (:70-73)
    const uint32_t sz = r1 * r1;
    const uint32_t b00 = pg[0] + pg[1] * r1 + pg[2] * sz, b10 = b00 + r1, b01 = b00 + sz, b11 = b10 + sz;
    idx[0] = b00; idx[1] = b00 + 1u; idx[2] = b10; idx[3] = b10 + 1u;
    idx[4] = b01; idx[5] = b01 + 1u; idx[6] = b11; idx[7] = b11 + 1u;
  } else {                                                        // x ^ y*2654435761 ^ z*805459861 (:78-82)
    const uint32_t hx0 = pg[0], hx1 = pg[0] + 1u;
    const uint32_t hy0 = pg[1] * 2654435761u, hy1 = hy0 + 2654435761u;
    const uint32_t hz0 = pg[2] * 805459861u, hz1 = hz0 + 805459861u;
    const uint32_t a00 = hy0 ^ hz0, a10 = hy1 ^ hz0, a01 = hy0 ^ hz1, a11 = hy1 ^ hz1;
    idx[0] = hx0 ^ a00; idx[1] = hx1 ^ a00; idx[2] = hx0 ^ a10; idx[3] = hx1 ^ a10;
    idx[4] = hx0 ^ a01; idx[5] = hx1 ^ a01; idx[6] = hx0 ^ a11; idx[7] = hx1 ^ a11;
    const uint32_t hm = lv.hmask[l];
    if (hm) {                                                     // `% 2^k` without the integer-division sequence
#pragma unroll
      for (int c = 0; c < 8; ++c) idx[c] &= hm;
    } else {
      const uint32_t hs = lv.hsize[l];
#pragma unroll
      for (int c = 0; c < 8; ++c) idx[c] = umod_cold(idx[c], hs);
    }
  }
}

template <bool HALF> struct TableT;
template <> struct TableT<true> {
  using vec = __half2;
  static __device__ __forceinline__ float2 load(const void* table, uint32_t entry) {
#ifdef NOF_EXP_STAGE_L0   // ablation only (profiles/README.md): `table` may point into shared memory, so no ld.global.nc
    const __half2 h = *(reinterpret_cast<const __half2*>(table) + entry);
#else
    const __half2 h = __ldg(reinterpret_cast<const __half2*>(table) + entry);
#endif
    return __half22float2(h);
  }
};
template <> struct TableT<false> {
  using vec = float2;
  static __device__ __forceinline__ float2 load(const void* table, uint32_t entry) {
    return __ldg(reinterpret_cast<const float2*>(table) + entry);
  }
};

// Setup of the R rays of one group (threads 0..R-1), nerf_runner.py:1045-1057,1282-1283.
__device__ __forceinline__ void setup_ray(RayS& rs, const StepArgs& a, int ray) {
  rs.ray = ray;
  rs.active = ray < a.p.N;
  rs.sumw = 0.f; rs.rgb[0] = rs.rgb[1] = rs.rgb[2] = 0.f; rs.anyvalid = 0;
#pragma unroll
  for (int i = 0; i < KC; ++i) rs.dviews[i] = 0.f;
  if (!rs.active) {
    rs.frame = 0; rs.depth = 0.f; rs.ray_w_base = 0.f;
    for (int i = 0; i < 3; ++i) { rs.dir[i] = 0.f; rs.u[i] = 0.f; rs.dw[i] = 0.f; rs.gt[i] = 0.f; }
    for (int i = 0; i < 12; ++i) rs.tf[i] = 0.f;
    for (int i = 0; i < MAX_V; ++i) rs.views[i] = 0.f;
    return;
  }
  const float* row = a.p.rays + (size_t)ray * a.p.ray_dim;
  const float dx = row[0], dy = row[1], dz = row[2];
  rs.dir[0] = dx; rs.dir[1] = dy; rs.dir[2] = dz;
  const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
  rs.u[0] = dx / nrm; rs.u[1] = dy / nrm; rs.u[2] = dz / nrm;
  rs.gt[0] = row[3]; rs.gt[1] = row[4]; rs.gt[2] = row[5];
  rs.depth = row[6];
  rs.frame = min(max((int)row[8], 0), a.p.F - 1);
  const float type = row[9];
  rs.ray_w_base = (type == 0.f) ? ((rs.frame == 0) ? a.p.first_frame_weight : 1.0f) : 0.f;   // nerf_runner.py:693-698,723
  const float* T = a.p.tf + (size_t)rs.frame * 12;
#pragma unroll
  for (int i = 0; i < 12; ++i) rs.tf[i] = T[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) rs.dw[i] = T[i * 4 + 0] * rs.u[0] + T[i * 4 + 1] * rs.u[1] + T[i * 4 + 2] * rs.u[2];
  for (int j = 0; j < a.p.ff; ++j) rs.views[j] = a.p.feat[(size_t)rs.frame * a.p.ff + j];
  sh3(rs.dw, rs.views + a.p.ff);
}

// Truncation of this step: a device scalar when the schedule anneals it (NofStep.trunc_ptr), else the launch constant.
__device__ __forceinline__ float step_trunc(const StepArgs& a) { return a.p.trunc_ptr ? __ldg(a.p.trunc_ptr) : a.p.trunc; }

// Raw (un-normalised) compositing weight, nerf_runner.py:1152-1159.
__device__ __forceinline__ float raw_weight(const StepArgs& a, float z, float depth) {
  if (depth > a.p.far_sc) return 0.f;
  const float trunc = step_trunc(a);
  const float s = (depth - z) / trunc;
  float w = sigmoidf_(s * a.p.sdf_lambda) * sigmoidf_(-s * a.p.sdf_lambda);
  const float dz = z - depth;
  const bool m = (dz <= trunc * a.p.neg_trunc_ratio) && (dz >= -trunc);
  return m ? w : 0.f;
}

// World point of this thread's sample: x = R (dir*z) + t (nerf_runner.py:1083,1242-1243).
__device__ __forceinline__ void world_point(const RayS& rs, float z, float pc[3], float x[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) pc[j] = rs.dir[j] * z;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    x[i] = fmaf(rs.tf[i * 4 + 2], pc[2], fmaf(rs.tf[i * 4 + 1], pc[1], rs.tf[i * 4 + 0] * pc[0])) + rs.tf[i * 4 + 3];
}

// One level of the multires gather for one point: enc (C=2) and, if WANT_J, J[d][c] = d enc/d u (gridencoder.cu:155-245).
// Trilinear interpolation as nested lerps (x, then y, then z): the differences each lerp needs ARE the Jacobian terms, so
// enc + J cost ~52 flops instead of ~104 for the weight-per-corner form (differs from it by fp32 rounding only).
template <bool HALF, bool WANT_J>
__device__ __forceinline__ void gather_level(const void* table, const LevelS& lv, int l, const float u[3], float enc[2],
                                             float J[3][2]) {
  const float scale = lv.scale[l];
  const uint32_t off = lv.off[l];
  float fr[3];
  uint32_t pg[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float p = fmaf(u[d], scale, 0.5f);
    const float fl = floorf(p);
    pg[d] = (uint32_t)fl;
    fr[d] = p - fl;
  }
  uint32_t idx[8];
  corner_indices(lv, l, pg, idx);
  float2 f[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#ifdef NOF_EXP_NO_GATHER  // profiling experiment only: same index arithmetic, no table loads
    f[c] = make_float2(__uint_as_float((off + idx[c]) & 0x3fu) * 1e30f, 0.25f);
#else
    f[c] = TableT<HALF>::load(table, off + idx[c]);
#endif
  }
  // x: pairs (0,1) (2,3) (4,5) (6,7) -> lx[yz], dx[yz]
  float2 dx[4], lx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    dx[k] = make_float2(f[2 * k + 1].x - f[2 * k].x, f[2 * k + 1].y - f[2 * k].y);
    lx[k] = make_float2(fmaf(fr[0], dx[k].x, f[2 * k].x), fmaf(fr[0], dx[k].y, f[2 * k].y));
  }
  // y: (lx[0],lx[1]) at z=0, (lx[2],lx[3]) at z=1
  float2 dy[2], ly[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    dy[k] = make_float2(lx[2 * k + 1].x - lx[2 * k].x, lx[2 * k + 1].y - lx[2 * k].y);
    ly[k] = make_float2(fmaf(fr[1], dy[k].x, lx[2 * k].x), fmaf(fr[1], dy[k].y, lx[2 * k].y));
  }
  const float2 dz = make_float2(ly[1].x - ly[0].x, ly[1].y - ly[0].y);
  enc[0] = fmaf(fr[2], dz.x, ly[0].x);
  enc[1] = fmaf(fr[2], dz.y, ly[0].y);
  if (WANT_J) {
    const float wy1 = fr[1], wy0 = 1.f - fr[1], wz1 = fr[2], wz0 = 1.f - fr[2];
    const float w00 = wy0 * wz0, w10 = wy1 * wz0, w01 = wy0 * wz1, w11 = wy1 * wz1;      // [y][z] ; dx index k = y + 2 z
    J[0][0] = scale * fmaf(w11, dx[3].x, fmaf(w01, dx[2].x, fmaf(w10, dx[1].x, w00 * dx[0].x)));
    J[0][1] = scale * fmaf(w11, dx[3].y, fmaf(w01, dx[2].y, fmaf(w10, dx[1].y, w00 * dx[0].y)));
    J[1][0] = scale * fmaf(wz1, dy[1].x, wz0 * dy[0].x);
    J[1][1] = scale * fmaf(wz1, dy[1].y, wz0 * dy[0].y);
    J[2][0] = scale * dz.x;
    J[2][1] = scale * dz.y;
  }
}

// Scatter dL/d enc of one level into the fp32 gradient table (gridencoder.cu:250-336, fp32 accumulate instead of
// fp16 atomics) — one vectorised reduction per corner.
__device__ __forceinline__ void scatter_level(float* grad_table, const LevelS& lv, int l, const float u[3], float g0, float g1) {
  const float scale = lv.scale[l];
  const uint32_t off = lv.off[l];
  float fr[3];
  uint32_t pg[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float p = fmaf(u[d], scale, 0.5f);
    const float fl = floorf(p);
    pg[d] = (uint32_t)fl;
    fr[d] = p - fl;
  }
  uint32_t idx[8];
  corner_indices(lv, l, pg, idx);
  const float wx0 = 1.f - fr[0], wx1 = fr[0];
  const float wy[2] = {1.f - fr[1], fr[1]}, wz[2] = {1.f - fr[2], fr[2]};
  float* base = grad_table + (size_t)off * 2;
  const bool pairable = (off & 1u) == 0u;                          // always true for grid.py's offsets (multiples of 8)
  // The x / x+1 corners of a (y,z) pair are neighbours in memory whenever their entries differ only in bit 0 (dense levels:
  // even entry; hashed levels: even x, because (x^A) and ((x|1)^A) differ in bit 0 only; level offsets are multiples of 8):
  // then ONE 16-byte reduction carries both. The L2 atomic unit is the binding resource of the scatter and charges per lane
  // operation, not per byte (profiles/red_bench.cu).
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t e0 = idx[2 * k], e1 = idx[2 * k + 1];
    const float wyz = wy[k & 1] * wz[k >> 1];
    const float w0 = wx0 * wyz, w1 = wx1 * wyz;
#ifdef NOF_EXP_NO_RED     // profiling experiment only (profiles/README.md): what the kernel costs without the reductions
    if (w0 * g0 == 123456.f) red_add_v2(base + (size_t)e0 * 2, w0 * g0, w1 * g1);
#else
    if ((e0 ^ e1) == 1u && pairable) {
      const bool sw = (e0 & 1u) != 0u;                           // e0 is the odd entry: swap the halves
      const float a0 = sw ? w1 : w0, a1 = sw ? w0 : w1;
      red_add_v4(base + (size_t)(e0 & ~1u) * 2, a0 * g0, a0 * g1, a1 * g0, a1 * g1);
    } else {
      red_add_v2(base + (size_t)e0 * 2, w0 * g0, w0 * g1);
      red_add_v2(base + (size_t)e1 * 2, w1 * g0, w1 * g1);
    }
#endif
  }
}

// scatter_level plus the eikonal term's table gradient: corner k additionally receives 1/2 scale (g . dw_k/df) q_c, with dw_k/df_d the
// derivative of the trilinear weight (w_k = prod_d (c_d ? f_d : 1 - f_d)), q = d sdf / d enc of this level, g = dL/dn.
__device__ __forceinline__ void scatter_level_eik(float* grad_table, const LevelS& lv, int l, const float u[3], float g0, float g1, float q0, float q1,
                                                  const float eg[3]) {
  const float scale = lv.scale[l];
  const uint32_t off = lv.off[l];
  float fr[3];
  uint32_t pg[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float p = fmaf(u[d], scale, 0.5f);
    const float fl = floorf(p);
    pg[d] = (uint32_t)fl;
    fr[d] = p - fl;
  }
  uint32_t idx[8];
  corner_indices(lv, l, pg, idx);
  float* base = grad_table + (size_t)off * 2;
  const float hs = 0.5f * scale;
  const bool pairable = (off & 1u) == 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {                               // (y, z) pairs; the x / x+1 corners share one 16-byte reduction when adjacent
    const int cy = k & 1, cz = k >> 1;
    const float wy = cy ? fr[1] : 1.f - fr[1], wz = cz ? fr[2] : 1.f - fr[2];
    const float wyz = wy * wz;
    const float dyz = (cy ? eg[1] : -eg[1]) * wz + (cz ? eg[2] : -eg[2]) * wy;       // d(wy wz)/df . g over y, z
    float v[2][2];
#pragma unroll
    for (int cx = 0; cx < 2; ++cx) {
      const float wx = cx ? fr[0] : 1.f - fr[0];
      const float w = wx * wyz;
      const float dw = (cx ? eg[0] : -eg[0]) * wyz + wx * dyz;
      const float kq = hs * dw;
      v[cx][0] = fmaf(kq, q0, w * g0);
      v[cx][1] = fmaf(kq, q1, w * g1);
    }
    const uint32_t e0 = idx[2 * k], e1 = idx[2 * k + 1];
    if ((e0 ^ e1) == 1u && pairable) {
      const bool sw = (e0 & 1u) != 0u;                        // e0 is the odd entry: swap the halves
      red_add_v4(base + (size_t)(e0 & ~1u) * 2, sw ? v[1][0] : v[0][0], sw ? v[1][1] : v[0][1], sw ? v[0][0] : v[1][0], sw ? v[0][1] : v[1][1]);
    } else {
      red_add_v2(base + (size_t)e0 * 2, v[0][0], v[0][1]);
      red_add_v2(base + (size_t)e1 * 2, v[1][0], v[1][1]);
    }
  }
}

// Loss seeds for one sample (nerf_runner.py:693-732, nerf_helpers.py:367-399). Inputs: network output (rgb logits,
// sdf), normalised compositing weight w (already 0 for invalid samples), ray weight. Returns dL/draw (unscaled) and
// accumulates the per-term loss values into acc[5] = {total, rgb(unused here), fs, sdf, fs_rgb}.
__device__ __forceinline__ void loss_seeds(const StepArgs& a, const RayS& rs, const float out[4], float z, float w, bool valid,
                                           float ray_w, float d_out[4], float acc[5]) {
  const float sw = valid ? ray_w : 0.f;
  const float depth = rs.depth;
  float rgb_s[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rgb_s[c] = sigmoidf_(out[c]);
    const float dmap = a.p.rgb_weight * 2.f * (rs.rgb[c] - rs.gt[c]) * ray_w * a.inv_N3;
    d_out[c] = dmap * w * rgb_s[c] * (1.f - rgb_s[c]);
  }
  const float sdf = out[3];
  const float tr = step_trunc(a);
  const bool front = z < depth - tr;
  const bool back = z > depth + tr * a.p.neg_trunc_ratio;
  const bool valid_depth = (depth >= a.p.near_sc) && (depth <= a.p.far_sc);
  const bool m_sdf = !front && !back && valid_depth;
  float ds = 0.f;
  // uncertain free space (nerf_helpers.py:387-389)
  if (depth > a.p.far_sc && sdf < a.p.fs_sdf) {
    const float e = sdf - a.p.fs_sdf;
    acc[2] += a.p.fs_weight * 0.5f * e * e * sw * a.inv_NS;
    ds += a.p.fs_weight * e * sw * a.inv_NS;
  }
  // empty space in front of the surface (nerf_helpers.py:391-393)
  if (front && depth <= a.p.far_sc && sdf < 1.f) {
    acc[2] += a.p.fs_weight * a.p.empty_weight * fabsf(sdf - 1.f) * sw * a.inv_NS;
    ds += -a.p.fs_weight * a.p.empty_weight * sw * a.inv_NS;
  }
  // truncated sdf near the surface (nerf_helpers.py:395)
  if (m_sdf) {
    const float e = (z + sdf * tr) - depth;
    acc[3] += a.p.trunc_weight * 0.5f * e * e * sw * a.inv_NS;
    ds += a.p.trunc_weight * e * tr * sw * a.inv_NS;
  }
  if (a.p.fs_rgb_weight > 0.f && front) {       // nerf_runner.py:730-732
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e = rgb_s[c] - 1.f;
      acc[4] += a.p.fs_rgb_weight * e * e * sw * a.inv_NS3;
      d_out[c] += a.p.fs_rgb_weight * 2.f * e * sw * a.inv_NS3 * rgb_s[c] * (1.f - rgb_s[c]);
    }
  }
  d_out[3] = ds;
}

}  // namespace nof
