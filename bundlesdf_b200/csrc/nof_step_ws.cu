// Fused forward + loss + backward of one train step, AMP policy — warp-specialised pipeline (one persistent CTA per SM).
//
// The reference runs this as ~300 PyTorch / cuBLAS / custom kernels (nerf_runner.py:1083-1088, 1227-1304, 1132-1169, 679-758;
// grid.py:34-99; gridencoder.cu:107-365). Here one CTA streams 128-point tiles through a ring of three shared-memory slots; five
// warp roles work on different tiles at the same time and hand tiles over with mbarriers (no block-wide barrier in steady state):
//
//   GATHER  (4 warps)  tile t+1.. : ray setup, sample points, multires hash-grid gather + d enc/d x  -> X0 (fp16 operand), z stash
//   MMA     (1 thread) tile t     : the ten chained GEMMs of NeRFSmall forward + dgrad on tcgen05 (M = 128 = the tile, accumulator in TMEM)
//   EPILOG  (8 warps)  tile t     : TMEM -> bias / ReLU / ReLU-mask -> next fp16 operand; compositing partial sums; loss seeds
//   WGRAD   (4 warps)  tile t     : weight gradients on mma.sync, register accumulators that persist over all tiles of the CTA
//   SCATTER (4 warps)  tile t-1   : run-merged grid-gradient reductions + pose Jacobian sums, per-ray finalisation
//
// Rays may span tiles (S up to 384 samples): tiles are forwarded in stream order and a tile's backward is issued as soon as every
// ray it touches has been forwarded completely (the RGB loss seed needs the composited colour of the whole ray); with S <= 128 a
// tile holds whole rays and forward / backward alternate. The per-ray inputs of the colour net (SH of the view direction, frame
// feature) are identical for all samples of a ray, so their product with W3 is computed ONCE per ray and enters layer 3 as a per-ray
// bias; the tile GEMM only carries the 15 geometry features (K = 16), and the matching gradients (dW3[:, views], d views) come from
// the per-ray column sums of dY3 that the weight-gradient warps produce anyway for the bias gradient.
#include "nof_mlp_image.cuh"

namespace nof {
namespace ws {
using namespace prim;

constexpr int PT = 128;                       // points per tile = UMMA M
constexpr int NS = 3;                         // tile slots in flight
constexpr int NRAY = 12;                      // ray-state ring: 3 tiles x 4 rays
using img::KG;
using img::VPAD;
constexpr int DES = PT + PT / 8;              // padded point stride of the transposed arrays (index pt + pt/8)
__device__ __forceinline__ int des_idx(int pt) { return pt + (pt >> 3); }

constexpr int NW_EPI = 8, NW_WG = 4, NW_GA = 4, NW_SC = 4;
constexpr int W_EPI0 = 0, W_WG0 = 8, W_GA0 = 12, W_SC0 = 16, W_MMA = 20;
constexpr int NT = 24 * 32;                   // six warpgroups; the last one holds the MMA issuer and three parked warps
constexpr int BAR_EPI = 1, BAR_GA = 2, BAR_SC = 3;
constexpr int REG_EPI = 80, REG_WG = 104, REG_GA = 104, REG_SC = 72, REG_MISC = 40;    // 768 threads x 80 registers re-dealt per role

enum { Q_START = 1, Q_END = 2 };

struct RayW {                                 // per-ray shared state (ring of NRAY)
  float dir[3], u[3], dw[3], gt[3];
  float depth, ray_w_base;
  float tf[12];
  int frame, ray, active;
  float sumw;                                 // sum of raw compositing weights (GATHER, all tiles of the ray)
  int anyvalid;
  float rgbacc[3];                            // sum of w_raw * sigmoid(rgb logit) over valid samples (EPILOG)
  float psum[6];                              // sum g_x, sum g_x * z (SCATTER) -> dL/dt, dL/dR
  float views[VPAD];                          // [frame feature | SH(dw)] rounded to fp16 (what the colour net sees under autocast)
  float vb[64];                               // W3[:, views] . views : the ray's contribution to the colour net's first pre-activation
  float cr[64];                               // sum over the ray's samples of dY3 (WGRAD) -> d views, dW3[:, views]
};

struct TileHdr {
  int grp;                                    // ray group of this tile, -1 = end of stream
  int qray[4];                                // ring index of the ray each 32-row quadrant belongs to
  int qs0[4];                                 // sample index of the quadrant's first row
  unsigned qvalid[4];                         // in-bounds mask of the quadrant's 32 samples
  int qflags[4];                              // Q_START: the ray's first samples, Q_END: its last
  int pad[3];
};

struct Plan {
  int w1, w2, w3g, w4, w5, bias, w3v, lv, img_bytes;          // operand image (one TMA bulk copy)
  int slot0, slot_bytes, x0, x1, xg, x3, x4, zs, out, hdr;    // slot-relative offsets
  int d_o, rays, bars, misc, gtab, total;
};
struct GemmTab {                               // one chained GEMM as the MMA issuer needs it (built once per CTA)
  uint32_t a_off;                              // A operand: byte offset inside the tile slot, or absolute (a_abs) shared-memory address
  uint32_t a_abs;
  uint32_t a_hi;                               // high word of the A descriptor (SBO, version)
  uint32_t b_lo, b_hi;                         // B descriptor for k-step 0
  uint32_t b_step;                             // added to b_lo per k-step
  uint32_t idesc;
  uint32_t nks;
};
__host__ __device__ inline Plan make_plan(int KE) {
  Plan s;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) / 128 * 128; return r; };
  const img::ImagePlan ip = img::make_image_plan(KE);
  s.w1 = ip.w1; s.w2 = ip.w2; s.w3g = ip.w3g; s.w4 = ip.w4; s.w5 = ip.w5; s.bias = ip.bias; s.w3v = ip.w3v; s.lv = ip.lv;
  s.img_bytes = ip.bytes;
  o = ip.bytes;
  s.slot0 = o;
  int q = 0;
  auto stake = [&](int bytes) { int r = q; q += (bytes + 127) / 128 * 128; return r; };
  s.x0 = stake(PT * KE * 2);
  s.x1 = stake(PT * 64 * 2);
  s.xg = stake(PT * KG * 2);
  s.x3 = stake(PT * 64 * 2);                  // x3|x4 also hold dEnc fp32 [KE][DES] (transposed) at the end of the backward
  s.x4 = stake(PT * 64 * 2);
  s.zs = stake(2 * DES * 4);                  // z and raw compositing weight of every point, padded order
  s.out = stake(PT * 4 * 2);                  // network output (rgb logits, sdf) as fp16 (its precision under autocast)
  s.hdr = stake((int)sizeof(TileHdr));
  s.slot_bytes = q;
  o += NS * q;
  s.d_o = take(PT * KG * 2);
  s.rays = take(NRAY * (int)sizeof(RayW));
  s.bars = take(24 * 8);
  s.misc = take(64);
  s.gtab = take(10 * (int)sizeof(GemmTab));
  s.total = o;
  return s;
}

// barrier slots inside Plan::bars
enum { B_IMG = 0, B_MMA = 1, B_OPND = 2, B_WGD = 3, B_GFULL = 4, B_SFREE = 7, B_SFULL = 10, B_RGA = 13, B_RSC = 14, B_REP = 15, B_DY = 16 };

// A waiter must observe EVERY phase of an mbarrier it uses (try_wait.parity cannot tell phase k from phase k-2): count the phases
// seen and wait them off one by one up to the phase that is needed.
struct PhaseWaiter {
  uint32_t seen;
  __device__ __forceinline__ bool until(uint64_t* bar, uint32_t need, volatile int* abort_flag) {
#pragma unroll 1
    while (seen < need) {
      if (!mbar_wait(bar, seen & 1u, abort_flag)) return false;
      ++seen;
    }
    return true;
  }
};

// ------------------------------------------------------------------------------------------------ small helpers
template <typename R>
__device__ __forceinline__ void world_point_w(const R& rs, float z, float pc[3], float x[3]) {       // nerf_runner.py:1083,1242-1243
#pragma unroll
  for (int j = 0; j < 3; ++j) pc[j] = rs.dir[j] * z;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    x[i] = fmaf(rs.tf[i * 4 + 2], pc[2], fmaf(rs.tf[i * 4 + 1], pc[1], rs.tf[i * 4 + 0] * pc[0])) + rs.tf[i * 4 + 3];
}

// Scalar part of the ray setup (one lane), nerf_runner.py:1045-1057,1282-1283.
__device__ __forceinline__ void setup_ray_w(RayW& rs, const StepArgs& a, int ray) {
  rs.ray = ray;
  rs.active = ray < a.p.N;
  rs.sumw = 0.f; rs.anyvalid = 0;
  rs.rgbacc[0] = rs.rgbacc[1] = rs.rgbacc[2] = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) rs.psum[i] = 0.f;
#pragma unroll
  for (int i = 0; i < VPAD; ++i) rs.views[i] = 0.f;
  if (!rs.active) {
    rs.frame = 0; rs.depth = 0.f; rs.ray_w_base = 0.f;
    for (int i = 0; i < 3; ++i) { rs.dir[i] = 0.f; rs.u[i] = 0.f; rs.dw[i] = 0.f; rs.gt[i] = 0.f; }
    for (int i = 0; i < 12; ++i) rs.tf[i] = 0.f;
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
  float v[MAX_V];
  for (int j = 0; j < a.p.ff; ++j) v[j] = a.p.feat[(size_t)rs.frame * a.p.ff + j];
  sh3(rs.dw, v + a.p.ff);
  for (int j = 0; j < a.V; ++j) rs.views[j] = __half2float(__float2half_rn(v[j]));           // autocast feeds the colour net fp16 inputs
}

// Loss seeds for one sample (nerf_runner.py:693-732, nerf_helpers.py:367-399): same arithmetic as nof_step_common.cuh loss_seeds,
// with the composited colour passed in (it is complete only once every tile of the ray has been forwarded).
__device__ __forceinline__ void loss_seeds_w(const StepArgs& a, float depth, const float gt[3], const float rgbm[3], const float out[4], float z,
                                             float w, bool valid, float ray_w, float d_out[4], float acc[5]) {
  const float sw = valid ? ray_w : 0.f;
  float rgb_s[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rgb_s[c] = sigmoidf_(out[c]);
    const float dmap = a.p.rgb_weight * 2.f * (rgbm[c] - gt[c]) * ray_w * a.inv_N3;
    d_out[c] = dmap * w * rgb_s[c] * (1.f - rgb_s[c]);
  }
  const float sdf = out[3];
  const float tr = step_trunc(a);
  const bool front = z < depth - tr;
  const bool back = z > depth + tr * a.p.neg_trunc_ratio;
  const bool valid_depth = (depth >= a.p.near_sc) && (depth <= a.p.far_sc);
  const bool m_sdf = !front && !back && valid_depth;
  float ds = 0.f;
  if (depth > a.p.far_sc && sdf < a.p.fs_sdf) {                 // uncertain free space (nerf_helpers.py:387-389)
    const float e = sdf - a.p.fs_sdf;
    acc[2] += a.p.fs_weight * 0.5f * e * e * sw * a.inv_NS;
    ds += a.p.fs_weight * e * sw * a.inv_NS;
  }
  if (front && depth <= a.p.far_sc && sdf < 1.f) {              // empty space in front of the surface (:391-393)
    acc[2] += a.p.fs_weight * a.p.empty_weight * fabsf(sdf - 1.f) * sw * a.inv_NS;
    ds += -a.p.fs_weight * a.p.empty_weight * sw * a.inv_NS;
  }
  if (m_sdf) {                                                  // truncated sdf near the surface (:395)
    const float e = (z + sdf * tr) - depth;
    acc[3] += a.p.trunc_weight * 0.5f * e * e * sw * a.inv_NS;
    ds += a.p.trunc_weight * e * tr * sw * a.inv_NS;
  }
  if (a.p.fs_rgb_weight > 0.f && front) {                       // nerf_runner.py:730-732
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e = rgb_s[c] - 1.f;
      acc[4] += a.p.fs_rgb_weight * e * e * sw * a.inv_NS3;
      d_out[c] += a.p.fs_rgb_weight * 2.f * e * sw * a.inv_NS3 * rgb_s[c] * (1.f - rgb_s[c]);
    }
  }
  d_out[3] = ds;
}

// The order in which tiles are forwarded (F) and back-propagated (B): identical, deterministic state machine in the MMA, EPILOG and
// WGRAD roles. B(b) is issued as soon as every ray touching tile b is forwarded; otherwise the next tile is forwarded.
struct Seq {
  int f, b, G, Sp;
  bool done;
  __device__ __forceinline__ bool can_b() const {
    if (b >= f) return false;
    const int tb = b % G, gb = b - tb;
    const int r_last = (tb * PT + PT - 1) / Sp;
    const int dep = ((r_last + 1) * Sp - 1) / PT;               // last tile (in the group) holding samples of the rays in tile b
    return f > gb + dep;
  }
};

// wgrad on mma.sync reading core-matrix buffers: dW[strip*16..+16][nt0*8 .. +NTU*8) += dY^T X over the 128 points of the tile.
// QSUM: the column sums of dY per 32-row quadrant (rows g8, g8+8 of the strip; lanes with t4 == 0) are added to cr of the quadrant's ray.
template <int NTU, bool QSUM>
__device__ __forceinline__ void wgrad_item(uint32_t dY, int Ky, uint32_t X, int Kx, int strip, int nt0, float (*acc)[4], float* bias2, bool do_bias,
                                           int lane, RayW* rays, const TileHdr* hdr) {
  const uint32_t ones = 0x3C003C00u;
  const int pa = (lane & 7) + (lane >> 4) * 8, oa = strip * 16 + ((lane >> 3) & 1) * 8;       // A: rows p, cols o (dY^T)
  const int pb = (lane & 7) + ((lane >> 3) & 1) * 8, ib = (lane >> 4) * 8;                      // B: rows p, cols i
#pragma unroll
  for (int q = 0; q < 4; ++q) {                                  // the tile's four 32-row quadrants, two k-steps each
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ks = 2 * q + h;
      uint32_t af[4];
      ldsm_x4_t(af, dY + cm_off(ks * 16 + pa, oa, Ky));
#pragma unroll
      for (int np = 0; np < (NTU + 1) / 2; ++np) {
        uint32_t bf[4];
        ldsm_x4_t(bf, X + cm_off(ks * 16 + pb, (nt0 + np * 2) * 8 + ib, Kx));
        mma16816(acc[np * 2], af, bf[0], bf[1]);
        if (np * 2 + 1 < NTU) mma16816(acc[np * 2 + 1], af, bf[2], bf[3]);
      }
      if (do_bias) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(t, af, ones, ones);
        q0 += t[0];
        q1 += t[2];
      }
    }
    bias2[0] += q0;
    bias2[1] += q1;
    if (QSUM && (lane & 3) == 0) {
      RayW& rq = rays[hdr->qray[q]];
      if (q0 != 0.f) atomicAdd(&rq.cr[strip * 16 + (lane >> 2)], q0);
      if (q1 != 0.f) atomicAdd(&rq.cr[strip * 16 + (lane >> 2) + 8], q1);
    }
  }
}
template <int CNT>
__device__ __forceinline__ void flush_item(float* G, int wofs, int bofs, int ncols, int nrows, int ld, int kofs, int strip, int nt0, const float (*acc)[4],
                                           const float* bias2, bool has_bias, int g8, int t4) {
#pragma unroll
  for (int nt = 0; nt < CNT; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = strip * 16 + g8 + h * 8, i = (nt0 + nt) * 8 + 2 * t4;
      const float v0 = acc[nt][h * 2], v1 = acc[nt][h * 2 + 1];
      if (o >= nrows || i >= ncols) continue;
      const size_t e = (size_t)wofs + (size_t)o * ld + kofs + i;
      if (i + 1 < ncols && (reinterpret_cast<uintptr_t>(G + e) & 7u) == 0u) {
        if (v0 != 0.f || v1 != 0.f) red_add_v2(G + e, v0, v1);
      } else {
        if (v0 != 0.f) red_add(G + e, v0);
        if (i + 1 < ncols && v1 != 0.f) red_add(G + e + 1, v1);
      }
    }
  if (has_bias && t4 == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = strip * 16 + g8 + h * 8;
      if (o < nrows && bias2[h] != 0.f) red_add(G + bofs + o, bias2[h]);
    }
  }
}
template <int N>
__device__ __forceinline__ void zero_acc(float (*acc)[4]) {
#pragma unroll
  for (int nt = 0; nt < N; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
}

// One GEMM D[128 x N] = A[128 x K] * B on the tensor core, issued by ONE thread; completion arrives on `bar`.
//   B_MN == 0: B = W stored [N rows x K] (forward: D = A W^T), K-major;  B_MN == 1: B = W stored [K rows x N] (dgrad: D = A W), MN-major view.
template <int N, int K, int B_MN>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, int KA, uint32_t b_addr, int KB, uint64_t* bar) {
  constexpr uint32_t idesc = umma_idesc(128, N, 0, B_MN);
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t ad = umma_desc(a_addr + ks * 256, 128, KA * 16);
    const uint64_t bd = B_MN ? umma_desc(b_addr + ks * 2 * (KB * 16), KB * 16, 128) : umma_desc(b_addr + ks * 256, 128, KB * 16);
    umma_f16(tmem_d, ad, bd, idesc, ks > 0 ? 1u : 0u);
  }
  umma_commit(bar);
}

// largest |x| of a packed half2 pair vs the fp16 range, without unpacking: inf / nan have all exponent bits set
__device__ __forceinline__ bool h2_bad(uint32_t w) { return ((w & 0x7C00u) == 0x7C00u) || ((w & 0x7C000000u) == 0x7C000000u); }

// ------------------------------------------------------------------------------------------------ the kernel
template <int KE_, bool POSE>
__global__ void __launch_bounds__(NT, 1) step_ws_kernel(const StepArgs a) {
  constexpr int KE = KE_;
  extern __shared__ __align__(128) unsigned char smem[];
  const Plan sp = make_plan(KE);
  const float* sB = reinterpret_cast<const float*>(smem + sp.bias);
  const __half* sW3v = reinterpret_cast<const __half*>(smem + sp.w3v);
  const LevelS& lv = *reinterpret_cast<const LevelS*>(smem + sp.lv);
  RayW* rays = reinterpret_cast<RayW*>(smem + sp.rays);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  volatile int* s_abort = reinterpret_cast<volatile int*>(smem + sp.misc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + sp.misc + 4);
  int* s_grp = reinterpret_cast<int*>(smem + sp.misc + 8);
  const uint32_t sbase = smem_u32(smem);
  unsigned char* pDO = smem + sp.d_o;
  const uint32_t aDO = sbase + sp.d_o;
  const uint32_t aW1 = sbase + sp.w1, aW2 = sbase + sp.w2, aW3 = sbase + sp.w3g, aW4 = sbase + sp.w4, aW5 = sbase + sp.w5;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = a.E, V = a.V, L = a.p.L, S = a.p.S, Sp = a.Sp, R = a.R;
  const int G = (R * Sp) / PT;                                   // tiles per ray group
  const float scale_ls = a.p.loss_scale ? *a.p.loss_scale : 1.0f;

  // ---- barriers, TMEM allocation, operand image
  if (tid == 0) {
    mbar_init(&bars[B_IMG], 1);
    mbar_init(&bars[B_MMA], 1);
    mbar_init(&bars[B_OPND], NW_EPI);
    mbar_init(&bars[B_WGD], NW_WG);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bars[B_GFULL + i], NW_GA);
      mbar_init(&bars[B_SFREE + i], NW_SC + NW_WG);
      mbar_init(&bars[B_SFULL + i], NW_EPI);
    }
    mbar_init(&bars[B_RGA], NW_GA);                              // role-internal barriers (bounded, unlike bar.sync)
    mbar_init(&bars[B_RSC], NW_SC);
    mbar_init(&bars[B_REP], NW_EPI);
    mbar_init(&bars[B_DY], NW_EPI);                              // EPILOG -> WGRAD: a dY operand of the backward is in shared memory
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (tid == 0) {
    mbar_expect_tx(&bars[B_IMG], (uint32_t)sp.img_bytes);
    tma_bulk_g2s(smem, a.wpack, (uint32_t)sp.img_bytes, &bars[B_IMG]);
  }
  bool ok = mbar_wait(&bars[B_IMG], 0);
  fence_async_smem();
  __syncthreads();

  auto slot_ptr = [&](int t) { return smem + sp.slot0 + (t % NS) * sp.slot_bytes; };
#ifdef NOF_WS_PROF   // role profile of CTA 0: cycles each warp spends waiting on mbarriers vs its whole role loop (profiles/ws_roles.py)
  long long prof_wait = 0, prof_t0 = clock64();
#define WS_WAIT(expr) do { const long long _t = clock64(); expr; prof_wait += clock64() - _t; } while (0)
#else
#define WS_WAIT(expr) do { expr; } while (0)
#endif
  // role-internal barrier over `bar` (count = warps of the role)
  auto role_sync = [&](uint64_t* bar, uint32_t& phase) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
    WS_WAIT(ok &= mbar_wait(bar, phase & 1u, s_abort));
    ++phase;
  };
  int* tile_ticket = reinterpret_cast<int*>(static_cast<char*>(a.wpack) + kWPackBytes - 16);

  const int wgp = warp >> 2;                                     // warpgroup = role (setmaxnreg is a warpgroup-wide instruction)
  if (wgp == W_GA0 / 4) {
    // ================================================================================== GATHER
    reg_inc<REG_GA>();
    const int gq = warp - W_GA0, row = gq * 32 + lane;
    uint32_t rph = 0;
    int t = 0;
    for (int gord = 0;; ++gord) {
      if (gq == 0 && lane == 0) *s_grp = (gord == 0) ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(tile_ticket, 1);
      role_sync(&bars[B_RGA], rph);
      const int grp = *s_grp;
      const bool last = grp >= a.n_groups;
      for (int j = 0; j < (last ? 1 : G); ++j, ++t) {
        unsigned char* sl = slot_ptr(t);
        TileHdr* hdr = reinterpret_cast<TileHdr*>(sl + sp.hdr);
        WS_WAIT(ok &= mbar_wait(&bars[B_SFREE + t % NS], ((uint32_t)(t / NS) & 1u) ^ 1u, s_abort));
        if (last) {
          if (gq == 0 && lane == 0) hdr->grp = -1;
          break;
        }
        const int o = j * PT + gq * 32;                         // offset of this quadrant in the group's point stream
        const int r = o / Sp, s0 = o - r * Sp;
        const int ring = (gord * R + r) % NRAY;
        RayW& rs = rays[ring];
        const bool start = s0 == 0, end = s0 + 32 >= Sp;
        if (start) {
          if (lane == 0) setup_ray_w(rs, a, grp * R + r);
          __syncwarp();
          for (int oo = lane; oo < 64; oo += 32) {              // the ray's share of the colour net's first layer
            float acc = 0.f;
            for (int v = 0; v < V; ++v) acc = fmaf(__half2float(sW3v[oo * VPAD + v]), rs.views[v], acc);
            rs.vb[oo] = acc;
            rs.cr[oo] = 0.f;
          }
        }
        if (lane == 0) {
          hdr->qray[gq] = ring; hdr->qs0[gq] = s0; hdr->qflags[gq] = (start ? Q_START : 0) | (end ? Q_END : 0);
          if (gq == 0) hdr->grp = grp;
        }
        role_sync(&bars[B_RGA], rph);                               // ray state of this tile visible to all four quadrants
        const int sidx = s0 + lane;
        const bool active = rs.active && sidx < S;
        const float z = active ? a.p.z_vals[(size_t)rs.ray * S + sidx] : 0.f;
        float pc[3], x[3], u[3];
        world_point_w(rs, z, pc, x);
        const bool valid = active && fabsf(x[0]) <= 1.f && fabsf(x[1]) <= 1.f && fabsf(x[2]) <= 1.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = (x[d] + 1.0f) * 0.5f;
        const float w_raw = active ? raw_weight(a, z, rs.depth) : 0.f;
        float* zs = reinterpret_cast<float*>(sl + sp.zs);
        zs[des_idx(row)] = z;
        zs[DES + des_idx(row)] = w_raw;
        const float wsum = warp_sum(w_raw);
        const unsigned anyv = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) {
          hdr->qvalid[gq] = anyv;
          if (wsum != 0.f) atomicAdd(&rs.sumw, wsum);
          if (anyv) atomicOr(&rs.anyvalid, 1);
        }
        unsigned char* pX0 = sl + sp.x0;
        __half2* Jslot = reinterpret_cast<__half2*>(a.jws) + ((size_t)blockIdx.x * NS + (t % NS)) * (MAX_L * 3) * PT;
#pragma unroll 1
        for (int c = 0; c < KE / 8; ++c) {                       // four levels = one 16-byte chunk of the operand row
          uint32_t w[4] = {0u, 0u, 0u, 0u};
          if (valid) {
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {                     // two levels in flight (16 loads), twice
#pragma unroll
              for (int k1 = 0; k1 < 2; ++k1) {
                const int k = 2 * k2 + k1, l = 4 * c + k;
                if (l < L) {
                  float enc[2], J[3][2];
                  gather_level<true, POSE>(a.p.table_f16, lv, l, u, enc, J);
                  if (POSE) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) Jslot[(size_t)(l * 3 + d) * PT + row] = __floats2half2_rn(J[d][0], J[d][1]);
                  }
                  w[k] = pack_h2(enc[0], enc[1]);
                }
              }
            }
          }
          *reinterpret_cast<uint4*>(pX0 + cm_off(row, c * 8, KE)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_GFULL + t % NS]);
      }
      if (last) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_GFULL + t % NS]);
        break;
      }
      if (!ok) break;
    }
  } else if (wgp == W_SC0 / 4) {
    // ================================================================================== SCATTER
    // Thread = 8 CONSECUTIVE samples of one ray x one level (two levels per thread, one after the other). Consecutive samples mostly fall
    // into the same grid cell, so the 8 corner contributions are summed in registers over the run and leave as ONE set of reductions per
    // run: ~3.5x fewer operations on the L2 atomic unit, which is what bounds the scatter (profiles/red_bench.cu).
    reg_dec<REG_SC>();
    const int sw = warp - W_SC0;
    uint32_t rph = 0;
    for (int t = 0;; ++t) {
      unsigned char* sl = slot_ptr(t);
      const TileHdr* hdr = reinterpret_cast<const TileHdr*>(sl + sp.hdr);
      WS_WAIT(ok &= mbar_wait(&bars[B_SFULL + t % NS], (uint32_t)(t / NS) & 1u, s_abort));
      if (!ok || hdr->grp < 0) break;
      const float* zs = reinterpret_cast<const float*>(sl + sp.zs);
      const float* dEb = reinterpret_cast<const float*>(sl + sp.x3);
      const __half2* Jslot = reinterpret_cast<const __half2*>(a.jws) + ((size_t)blockIdx.x * NS + (t % NS)) * (MAX_L * 3) * PT;
      const int sg = lane & 15, p0 = 8 * sg, q = p0 >> 5;
      RayW& r8 = rays[hdr->qray[q]];
      const unsigned vm = hdr->qvalid[q] >> (p0 & 31);
      float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int rep = 0; rep < 2; ++rep) {
        const int l = rep * 8 + 2 * sw + (lane >> 4);
        if (l >= L) continue;
        const float scale = lv.scale[l];
        const uint32_t off = lv.off[l];
        const float* dE0 = dEb + (size_t)(2 * l) * DES;
        float acc[8][2];
        uint32_t cpg[3] = {0u, 0u, 0u};
        bool have = false;
#pragma unroll 1
        for (int j = 0; j <= 8; ++j) {
          bool live = false;
          uint32_t pg[3] = {0u, 0u, 0u};
          float fr[3] = {0.f, 0.f, 0.f}, g0 = 0.f, g1 = 0.f;
          if (j < 8 && ((vm >> j) & 1u)) {
            const int qi = des_idx(p0 + j);
            g0 = dE0[qi];
            g1 = dE0[DES + qi];
            live = (g0 != 0.f || g1 != 0.f);
            if (live) {
              const float zq = zs[qi];
              float pc[3], x[3];
              world_point_w(r8, zq, pc, x);                      // same arithmetic as the gather: same cell, same weights
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                const float pp = fmaf((x[d] + 1.0f) * 0.5f, scale, 0.5f);
                const float fl = floorf(pp);
                pg[d] = (uint32_t)fl;
                fr[d] = pp - fl;
              }
              if (POSE) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  const float2 jj = __half22float2(Jslot[(size_t)(l * 3 + d) * PT + p0 + j]);
                  const float gx = 0.5f * fmaf(g0, jj.x, g1 * jj.y);
                  st[d] += gx;
                  st[3 + d] = fmaf(gx, zq, st[3 + d]);
                }
              }
            }
          }
          const bool newcell = live && (!have || pg[0] != cpg[0] || pg[1] != cpg[1] || pg[2] != cpg[2]);
          if (have && (newcell || j == 8)) {                     // the run ended: one set of reductions for all its samples
            uint32_t idx[8];
            corner_indices(lv, l, cpg, idx);
            float* base = a.p.grad_table + (size_t)off * 2;
#pragma unroll
            for (int k = 0; k < 8; ++k) red_add_v2(base + (size_t)idx[k] * 2, acc[k][0], acc[k][1]);
            have = false;
          }
          if (newcell) {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c][0] = acc[c][1] = 0.f;
            cpg[0] = pg[0]; cpg[1] = pg[1]; cpg[2] = pg[2];
            have = true;
          }
          if (live) {
            const float wx[2] = {1.f - fr[0], fr[0]}, wy[2] = {1.f - fr[1], fr[1]}, wz[2] = {1.f - fr[2], fr[2]};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float w = wx[c & 1] * wy[(c >> 1) & 1] * wz[(c >> 2) & 1];
              acc[c][0] = fmaf(w, g0, acc[c][0]);
              acc[c][1] = fmaf(w, g1, acc[c][1]);
            }
          }
        }
      }
      if (POSE) {
        // x = R (dir z) + t  =>  dL/dR[i][j] = dir[j] * sum gi z ,  dL/dt[i] = sum gi. Lanes of one quadrant (4 sample groups x 2 levels) first
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          float v = st[i];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if ((lane & 19) == 0 && v != 0.f) atomicAdd(&r8.psum[i], v);
        }
      }
      role_sync(&bars[B_RSC], rph);                                  // every scatter warp's sums are in
      // rays whose last samples are in this tile: pose / frame-feature gradients. Warp sw looks after quadrant sw.
      if (hdr->qflags[sw] & Q_END) {
        RayW& r2 = rays[hdr->qray[sw]];
        if (r2.active && (a.p.grad_feat || (POSE && r2.frame != 0))) {
          float dv[MAX_V];
          const float c0 = r2.cr[lane], c1 = r2.cr[lane + 32];
#pragma unroll
          for (int v = 0; v < MAX_V; ++v) {
            float pv = 0.f;
            if (v < V) pv = fmaf(c0, __half2float(sW3v[lane * VPAD + v]), c1 * __half2float(sW3v[(lane + 32) * VPAD + v]));
            dv[v] = warp_sum(pv);
          }
          if (lane == 0) {
            if (a.p.grad_feat)
              for (int j = 0; j < a.p.ff; ++j)
                if (dv[j] != 0.f) red_add(a.p.grad_feat + (size_t)r2.frame * a.p.ff + j, dv[j]);
            if (POSE && r2.frame != 0) {
              float gd[3];
              sh3_backward(r2.dw, dv + a.p.ff, gd);
#pragma unroll
              for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  const float v = fmaf(gd[i], r2.u[j], r2.psum[3 + i] * r2.dir[j]);
                  if (v != 0.f) red_add(a.p.grad_tf + (size_t)r2.frame * 12 + i * 4 + j, v);
                }
                if (r2.psum[i] != 0.f) red_add(a.p.grad_tf + (size_t)r2.frame * 12 + i * 4 + 3, r2.psum[i]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_SFREE + t % NS]);
    }
  } else if (wgp == W_WG0 / 4) {
    // ================================================================================== WGRAD
    reg_inc<REG_WG>();
    const int ww = warp - W_WG0;
    const int g8 = lane >> 2, t4 = lane & 3;
    float wg1[KE / 8][4], wg2[2][4], wg3[2][4], wg4[8][4], wg5[2][4];
    float wb1[2] = {0.f, 0.f}, wb2[2] = {0.f, 0.f}, wb3[2] = {0.f, 0.f}, wb4[2] = {0.f, 0.f}, wb5[2] = {0.f, 0.f};
    float w3v[9];
    zero_acc<KE / 8>(wg1); zero_acc<2>(wg2); zero_acc<2>(wg3); zero_acc<8>(wg4); zero_acc<2>(wg5);
#pragma unroll
    for (int k = 0; k < 9; ++k) w3v[k] = 0.f;
    Seq seq{0, 0, G, Sp, false};
    PhaseWaiter pdy{0};
    uint32_t nd = 0;                                              // dY operands announced so far (5 per backward)
    auto wg_done = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_WGD]);
    };
    while (ok) {
      if (seq.can_b()) {
        const int t = seq.b;
        unsigned char* sl = slot_ptr(t);
        const TileHdr* hdr = reinterpret_cast<const TileHdr*>(sl + sp.hdr);
        const uint32_t sa = smem_u32(sl);
        const uint32_t aX0 = sa + sp.x0, aX1 = sa + sp.x1, aXG = sa + sp.xg, aX3 = sa + sp.x3, aX4 = sa + sp.x4;
        WS_WAIT(ok &= pdy.until(&bars[B_DY], ++nd, s_abort));                      // dOut ready (seeds)
        wgrad_item<2, false>(aDO, KG, aX4, 64, 0, 2 * ww, wg5, wb5, ww == 0, lane, rays, hdr);
        wg_done();
        WS_WAIT(ok &= pdy.until(&bars[B_DY], ++nd, s_abort));                // dY4 ready
        wgrad_item<8, false>(aX4, 64, aX3, 64, ww, 0, wg4, wb4, true, lane, rays, hdr);
        wg_done();
        WS_WAIT(ok &= pdy.until(&bars[B_DY], ++nd, s_abort));                // dY3 ready
        {
          wgrad_item<2, true>(aX3, 64, aXG, KG, ww, 0, wg3, wb3, true, lane, rays, hdr);
          __syncwarp();
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {                          // rays complete in this tile: dW3[:, views] += c_r (x) views
            if (!(hdr->qflags[q] & Q_END)) continue;
            const RayW& rq = rays[hdr->qray[q]];
            const float c = rq.cr[ww * 16 + (lane & 15)];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
              const int v = (lane >> 4) + 2 * k;
              if (v < V) w3v[k] = fmaf(c, rq.views[v], w3v[k]);
            }
          }
        }
        wg_done();
        WS_WAIT(ok &= pdy.until(&bars[B_DY], ++nd, s_abort));                // dH2 ready
        wgrad_item<2, false>(aDO, KG, aX1, 64, 0, 2 * ww, wg2, wb2, ww == 0, lane, rays, hdr);
        wg_done();
        WS_WAIT(ok &= pdy.until(&bars[B_DY], ++nd, s_abort));                // dY1 ready
        wgrad_item<KE / 8, false>(aX1, 64, aX0, KE, ww, 0, wg1, wb1, true, lane, rays, hdr);
        wg_done();
        if (lane == 0) mbar_arrive(&bars[B_SFREE + t % NS]);    // this role no longer reads the slot
        ++seq.b;
      } else if (!seq.done) {
        const int t = seq.f;
        WS_WAIT(ok &= mbar_wait(&bars[B_GFULL + t % NS], (uint32_t)(t / NS) & 1u, s_abort));
        if (reinterpret_cast<const TileHdr*>(slot_ptr(t) + sp.hdr)->grp < 0) seq.done = true;
        else ++seq.f;
      } else {
        break;
      }
    }
    // ---- flush the weight gradients (once per CTA)
    {
      float* Gm = a.p.grad_mlp;
      const int K3 = V + 15;
      flush_item<KE / 8>(Gm, a.po[0], a.po[1], E, 64, E, 0, ww, 0, wg1, wb1, true, g8, t4);
      flush_item<2>(Gm, a.po[2], a.po[3], 64, 16, 64, 0, 0, 2 * ww, wg2, wb2, ww == 0, g8, t4);
      flush_item<2>(Gm, a.po[4], a.po[5], 15, 64, K3, V, ww, 0, wg3, wb3, true, g8, t4);
      flush_item<8>(Gm, a.po[6], a.po[7], 64, 64, 64, 0, ww, 0, wg4, wb4, true, g8, t4);
      flush_item<2>(Gm, a.po[8], a.po[9], 64, 3, 64, 0, 0, 2 * ww, wg5, wb5, ww == 0, g8, t4);
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int v = (lane >> 4) + 2 * k, o = ww * 16 + (lane & 15);
        if (v < V && w3v[k] != 0.f) red_add(Gm + a.po[4] + o * K3 + v, w3v[k]);
      }
    }
  } else if (wgp == W_MMA / 4) {
    // ================================================================================== MMA issuer (one thread; three parked warps)
    reg_dec<REG_MISC>();
    if (warp == W_MMA && lane == 0) {
      // The ten GEMMs of a tile as a table: F1..F5 (forward, B = W^T K-major), B5..B1 (dgrad, B = W as MN-major view). The issue loop below
      // is a few dozen instructions: this thread runs rarely and must not drag kilobytes of code through the instruction cache every time.
      GemmTab* gt = reinterpret_cast<GemmTab*>(smem + sp.gtab);
      {
        auto fill = [&](int i, uint32_t a_off, uint32_t a_abs, int KA, uint32_t b_addr, int KB, int N, int K, int b_mn) {
          GemmTab g;
          g.a_off = a_off; g.a_abs = a_abs;
          const uint64_t ad = umma_desc(0, 128, KA * 16);
          g.a_hi = (uint32_t)(ad >> 32);
          const uint64_t bd = b_mn ? umma_desc(b_addr, KB * 16, 128) : umma_desc(b_addr, 128, KB * 16);
          g.b_lo = (uint32_t)bd; g.b_hi = (uint32_t)(bd >> 32);
          g.b_step = b_mn ? (uint32_t)(2 * KB) : 16u;             // per k-step of 16: two 8-row groups of an MN-major buffer, or 256 bytes of a K-major one (>> 4)
          g.idesc = umma_idesc(128, N, 0, b_mn);
          g.nks = (uint32_t)(K / 16);
          gt[i] = g;
        };
        fill(0, sp.x0, 0, KE, aW1, KE, 64, KE, 0);
        fill(1, sp.x1, 0, 64, aW2, 64, 16, 64, 0);
        fill(2, sp.xg, 0, KG, aW3, KG, 64, KG, 0);
        fill(3, sp.x3, 0, 64, aW4, 64, 64, 64, 0);
        fill(4, sp.x4, 0, 64, aW5, 64, 16, 64, 0);
        fill(5, 0, aDO, KG, aW5, 64, 64, 16, 1);
        fill(6, sp.x4, 0, 64, aW4, 64, 64, 64, 1);
        fill(7, sp.x3, 0, 64, aW3, KG, KG, 64, 1);
        fill(8, 0, aDO, KG, aW2, 64, 64, 16, 1);
        fill(9, sp.x1, 0, 64, aW1, KE, KE, 64, 1);
      }
      Seq seq{0, 0, G, Sp, false};
      uint32_t ec = 0;
      PhaseWaiter pw{0};
      while (ok) {
        int g0, t;
        if (seq.can_b()) {
          t = seq.b++;
          g0 = 5;
          ec += 1;                                               // the seeds step precedes the first dgrad
        } else if (!seq.done) {
          t = seq.f;
          WS_WAIT(ok &= mbar_wait(&bars[B_GFULL + t % NS], (uint32_t)(t / NS) & 1u, s_abort));
          if (reinterpret_cast<const TileHdr*>(slot_ptr(t) + sp.hdr)->grp < 0) { seq.done = true; continue; }
          ++seq.f;
          g0 = 0;
        } else {
          break;
        }
        const uint32_t sa = smem_u32(slot_ptr(t));
#pragma unroll 1
        for (int g = g0; g < g0 + 5; ++g) {
          WS_WAIT(ok &= pw.until(&bars[B_OPND], ec, s_abort));   // every EPILOG step that precedes this GEMM is complete
          tc_fence_after();
          const GemmTab G_ = gt[g];
          const uint32_t a_addr = G_.a_abs ? G_.a_abs : sa + G_.a_off;
          uint32_t a_lo = ((a_addr >> 4) & 0x3FFFu) | (8u << 16), b_lo = G_.b_lo;      // LBO = 128 bytes
#pragma unroll 1
          for (uint32_t ks = 0; ks < G_.nks; ++ks) {
            umma_f16(tmem, ((uint64_t)G_.a_hi << 32) | a_lo, ((uint64_t)G_.b_hi << 32) | b_lo, G_.idesc, ks);
            a_lo += 16u;                                         // 256 bytes further along K
            b_lo += G_.b_step;
          }
          umma_commit(&bars[B_MMA]);
          ++ec;
        }
        WS_WAIT(ok &= pw.until(&bars[B_OPND], ec, s_abort));     // observe the op's last step too: no phase of B_OPND is ever skipped
      }
    }
    __syncwarp();
  } else {
    // ================================================================================== EPILOG (warps 0..7)
    const int quad = warp & 3, half = warp >> 2, row = quad * 32 + lane;
    const bool owner = half == 0;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    Seq seq{0, 0, G, Sp, false};
    uint32_t gc = 0, wc = 0, rph = 0;                              // GEMMs waited for, weight-gradient passes of finished backward ops
    PhaseWaiter pwg{0};
    float loss_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float n_valid_s = 0.f, n_valid_r = 0.f;
    bool overflow = false;
    auto mma_wait = [&]() {
      WS_WAIT(ok &= mbar_wait(&bars[B_MMA], gc & 1u, s_abort));
      ++gc;
      tc_fence_after();
    };
    auto step_done = [&](bool dy) {                               // publish smem writes to the tensor core, retire TMEM reads, signal
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[B_OPND]);
        if (dy) mbar_arrive(&bars[B_DY]);
      }
    };
    // relu(v + bias [+ per-ray bias]) of this thread's 32 columns -> fp16 operand row
    auto fwd_relu_store = [&](unsigned char* pX, const float* bias, const float* rb) {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = half * 32 + ch * 8 + 2 * j;
          const float x0 = v[ch * 8 + 2 * j] + bias[c] + rb[c], x1 = v[ch * 8 + 2 * j + 1] + bias[c + 1] + rb[c + 1];
          w[j] = pack_h2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
        }
        *reinterpret_cast<uint4*>(pX + cm_off(row, half * 32 + ch * 8, 64)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    };
    // dY = dgrad output masked by the ReLU of the stored activation, written in place of that activation
    auto bwd_mask_store = [&](unsigned char* pX, uint32_t wg_need) {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
      WS_WAIT(ok &= pwg.until(&bars[B_WGD], wg_need, s_abort));           // the weight-gradient warps are done reading this activation
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        unsigned char* p = pX + cm_off(row, half * 32 + ch * 8, 64);
        const uint4 m = *reinterpret_cast<const uint4*>(p);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2 mh = *reinterpret_cast<const __half2*>(&mw[j]);
          const uint32_t keep = __hgt2_mask(mh, __float2half2_rn(0.f));
          o[j] = pack_h2(v[ch * 8 + 2 * j], v[ch * 8 + 2 * j + 1]) & keep;
          overflow |= h2_bad(o[j]);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    };
    while (ok) {
      if (seq.can_b()) {
        // ---------------------------------------------------------------- backward of tile seq.b
        const int t = seq.b;
        unsigned char* sl = slot_ptr(t);
        const TileHdr* hdr = reinterpret_cast<const TileHdr*>(sl + sp.hdr);
        const RayW& rs = rays[hdr->qray[quad]];
        const bool valid = (hdr->qvalid[quad] >> lane) & 1u;
        const int sidx = hdr->qs0[quad] + lane;
        const bool active = rs.active && sidx < S;
        role_sync(&bars[B_REP], rph);                                // every EPILOG warp's compositing sums of the forwarded tiles are in
        WS_WAIT(ok &= pwg.until(&bars[B_WGD], wc, s_abort));              // (observe the previous backward's last weight-gradient phase)
        float dsdf_s = 0.f;
        if (owner) {
          const float* zs = reinterpret_cast<const float*>(sl + sp.zs);
          const float z = zs[des_idx(row)], w_raw = zs[DES + des_idx(row)];
          const uint2 oh = *reinterpret_cast<const uint2*>(sl + sp.out + row * 8);
          const float2 o01 = __half22float2(*reinterpret_cast<const __half2*>(&oh.x)), o23 = __half22float2(*reinterpret_cast<const __half2*>(&oh.y));
          const float out4[4] = {o01.x, o01.y, o23.x, o23.y};
          const float den = rs.sumw + 1e-10f;
          const float w = valid ? w_raw / den : 0.f;
          const float rgbm[3] = {rs.rgbacc[0] / den, rs.rgbacc[1] / den, rs.rgbacc[2] / den};
          const float ray_w = rs.ray_w_base * (rs.anyvalid ? 1.f : 0.f);
          float d_out[4];
          loss_seeds_w(a, rs.depth, rs.gt, rgbm, out4, z, w, valid, active ? ray_w : 0.f, d_out, loss_acc);
          if (!active) { d_out[0] = d_out[1] = d_out[2] = d_out[3] = 0.f; }
          if (valid) n_valid_s += 1.f;
          if (sidx == 0 && rs.active) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float dd = rgbm[c] - rs.gt[c]; e += dd * dd; }
            loss_acc[1] += a.p.rgb_weight * e * ray_w * a.inv_N3;
            if (rs.anyvalid && rs.ray_w_base != 0.f) n_valid_r += 1.f;
            if (a.p.rgb_map) {
#pragma unroll
              for (int c = 0; c < 3; ++c) a.p.rgb_map[(size_t)rs.ray * 3 + c] = rgbm[c];
            }
          }
          if (active) {
            const size_t pi = (size_t)rs.ray * S + sidx;
            if (a.p.raw) *reinterpret_cast<float4*>(a.p.raw + pi * 4) = make_float4(out4[0], out4[1], out4[2], out4[3]);
            if (a.p.valid_samples) a.p.valid_samples[pi] = valid ? 1 : 0;
            if (a.p.weights) a.p.weights[pi] = w;
          }
          dsdf_s = d_out[3] * scale_ls;
          const float s0 = d_out[0] * scale_ls, s1 = d_out[1] * scale_ls, s2 = d_out[2] * scale_ls;
          overflow |= !(fabsf(s0) <= 65504.f) || !(fabsf(s1) <= 65504.f) || !(fabsf(s2) <= 65504.f) || !(fabsf(dsdf_s) <= 65504.f);
          *reinterpret_cast<uint4*>(pDO + cm_off(row, 0, KG)) = make_uint4(pack_h2(s0, s1), pack_h2(s2, 0.f), 0u, 0u);
          *reinterpret_cast<uint4*>(pDO + cm_off(row, 8, KG)) = make_uint4(0u, 0u, 0u, 0u);
        }
        step_done(true);                                          // dOut visible
#pragma unroll 1
        for (int ly = 0; ly < 5; ++ly) {                          // layers 5, 4, 3, 2, 1
          mma_wait();
          if (ly == 2) {
            // ---- layer 3: d geo (15) -> dH2[1..15], dH2[0] = d sdf
            float v[8];
            TmemLd<8>::ld(trow + half * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = half * 8 + j;
              if (col < 15) {
                overflow |= !(fabsf(v[j]) <= 65504.f);
                *reinterpret_cast<__half*>(pDO + cm_off(row, 1 + col, KG)) = __float2half_rn(v[j]);
              }
            }
            if (owner) *reinterpret_cast<__half*>(pDO + cm_off(row, 0, KG)) = __float2half_rn(dsdf_s);
            WS_WAIT(ok &= pwg.until(&bars[B_WGD], wc + 3, s_abort));      // every phase of a barrier must be observed before the next one can complete
          } else if (ly == 4) {
            // ---- layer 1: dEnc fp32, transposed [column][DES], over X3|X4 (dead: wgrad3 / wgrad4 were observed done at layer 2)
            constexpr int NH = KE / 2;
            float v[NH];
            TmemLd<NH>::ld(trow + half * NH, v);
            float* dE = reinterpret_cast<float*>(sl + sp.x3) + des_idx(row);
#pragma unroll
            for (int j = 0; j < NH; ++j) dE[(half * NH + j) * DES] = v[j];
          } else {
            // ---- layers 5, 4, 2: dY over the stored activation (X4, X3, X1) once the weight-gradient warps have read it
            bwd_mask_store(sl + (ly == 0 ? sp.x4 : ly == 1 ? sp.x3 : sp.x1), wc + (ly == 0 ? 1u : ly == 1 ? 2u : 4u));
          }
          if (ly < 4) step_done(true);
        }
        wc += 5;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[B_SFULL + t % NS]);
          mbar_arrive(&bars[B_OPND]);
        }
        ++seq.b;
      } else if (!seq.done) {
        // ---------------------------------------------------------------- forward of tile seq.f
        const int t = seq.f;
        unsigned char* sl = slot_ptr(t);
        WS_WAIT(ok &= mbar_wait(&bars[B_GFULL + t % NS], (uint32_t)(t / NS) & 1u, s_abort));
        const TileHdr* hdr = reinterpret_cast<const TileHdr*>(sl + sp.hdr);
        if (hdr->grp < 0) { seq.done = true; continue; }
        RayW& rs = rays[hdr->qray[quad]];
        const bool valid = (hdr->qvalid[quad] >> lane) & 1u;
#pragma unroll 1
        for (int ly = 0; ly < 5; ++ly) {
          mma_wait();
          if (ly == 1) {
            // L2: 64 -> 16 (sdf | geo 15)
            float v[8];
            TmemLd<8>::ld(trow + half * 8, v);
            unsigned char* pXG = sl + sp.xg;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = half * 8 + j;
              const __half hv = __float2half_rn(v[j] + sB[64 + col]);
              if (col == 0) *reinterpret_cast<__half*>(sl + sp.out + row * 8 + 6) = hv;
              else *reinterpret_cast<__half*>(pXG + cm_off(row, col - 1, KG)) = hv;
            }
            if (!owner) *reinterpret_cast<__half*>(pXG + cm_off(row, 15, KG)) = __float2half_rn(0.f);
          } else if (ly == 4) {
            // L5: 64 -> 3; compositing partial sums of this quadrant
            if (owner) {
              float v[8];
              TmemLd<8>::ld(trow, v);
              const float w_raw = reinterpret_cast<const float*>(sl + sp.zs)[DES + des_idx(row)];
              __half h[3];
              float pr[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                h[c] = __float2half_rn(v[c] + sB[208 + c]);
                pr[c] = warp_sum(valid ? w_raw * sigmoidf_(__half2float(h[c])) : 0.f);
              }
              *reinterpret_cast<__half2*>(sl + sp.out + row * 8) = __halves2half2(h[0], h[1]);
              *reinterpret_cast<__half*>(sl + sp.out + row * 8 + 4) = h[2];
              if (lane == 0 && (pr[0] != 0.f || pr[1] != 0.f || pr[2] != 0.f)) {
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(&rs.rgbacc[c], pr[c]);
              }
            }
          } else {
            // L1 (E -> 64), L3 (geo 15 + the ray's view/feature share as a bias -> 64), L4 (64 -> 64): ReLU -> next operand
            fwd_relu_store(sl + (ly == 0 ? sp.x1 : ly == 2 ? sp.x3 : sp.x4), sB + (ly == 0 ? 0 : ly == 2 ? 80 : 144), ly == 2 ? rs.vb : sB + 216);
          }
          step_done(false);
        }
        ++seq.f;
      } else {
        break;
      }
    }
    // end of stream: pass the sentinel tile on to the scatter warps
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[B_SFULL + seq.f % NS]);
    {
      loss_acc[0] = loss_acc[1] + loss_acc[2] + loss_acc[3] + loss_acc[4];
      float vals[7] = {loss_acc[0], loss_acc[1], loss_acc[2], loss_acc[3], loss_acc[4], n_valid_s, n_valid_r};
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const float v = warp_sum(vals[i]);
        if (lane == 0 && v != 0.f) red_add(a.p.losses + i, v);
      }
      const unsigned ov = __ballot_sync(0xffffffffu, overflow);
      if (lane == 0 && ov && a.p.found_inf) atomicCAS(a.p.found_inf, 0, 1);
    }
  }

#ifdef NOF_WS_PROF
  if (blockIdx.x == 0 && lane == 0) {
    long long* pr = reinterpret_cast<long long*>(static_cast<char*>(a.wpack) + 24 * 1024) + warp * 2;
    pr[0] = clock64() - prof_t0;
    pr[1] = prof_wait;
  }
#endif
  // ---- every role is done (or gave up): report, release TMEM
  if (!ok) {
    *s_abort = 1;
    if (a.p.found_inf) atomicExch(a.p.found_inf, 2);              // a bounded wait timed out: results are invalid
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace ws

size_t step_ws_smem(int KE) { return (size_t)ws::make_plan(KE).total; }
size_t step_ws_jscratch(int blocks) { return (size_t)blocks * ws::NS * (MAX_L * 3) * ws::PT * 4; }

// Layout of the ray groups for the streaming kernel: Sp = padded samples per ray (multiple of 32), R rays per group, R*Sp a multiple of 128,
// no ray spanning more than three tiles.
bool step_ws_tiling(int S, int* Sp_out, int* R_out) {
  int Sp, R;
  if (S <= 32) { Sp = 32; R = 4; }
  else if (S <= 64) { Sp = 64; R = 2; }
  else if (S <= 128) { Sp = 128; R = 1; }
  else if (S <= 192) { Sp = 192; R = 2; }
  else if (S <= 256) { Sp = 256; R = 1; }
  else if (S <= 320) { Sp = 320; R = 2; }
  else if (S <= 384) { Sp = 384; R = 1; }
  else return false;
  *Sp_out = Sp; *R_out = R;
  return true;
}

template <int KE, bool POSE>
static int launch_ws(const StepArgs& a, int blocks, cudaStream_t st) {
  const size_t smem = step_ws_smem(KE);
  static_assert(sizeof(ws::TileHdr) == 80, "TileHdr layout");
  cudaFuncSetAttribute(ws::step_ws_kernel<KE, POSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per-device attribute
  const ws::Plan sp = ws::make_plan(KE);
  if ((size_t)sp.img_bytes + 16 > kWPackBytes) { set_error("nof_step_fused(ws): operand image too large"); return NOF_E_INVALID; }
  img::pack_image_kernel<KE><<<(sp.img_bytes / 4 + 255) / 256, 256, 0, st>>>(a);
  ws::step_ws_kernel<KE, POSE><<<blocks, ws::NT, smem, st>>>(a);
  return check_launch("step_ws_kernel");
}

int step_ws_dispatch(const StepArgs& a, int blocks, cudaStream_t st) {
  const bool pose = a.p.need_pose_grad != 0;
  if (a.KE == 32) return pose ? launch_ws<32, true>(a, blocks, st) : launch_ws<32, false>(a, blocks, st);
  if (a.KE == 16) return pose ? launch_ws<16, true>(a, blocks, st) : launch_ws<16, false>(a, blocks, st);
  set_error("nof_step_fused(amp, ws): unsupported KE=%d", a.KE);
  return NOF_E_UNSUPPORTED;
}

}  // namespace nof
