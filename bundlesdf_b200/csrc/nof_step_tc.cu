// Fused forward + loss + backward of one train step, AMP policy, tcgen05 edition (tile PT = 128 points, S <= 128).
//
// Same structure as nof_step_amp.cu (two threads per point, every warp walks the phases of a tile together: one instruction stream per SM
// at a time, which is what keeps this ~8 k-instruction kernel inside the instruction cache — see profiles/README.md for what happens when
// the phases run as concurrent warp roles instead), but the ten chained GEMMs of the MLP forward and dgrad run on the 5th-generation
// tensor cores:
//   * one elected thread issues tcgen05.mma (M=128 = the whole tile, N = layer width, K = 16 per instruction) with both
//     operands read straight from shared memory through UMMA descriptors; the accumulator lives in TMEM (64 columns);
//   * completion is signalled by tcgen05.commit on an mbarrier; every thread then pulls ITS row (32 columns of it) out of
//     TMEM with tcgen05.ld.32x32b, applies bias / ReLU / the ReLU mask, converts to fp16 and writes the next operand;
//   * activations and weights are stored in the canonical no-swizzle "core matrix" layout (8 rows x 16 bytes contiguous),
//     which serves as K-major A/B for the forward GEMMs AND as MN-major B for the dgrad GEMMs (W is never transposed), and
//     which ldmatrix(.trans) can read for the wgrad GEMMs that stay on mma.sync with register accumulators;
//   * each dgrad MMA is issued asynchronously BEFORE the warps start the wgrad of the same layer, so the tensor-core
//     generations overlap.
// Ablation on B200 (profiles/README.md): the mma.sync + ldmatrix MLP phases cost 174 of 286 us per C2 launch; this kernel
// replaces 156 of the 264 mma.sync and 152 of the 240 ldmatrix per warp and tile by 26 tcgen05.mma per CTA and tile.
#include "nof_mlp_image.cuh"

namespace nof {
namespace tc {
using namespace prim;
using img::KG;
using img::VPAD;

// Bounded wait with the polling loop INLINE: every thread of the CTA waits ten times per tile with ~60 live registers (the weight-gradient
// accumulators); an out-of-line call there costs more than the loop's footprint saves in this lock-step kernel (unlike nof_step_ws.cu).
__device__ __forceinline__ bool mbar_wait_inl(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (int it = 0; it < (1 << 22); ++it)
    if (mbar_try(bar, parity)) return true;
  return false;
}

constexpr int PT = 128, NT = 256, NWARP = 8;
constexpr int DES = PT + PT / 8;                 // padded point stride of the transposed dEnc / z arrays: index pt + pt/8
__device__ __forceinline__ int des_idx(int pt) { return pt + (pt >> 3); }

struct RayT : RayS {                          // + the per-ray share of the colour net (nof_mlp_image.cuh)
  float vb[64];                               // W3[:, views] . views
  float cr[64];                               // sum over the ray's samples of dY3 -> d views, dW3[:, views]
};

struct Plan {
  int w1, w2, w3, w4, w5, bias, w3v, lv, img_bytes, x0, x1, xc, x3, x4, d_o, out, zs, rays, bar, tmem, l0, total;
};
__host__ __device__ inline Plan make_plan(int KE) {
  Plan s;
  const img::ImagePlan ip = img::make_image_plan(KE);
  s.w1 = ip.w1; s.w2 = ip.w2; s.w3 = ip.w3g; s.w4 = ip.w4; s.w5 = ip.w5; s.bias = ip.bias; s.w3v = ip.w3v; s.lv = ip.lv; s.img_bytes = ip.bytes;
  int o = ip.bytes;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) / 128 * 128; return r; };
  s.x0 = take(PT * KE * 2);
  s.x1 = take(PT * 64 * 2);
  s.xc = take(PT * KG * 2);                  // geo features (15 + pad): the only per-sample input of the colour net
  s.x3 = take(PT * 64 * 2);                 // x3|x4 also hold dEnc fp32 [KE][DES] (transposed) at the end of the backward
  s.x4 = take(PT * 64 * 2);
  s.d_o = take(PT * 16 * 2);
  s.out = take(PT * 4 * 4);
  s.zs = take(5 * DES * 4);                  // z, valid flag and u[3] of every point, in the scatter's padded order
  s.rays = take(2 * MAX_R * (int)sizeof(RayT));    // two tiles' rays: the next tile's are staged while this one computes
  s.bar = take(64);                          // [0] TMA staging barrier, [1] MMA completion barrier
  s.tmem = take(16);
#ifdef NOF_EXP_STAGE_L0   // ablation: level 0 of the fp16 table (17^3 entries) staged in shared memory by TMA
  s.l0 = take(4920 * 4);
#else
  s.l0 = o;
#endif
  s.total = o;
  return s;
}

// wgrad on mma.sync reading core-matrix buffers (same balanced split as nof_step_amp.cu): dW[strip*16..+16][nt0*8..] += dY^T X
// RSUM: the column sums of dY (which the bias gradient needs anyway) are also added, ray by ray (ks_per_ray k-steps each), to cr of
// the tile's rays: lanes with t4 == 0 hold rows g8 and g8 + 8 of the strip.
// and dW3[:, views] += c_r (x) views goes into w3v (rows g8 / g8 + 8 of the strip, view columns t4, t4 + 4, ...: every lane of a quad holds the
// same row sums, so the four lanes split the columns).
template <int NTU, bool RSUM = false>
__device__ __forceinline__ void wgrad_item(uint32_t dY, int Ky, uint32_t X, int Kx, int strip, int nt0, float (*acc)[4], float* bias2,
                                           bool do_bias, int lane, RayT* rays = nullptr, int ks_per_ray = 8, float (*w3v)[2] = nullptr, int V = 0) {
#ifdef NOF_EXP_NO_WGRAD
  return;
#endif
  const uint32_t ones = 0x3C003C00u;
  const int pa = (lane & 7) + (lane >> 4) * 8, oa = strip * 16 + ((lane >> 3) & 1) * 8;       // A: rows p, cols o (dY^T)
  const int pb = (lane & 7) + ((lane >> 3) & 1) * 8, ib = (lane >> 4) * 8;                      // B: rows p, cols i
  float r0 = 0.f, r1 = 0.f;
  for (int ks = 0; ks < PT / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4_t(a, dY + cm_off(ks * 16 + pa, oa, Ky));
#pragma unroll
    for (int np = 0; np < (NTU + 1) / 2; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, X + cm_off(ks * 16 + pb, (nt0 + np * 2) * 8 + ib, Kx));
      mma16816(acc[np * 2], a, b[0], b[1]);
      if (np * 2 + 1 < NTU) mma16816(acc[np * 2 + 1], a, b[2], b[3]);
    }
    if (do_bias) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(t, a, ones, ones);
      bias2[0] += t[0];
      bias2[1] += t[2];
      if (RSUM) {
        r0 += t[0];
        r1 += t[2];
        if ((ks + 1) % ks_per_ray == 0) {
          RayT& rq = rays[ks / ks_per_ray];
          if ((lane & 3) == 0) {                                // one strip <-> one warp: plain stores, no atomics
            rq.cr[strip * 16 + (lane >> 2)] = r0;
            rq.cr[strip * 16 + (lane >> 2) + 8] = r1;
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int v = (lane & 3) + 4 * k;
            if (v < V) {
              const float hv = __half2float(__float2half_rn(rq.views[v]));
              w3v[k][0] = fmaf(r0, hv, w3v[k][0]);
              w3v[k][1] = fmaf(r1, hv, w3v[k][1]);
            }
          }
          r0 = r1 = 0.f;
        }
      }
    }
  }
}
template <int NS, int NTL>
struct WSplit {
  static constexpr int ideal = (NS * NTL + NWARP - 1) / NWARP;
  static constexpr int CNT = ideal <= 1 ? 1 : (ideal <= 2 ? (NTL % 2 == 0 ? 2 : NTL) : (NTL % 4 == 0 ? 4 : NTL));
  static constexpr int GROUPS = NTL / CNT;
  static constexpr int ITEMS = NS * GROUPS;
  static_assert(CNT <= 4 && NTL % CNT == 0, "wgrad split");
};
template <int CNT>
__device__ __forceinline__ void flush_item(float* G, int wofs, int bofs, int ncols, int nrows, int strip, int nt0, const float (*acc)[4],
                                           const float* bias2, bool has_bias, int g8, int t4, int ld = -1, int kofs = 0) {
  if (ld < 0) ld = ncols;
#pragma unroll
  for (int nt = 0; nt < CNT; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = strip * 16 + g8 + h * 8, i = (nt0 + nt) * 8 + 2 * t4;
      const float v0 = acc[nt][h * 2], v1 = acc[nt][h * 2 + 1];
      if (o >= nrows || i >= ncols) continue;
      const size_t e = (size_t)wofs + (size_t)o * ld + kofs + i;
      if (i + 1 < ncols && (reinterpret_cast<uintptr_t>(G + e) & 7u) == 0u) {                       // the fragment's two columns in one 8-byte reduction
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

// One GEMM D[128 x N] (+)= A[128 x K] * B on the tensor core. A: core-matrix K-major buffer with row length KA.
//   B_MN == 0: B = W stored [N rows x K] (forward:  D = A W^T), K-major
//   B_MN == 1: B = W stored [K rows x N] (dgrad:    D = A W),   MN-major view of the same buffer, row length = N
// Issued by ONE thread; completion arrives on `bar`.
template <int N, int K, int B_MN>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, int KA, uint32_t b_addr, int KB, uint64_t* bar) {
  constexpr uint32_t idesc = umma_idesc(128, N, 0, B_MN);
#ifdef NOF_EXP_NO_TC
  return;
#endif
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t ad = umma_desc(a_addr + ks * 256, 128, KA * 16);                      // two 16-byte k-chunks, 128 B apart
    const uint64_t bd = B_MN ? umma_desc(b_addr + ks * 2 * (KB * 16), KB * 16, 128)      // MN-major: LBO = k-group stride, SBO = n-chunk stride
                             : umma_desc(b_addr + ks * 256, 128, KB * 16);               // K-major
    umma_f16(tmem_d, ad, bd, idesc, ks > 0 ? 1u : 0u);
  }
  umma_commit(bar);
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int KE_>
__global__ void __launch_bounds__(NT, 2) step_tc_kernel(const StepArgs a) {
  constexpr int KE = KE_;
  extern __shared__ __align__(128) unsigned char smem[];
  const Plan sp = make_plan(KE);
  float* sB = reinterpret_cast<float*>(smem + sp.bias);
  float* sOut = reinterpret_cast<float*>(smem + sp.out);
  float* sZ = reinterpret_cast<float*>(smem + sp.zs);
  RayT* sRayBase = reinterpret_cast<RayT*>(smem + sp.rays);
  const LevelS& lv = *reinterpret_cast<const LevelS*>(smem + sp.lv);
  const __half* sW3v = reinterpret_cast<const __half*>(smem + sp.w3v);
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + sp.bar);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + sp.tmem);
  int* s_next = reinterpret_cast<int*>(s_tmem + 1);          // [2]: the tile after the current one, by buffer parity
  const uint32_t sbase = smem_u32(smem);
  const uint32_t aW1 = sbase + sp.w1, aW2 = sbase + sp.w2, aW3 = sbase + sp.w3, aW4 = sbase + sp.w4, aW5 = sbase + sp.w5;
  const uint32_t aX0 = sbase + sp.x0, aX1 = sbase + sp.x1, aXC = sbase + sp.xc, aX3 = sbase + sp.x3, aX4 = sbase + sp.x4, aDO = sbase + sp.d_o;
  unsigned char* pX0 = smem + sp.x0; unsigned char* pX1 = smem + sp.x1; unsigned char* pXC = smem + sp.xc;
  unsigned char* pX3 = smem + sp.x3; unsigned char* pX4 = smem + sp.x4; unsigned char* pDO = smem + sp.d_o;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = a.E, V = a.V, L = a.p.L;
  const float scale_ls = a.p.loss_scale ? *a.p.loss_scale : 1.0f;

  // ---- barriers, TMEM allocation (warp 0), parameter staging
  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  // MLP operands: already fp16, core-matrix ordered and zero-padded (pack_mlp_kernel) — one bulk copy, no conversion
  const uint32_t pack_bytes = (uint32_t)sp.img_bytes;
#ifdef NOF_EXP_STAGE_L0
  const uint32_t l0_bytes = (uint32_t)(a.p.offsets[1] - a.p.offsets[0]) * 4u;          // level 0 of the table, 16-byte multiple (grid.py pads to 8 entries)
  const unsigned char* table0 = smem + sp.l0 - (size_t)a.p.offsets[0] * 4;              // so that table0[off0 + idx] is the staged copy
  if (tid == 0) {
    mbar_expect_tx(bar_tma, pack_bytes + l0_bytes);
    tma_bulk_g2s(smem + sp.w1, a.wpack, pack_bytes, bar_tma);
    tma_bulk_g2s(smem + sp.l0, static_cast<const unsigned char*>(a.p.table_f16) + (size_t)a.p.offsets[0] * 4, l0_bytes, bar_tma);
  }
#else
  if (tid == 0) {
    mbar_expect_tx(bar_tma, pack_bytes);
    tma_bulk_g2s(smem + sp.w1, a.wpack, pack_bytes, bar_tma);
  }
#endif
  bool tc_ok = mbar_wait_inl(bar_tma, 0);
  fence_async_smem();
  __syncthreads();

  using S1 = WSplit<4, KE / 8>;
  using S2 = WSplit<1, 8>;
  using S3 = WSplit<4, KG / 8>;
  using S4 = WSplit<4, 8>;
  using S5 = WSplit<1, 8>;
  float wg1[S1::CNT][4], wg2[S2::CNT][4], wg3[S3::CNT][4], wg4[S4::CNT][4], wg5[S5::CNT][4];
  float wb1[2] = {0.f, 0.f}, wb2[2] = {0.f, 0.f}, wb3[2] = {0.f, 0.f}, wb4[2] = {0.f, 0.f}, wb5[2] = {0.f, 0.f};
  zero_acc<S1::CNT>(wg1); zero_acc<S2::CNT>(wg2); zero_acc<S3::CNT>(wg3); zero_acc<S4::CNT>(wg4); zero_acc<S5::CNT>(wg5);
  float w3v[5][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};       // dW3[:, views] rows g8 / g8+8 of this warp's strip, columns t4 + 4k
  float loss_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float n_valid_s = 0.f, n_valid_r = 0.f;
  bool overflow = false;
  uint32_t phase = 0;                        // parity of bar_mma

  const int Sp = a.Sp, R = a.R, S = a.p.S;
  const int half = tid / PT;                 // gather/scatter: which half of the levels; epilogues: which half of the columns
  const int pt = tid - half * PT;            // = 32*(warp%4) + lane: the TMEM lane this thread may read
  const int rl = pt / Sp, sidx = pt - rl * Sp;
  const int l_beg = half ? (L + 1) >> 1 : 0, l_end = half ? L : (L + 1) >> 1;
  const bool owner = half == 0;
  __half2* Jslot = reinterpret_cast<__half2*>(a.jws) + (size_t)blockIdx.x * (MAX_L * 3) * PT;
  const int g8 = lane >> 2, t4 = lane & 3;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);     // TMEM address of this warp's lane quadrant

  // wait for the MMA generation, make TMEM readable
  auto mma_wait = [&]() {
#ifdef NOF_EXP_NO_TC
    return;
#endif
    tc_ok &= mbar_wait_inl(bar_mma, phase);
    phase ^= 1u;
    tc_fence_after();
  };
  // publish this thread's shared-memory writes to the tensor core, retire its TMEM reads, block barrier
  auto sync_for_mma = [&]() {
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    tc_fence_after();
  };

  // Tiles (ray groups) are handed out dynamically: first one = blockIdx.x, then an atomic ticket (zeroed by pack_mlp_kernel).
  // 2048 tiles over 296 CTAs of uneven cost (invalid samples skip gather and scatter): -8.5 % vs the static round-robin. Starting
  // the second CTA of each SM half a tile late was tried and does not help.
  int* tile_ticket = reinterpret_cast<int*>(static_cast<char*>(a.wpack) + kWPackBytes - 16);
  // Ray state is double-buffered: while tile t computes, the warps that idle through its compositing / loss phases stage the rays of
  // tile t+1 (one thread per ray: dependent global loads, ~1.5 us that used to sit in front of every tile behind a block barrier).
  int grp = blockIdx.x, cur = 0;
  if (tid < R && grp < a.n_groups) setup_ray(sRayBase[tid], a, grp * R + tid);
  __syncthreads();
  while (grp < a.n_groups) {
    RayT* sRay = sRayBase + cur * MAX_R;
    RayT* sRayNext = sRayBase + (cur ^ 1) * MAX_R;
    // ============ 1. the next tile's ticket; this tile's rays were staged during the previous tile (or above)
    if (tid == NT - 1) s_next[cur] = (int)gridDim.x + atomicAdd(tile_ticket, 1);
    // the rays' share of the colour net's first layer (their view / frame-feature inputs are the same for every sample): W3[:, views] . views,
    // on the fp16-rounded inputs autocast would feed; read as a per-ray bias by the layer-3 epilogue, several barriers from here
    for (int i = tid; i < R * 64; i += NT) {
      RayT& rq = sRay[i >> 6];
      const int o = i & 63;
      float acc = 0.f;
      for (int v = 0; v < V; ++v) acc = fmaf(__half2float(sW3v[o * VPAD + v]), __half2float(__float2half_rn(rq.views[v])), acc);
      rq.vb[o] = acc;
      rq.cr[o] = 0.f;
    }
    const RayT& rs = sRay[rl];
    const bool active = rs.active && sidx < S;
    const float z = active ? a.p.z_vals[(size_t)rs.ray * S + sidx] : 0.f;
    float pc[3], x[3], u[3];
    world_point(rs, z, pc, x);
    const bool valid = active && fabsf(x[0]) <= 1.f && fabsf(x[1]) <= 1.f && fabsf(x[2]) <= 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = (x[d] + 1.0f) * 0.5f;
    const float w_raw = active ? raw_weight(a, z, rs.depth) : 0.f;
    if (owner) {
      const int qi = des_idx(pt);                           // for the scatter, whose threads own other points
      sZ[qi] = z;
      sZ[DES + qi] = valid ? 1.f : 0.f;
#pragma unroll
      for (int d = 0; d < 3; ++d) sZ[(2 + d) * DES + qi] = u[d];
      const float ws = warp_sum(w_raw);
      const unsigned anyv = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) {
        if (ws != 0.f) atomicAdd(&sRay[rl].sumw, ws);
        if (anyv) atomicOr(&sRay[rl].anyvalid, 1);
      }
    }
    // ============ 2. this thread's half of the gather
#ifdef NOF_EXP_NO_GATHER
    if (false) {
#else
    if (valid) {
#endif
#pragma unroll 2
      for (int l = l_beg; l < l_end; ++l) {
        float enc[2], J[3][2];
        if (a.p.need_pose_grad) {
#ifdef NOF_EXP_STAGE_L0
          gather_level<true, true>(l == 0 ? (const void*)table0 : a.p.table_f16, lv, l, u, enc, J);
#else
          gather_level<true, true>(a.p.table_f16, lv, l, u, enc, J);
#endif
#pragma unroll
          for (int d = 0; d < 3; ++d) Jslot[(size_t)(l * 3 + d) * PT + pt] = __floats2half2_rn(J[d][0], J[d][1]);
        } else {
          gather_level<true, false>(a.p.table_f16, lv, l, u, enc, J);
        }
        *reinterpret_cast<uint32_t*>(pX0 + cm_off(pt, 2 * l, KE)) = pack_h2(enc[0], enc[1]);
      }
    } else {
      for (int l = l_beg; l < l_end; ++l) *reinterpret_cast<uint32_t*>(pX0 + cm_off(pt, 2 * l, KE)) = 0u;
    }
    if (owner) for (int j = E; j < KE; j += 2) *reinterpret_cast<uint32_t*>(pX0 + cm_off(pt, j, KE)) = 0u;
    sync_for_mma();
    // ============ 3. MLP forward: five tcgen05 GEMMs, epilogue = this thread's row, its half of the columns
    // ---- L1: E -> 64, ReLU
    if (tid == 0) issue_gemm<64, KE, 0>(tmem, aX0, KE, aW1, KE, bar_mma);
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = half * 32 + ch * 8 + 2 * j;
          w[j] = pack_h2(fmaxf(v[ch * 8 + 2 * j] + sB[c], 0.f), fmaxf(v[ch * 8 + 2 * j + 1] + sB[c + 1], 0.f));
        }
        *reinterpret_cast<uint4*>(pX1 + cm_off(pt, half * 32 + ch * 8, 64)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    sync_for_mma();
    // ---- L2: 64 -> 16 (sdf | geo 15), no activation
    if (tid == 0) issue_gemm<16, 64, 0>(tmem, aX1, 64, aW2, 64, bar_mma);
    mma_wait();
    {
      float v[8];
      TmemLd<8>::ld(trow + half * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = half * 8 + j;
        const __half hv = __float2half_rn(v[j] + sB[64 + col]);
        if (col == 0) sOut[pt * 4 + 3] = __half2float(hv);
        else *reinterpret_cast<__half*>(pXC + cm_off(pt, col - 1, KG)) = hv;
      }
      if (!owner) *reinterpret_cast<__half*>(pXC + cm_off(pt, 15, KG)) = __float2half_rn(0.f);
    }
    sync_for_mma();
    // ---- L3: geo 15 (+ the ray's view / feature share as a bias) -> 64, ReLU
    if (tid == 0) issue_gemm<64, KG, 0>(tmem, aXC, KG, aW3, KG, bar_mma);
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = half * 32 + ch * 8 + 2 * j;
          w[j] = pack_h2(fmaxf(v[ch * 8 + 2 * j] + sB[80 + c] + rs.vb[c], 0.f), fmaxf(v[ch * 8 + 2 * j + 1] + sB[80 + c + 1] + rs.vb[c + 1], 0.f));
        }
        *reinterpret_cast<uint4*>(pX3 + cm_off(pt, half * 32 + ch * 8, 64)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    sync_for_mma();
    // ---- L4: 64 -> 64, ReLU
    if (tid == 0) issue_gemm<64, 64, 0>(tmem, aX3, 64, aW4, 64, bar_mma);
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = half * 32 + ch * 8 + 2 * j;
          w[j] = pack_h2(fmaxf(v[ch * 8 + 2 * j] + sB[144 + c], 0.f), fmaxf(v[ch * 8 + 2 * j + 1] + sB[144 + c + 1], 0.f));
        }
        *reinterpret_cast<uint4*>(pX4 + cm_off(pt, half * 32 + ch * 8, 64)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    sync_for_mma();
    // ---- L5: 64 -> 3 (padded 16)
    if (tid == 0) issue_gemm<16, 64, 0>(tmem, aX4, 64, aW5, 64, bar_mma);
    mma_wait();
    if (owner) {
      float v[8];
      TmemLd<8>::ld(trow, v);
#pragma unroll
      for (int c = 0; c < 3; ++c) sOut[pt * 4 + c] = __half2float(__float2half_rn(v[c] + sB[208 + c]));
    }
    // ============ 4. compositing — once per point (owner threads: warps 0-3). sumw / anyvalid were completed many barriers ago and every
    // owner reads back only the sOut row it wrote itself, so no block barrier here; the other four warps stage the NEXT tile's rays meanwhile.
    float out4[4] = {0.f, 0.f, 0.f, 0.f};
    float w = 0.f;
    if (owner) {
#pragma unroll
      for (int c = 0; c < 4; ++c) out4[c] = sOut[pt * 4 + c];
      w = valid ? w_raw / (rs.sumw + 1e-10f) : 0.f;
      float pr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) pr[c] = warp_sum(w * sigmoidf_(out4[c]));
      if (lane == 0 && (pr[0] != 0.f || pr[1] != 0.f || pr[2] != 0.f)) {
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(&sRay[rl].rgb[c], pr[c]);
      }
      named_bar_sync(1, PT);                                // (C) rgb_map complete — the four owner warps only
    } else {
      const int ng = s_next[cur];                           // written at the top of this tile, several block barriers ago
      if (pt < R && ng < a.n_groups) setup_ray(sRayNext[pt], a, ng * R + pt);
    }
    // ============ 5. loss seeds
    float dsdf_s = 0.f;
    if (owner) {
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
      dsdf_s = d_out[3] * scale_ls;
      const float s0 = d_out[0] * scale_ls, s1 = d_out[1] * scale_ls, s2 = d_out[2] * scale_ls;
      overflow |= !(fabsf(s0) <= 65504.f) || !(fabsf(s1) <= 65504.f) || !(fabsf(s2) <= 65504.f) || !(fabsf(dsdf_s) <= 65504.f);
      *reinterpret_cast<uint4*>(pDO + cm_off(pt, 0, 16)) = make_uint4(pack_h2(s0, s1), pack_h2(s2, 0.f), 0u, 0u);
      *reinterpret_cast<uint4*>(pDO + cm_off(pt, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
    sync_for_mma();                                         // (S0) dOut visible to the warps and to the tensor core

    // ============ 6. backward: per layer the dgrad MMA is issued first (async), the warps do the wgrad meanwhile
    // ---- layer 5
    if (tid == 0) issue_gemm<64, 16, 1>(tmem, aDO, 16, aW5, 64, bar_mma);
    if (warp < S5::ITEMS) wgrad_item<S5::CNT>(aDO, 16, aX4, 64, 0, (warp % S5::GROUPS) * S5::CNT, wg5, wb5, (warp % S5::GROUPS) == 0, lane);
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
      tc_fence_before();
      __syncthreads();                                      // every warp finished reading X4 (wgrad5) and its TMEM rows
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        unsigned char* p = pX4 + cm_off(pt, half * 32 + ch * 8, 64);
        const uint4 m = *reinterpret_cast<const uint4*>(p);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // packed: round the pair, mask it with (activation > 0), test the exponents
          const uint32_t keep = __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
          o[j] = pack_h2(v[ch * 8 + 2 * j], v[ch * 8 + 2 * j + 1]) & keep;
          overflow |= ((o[j] & 0x7C00u) == 0x7C00u) || ((o[j] & 0x7C000000u) == 0x7C000000u);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    sync_for_mma();                                         // dY4 visible
    // ---- layer 4
    if (tid == 0) issue_gemm<64, 64, 1>(tmem, aX4, 64, aW4, 64, bar_mma);
    if (warp < S4::ITEMS)
      wgrad_item<S4::CNT>(aX4, 64, aX3, 64, warp / S4::GROUPS, (warp % S4::GROUPS) * S4::CNT, wg4, wb4, (warp % S4::GROUPS) == 0, lane);
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
      tc_fence_before();
      __syncthreads();                                      // wgrad4 finished reading X3
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        unsigned char* p = pX3 + cm_off(pt, half * 32 + ch * 8, 64);
        const uint4 m = *reinterpret_cast<const uint4*>(p);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // packed: round the pair, mask it with (activation > 0), test the exponents
          const uint32_t keep = __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
          o[j] = pack_h2(v[ch * 8 + 2 * j], v[ch * 8 + 2 * j + 1]) & keep;
          overflow |= ((o[j] & 0x7C00u) == 0x7C00u) || ((o[j] & 0x7C000000u) == 0x7C000000u);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    sync_for_mma();                                         // dY3 visible
    // ---- layer 3: dgrad -> dgeo (the view part needs no per-sample dgrad), wgrad dY3^T XG + the per-ray column sums of dY3
    if (tid == 0) issue_gemm<KG, 64, 1>(tmem, aX3, 64, aW3, KG, bar_mma);
    if (warp < S3::ITEMS)
      wgrad_item<S3::CNT, true>(aX3, 64, aXC, KG, warp / S3::GROUPS, (warp % S3::GROUPS) * S3::CNT, wg3, wb3, (warp % S3::GROUPS) == 0, lane, sRay,
                                Sp / 16, w3v, V);
    mma_wait();
    {
      float v[8];
      TmemLd<8>::ld(trow + half * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = half * 8 + j;
        if (col < 15) {
          overflow |= !(fabsf(v[j]) <= 65504.f);
          *reinterpret_cast<__half*>(pDO + cm_off(pt, 1 + col, 16)) = __float2half_rn(v[j]);
        }
      }
      if (owner) *reinterpret_cast<__half*>(pDO + cm_off(pt, 0, 16)) = __float2half_rn(dsdf_s);
    }
    sync_for_mma();                                         // dH2 visible
    // ---- layer 2
    if (tid == 0) issue_gemm<64, 16, 1>(tmem, aDO, 16, aW2, 64, bar_mma);
    if (warp < S2::ITEMS) wgrad_item<S2::CNT>(aDO, 16, aX1, 64, 0, (warp % S2::GROUPS) * S2::CNT, wg2, wb2, (warp % S2::GROUPS) == 0, lane);
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
      tc_fence_before();
      __syncthreads();                                      // wgrad2 finished reading X1
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        unsigned char* p = pX1 + cm_off(pt, half * 32 + ch * 8, 64);
        const uint4 m = *reinterpret_cast<const uint4*>(p);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // packed: round the pair, mask it with (activation > 0), test the exponents
          const uint32_t keep = __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
          o[j] = pack_h2(v[ch * 8 + 2 * j], v[ch * 8 + 2 * j + 1]) & keep;
          overflow |= ((o[j] & 0x7C00u) == 0x7C00u) || ((o[j] & 0x7C000000u) == 0x7C000000u);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    sync_for_mma();                                         // dY1 visible
    // ---- layer 1: dgrad -> dEnc fp32 (X3 region, dead), wgrad dY1^T X0
    if (tid == 0) issue_gemm<KE, 64, 1>(tmem, aX1, 64, aW1, KE, bar_mma);
    if (warp < S1::ITEMS)
      wgrad_item<S1::CNT>(aX1, 64, aX0, KE, warp / S1::GROUPS, (warp % S1::GROUPS) * S1::CNT, wg1, wb1, (warp % S1::GROUPS) == 0, lane);
    mma_wait();
    {
      constexpr int NH = KE / 2;                            // columns per thread
      float v[NH];
      TmemLd<NH>::ld(trow + half * NH, v);
      float* dE = reinterpret_cast<float*>(pX3) + des_idx(pt);                 // transposed: [column][DES], spills into X4 (dead)
#pragma unroll
      for (int j = 0; j < NH; ++j) dE[(half * NH + j) * DES] = v[j];
    }
    tc_fence_before();
    __syncthreads();                                        // dEnc visible to the scatter's (8 samples x 1 level) threads
    // ============ 7. grid-gradient scatter + pose Jacobian. Thread = 8 CONSECUTIVE samples of one ray x ONE level (warp w: levels
    // 2w, 2w+1; lane & 15: sample group). Consecutive samples mostly fall into the same grid cell (C2: one new cell every 17 samples
    // at the coarsest level, every 2.4 at the finest), so the 8 corner contributions are summed in registers over the run and go out
    // as ONE set of reductions per run: ~3.5x fewer operations on the L2 atomic unit, which is what bounds the scatter
    // (profiles/red_bench.cu: ~180 G lane-ops/s whatever the operand width).
    {
      const int sg = lane & 15;                             // (pairing coarse with fine levels in a warp to even out the runs per warp: +2.5 us)
      const int l = 2 * warp + (lane >> 4);
      const int p0 = 8 * sg;
      const RayS& r8 = sRay[p0 / Sp];
      float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#ifdef NOF_EXP_NO_SCATTER
      if (false) {
#else
      if (l < L) {
#endif
        const float scale = lv.scale[l];
        const uint32_t off = lv.off[l];
        const float* dE0 = reinterpret_cast<const float*>(pX3) + (size_t)(2 * l) * DES;
        float acc[8][2];
        uint32_t cpg[3] = {0u, 0u, 0u};
        bool have = false;
#pragma unroll 1
        for (int j = 0; j <= 8; ++j) {
          bool live = false;
          uint32_t pg[3] = {0u, 0u, 0u};
          float fr[3] = {0.f, 0.f, 0.f}, g0 = 0.f, g1 = 0.f;
          if (j < 8) {
            const int q = p0 + j, qi = des_idx(q);
            if (sZ[DES + qi] != 0.f) {
              g0 = dE0[qi];
              g1 = dE0[DES + qi];
              const float zq = sZ[qi];
#pragma unroll
              for (int d = 0; d < 3; ++d) {                 // same u as the gather used: same cell, same weights
                const float pp = fmaf(sZ[(2 + d) * DES + qi], scale, 0.5f);
                const float fl = floorf(pp);
                pg[d] = (uint32_t)fl;
                fr[d] = pp - fl;
              }
              live = (g0 != 0.f || g1 != 0.f);
              if (a.p.need_pose_grad && live) {
                float gx[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  const float2 jj = __half22float2(Jslot[(size_t)(l * 3 + d) * PT + q]);
                  gx[d] = 0.5f * fmaf(g0, jj.x, g1 * jj.y);
                  st[d] += gx[d];
                  st[3 + d] = fmaf(gx[d], zq, st[3 + d]);
                }
              }
            }
          }
          const bool newcell = live && (!have || pg[0] != cpg[0] || pg[1] != cpg[1] || pg[2] != cpg[2]);
          if (have && (newcell || j == 8)) {                // the run ended: one set of reductions for all its samples
            uint32_t idx[8];
            corner_indices(lv, l, cpg, idx);
            float* base = a.p.grad_table + (size_t)off * 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t e0 = idx[2 * k], e1 = idx[2 * k + 1];
#ifdef NOF_EXP_NO_RED
              if (acc[2 * k][0] == 123456.f) red_add_v2(base + (size_t)e0 * 2, acc[2 * k][0], acc[2 * k][1]);
#else   // one 8-byte reduction per corner: pairing x/x+1 into 16-byte ones (scatter_level) costs more issue slots than it saves here
              red_add_v2(base + (size_t)e0 * 2, acc[2 * k][0], acc[2 * k][1]);
              red_add_v2(base + (size_t)e1 * 2, acc[2 * k + 1][0], acc[2 * k + 1][1]);
#endif
            }
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
      if (a.p.need_pose_grad) {
        // x = R (dir z) + t  =>  dL/dR[i][j] = dir[j] * sum gi z ,  dL/dt[i] = sum gi  (gi = 0.5 gx[i]: u = (x+1)/2). Sum over the lanes
        // of the same ray: Sp/8 sample groups (16, 8 or 4 lanes) in each half-warp, then across the two level halves.
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          float v = st[i];
          for (int o = 1; o < Sp / 8; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          st[i] = v;
        }
        if ((lane & (Sp / 8 - 1)) == 0 && lane < 16 && r8.active && r8.frame != 0) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float v = st[3 + i] * r8.dir[j];
              if (v != 0.f) red_add(a.p.grad_tf + (size_t)r8.frame * 12 + i * 4 + j, v);
            }
            if (st[i] != 0.f) red_add(a.p.grad_tf + (size_t)r8.frame * 12 + i * 4 + 3, st[i]);
          }
        }
      }
    }
    __syncthreads();
    // rays of this tile: d views = W3[:, views]^T c_r -> frame-feature gradient and (through the SH Jacobian) the pose. Warp w looks after
    // ray w: lane v computes column v (64 multiply-adds, no shuffles), lane 0 finishes.
    if (warp < R && sRay[warp].active && (a.p.grad_feat || (a.p.need_pose_grad && sRay[warp].frame != 0))) {
      RayT& r2 = sRay[warp];
      if (lane < V) {
        float acc = 0.f;
#pragma unroll 8
        for (int o = 0; o < 64; ++o) acc = fmaf(r2.cr[o], __half2float(sW3v[o * VPAD + lane]), acc);
        r2.dviews[lane] = acc;
      }
      __syncwarp();
      if (lane == 0) {
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
              const float vv = gd[i] * r2.u[j];
              if (vv != 0.f) red_add(a.p.grad_tf + (size_t)r2.frame * 12 + i * 4 + j, vv);
            }
        }
      }
    }
    // no barrier here: the next tile works on the other ray buffer, and everything else it overwrites (sZ, X0, the Jacobian slot) was last
    // read before the barrier above
    grp = s_next[cur];
    cur ^= 1;
  }

  // ============ flush
  {
    float* G = a.p.grad_mlp;
    const int K3 = V + 15;
    if (warp < S1::ITEMS)
      flush_item<S1::CNT>(G, a.po[0], a.po[1], E, 64, warp / S1::GROUPS, (warp % S1::GROUPS) * S1::CNT, wg1, wb1, (warp % S1::GROUPS) == 0, g8, t4);
    if (warp < S2::ITEMS)
      flush_item<S2::CNT>(G, a.po[2], a.po[3], 64, 16, 0, (warp % S2::GROUPS) * S2::CNT, wg2, wb2, (warp % S2::GROUPS) == 0, g8, t4);
    if (warp < S3::ITEMS)
      flush_item<S3::CNT>(G, a.po[4], a.po[5], 15, 64, warp / S3::GROUPS, (warp % S3::GROUPS) * S3::CNT, wg3, wb3, (warp % S3::GROUPS) == 0, g8, t4, K3, V);
    if (warp < S3::ITEMS && (warp % S3::GROUPS) == 0) {        // the strip owners hold dW3[:, views]
      const int strip = warp / S3::GROUPS;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int v = t4 + 4 * k;
        if (v < V) {
          if (w3v[k][0] != 0.f) red_add(G + a.po[4] + (strip * 16 + g8) * K3 + v, w3v[k][0]);
          if (w3v[k][1] != 0.f) red_add(G + a.po[4] + (strip * 16 + g8 + 8) * K3 + v, w3v[k][1]);
        }
      }
    }
    if (warp < S4::ITEMS)
      flush_item<S4::CNT>(G, a.po[6], a.po[7], 64, 64, warp / S4::GROUPS, (warp % S4::GROUPS) * S4::CNT, wg4, wb4, (warp % S4::GROUPS) == 0, g8, t4);
    if (warp < S5::ITEMS)
      flush_item<S5::CNT>(G, a.po[8], a.po[9], 64, 3, 0, (warp % S5::GROUPS) * S5::CNT, wg5, wb5, (warp % S5::GROUPS) == 0, g8, t4);
  }
  {
    loss_acc[0] = loss_acc[1] + loss_acc[2] + loss_acc[3] + loss_acc[4];
    float vals[7] = {loss_acc[0], loss_acc[1], loss_acc[2], loss_acc[3], loss_acc[4], n_valid_s, n_valid_r};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float v = warp_sum(vals[i]);
      if (lane == 0 && v != 0.f) red_add(a.p.losses + i, v);
    }
    const unsigned ov = __ballot_sync(0xffffffffu, overflow);
    if (lane == 0 && ov && a.p.found_inf) atomicExch(a.p.found_inf, 1);
    const unsigned bad = __ballot_sync(0xffffffffu, !tc_ok);
    if (lane == 0 && bad && a.p.found_inf) atomicExch(a.p.found_inf, 2);      // an mbarrier wait timed out: results are invalid
  }
  // ---- release TMEM
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace tc

size_t step_tc_smem(int KE) { return (size_t)tc::make_plan(KE).total; }

template <int KE>
static int launch_tc(const StepArgs& a, int blocks, cudaStream_t st) {
  const size_t smem = step_tc_smem(KE);
  // function attributes are per device: set on every launch (a cheap host-side call) rather than once per process
  cudaFuncSetAttribute(tc::step_tc_kernel<KE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const img::ImagePlan ip = img::make_image_plan(KE);
  if ((size_t)ip.bytes + 16 > kWPackBytes) { set_error("nof_step_fused(tc): operand image too large"); return NOF_E_INVALID; }
  img::pack_image_kernel<KE><<<(ip.bytes / 4 + 255) / 256, 256, 0, st>>>(a);
  tc::step_tc_kernel<KE><<<blocks, tc::NT, smem, st>>>(a);
  return check_launch("step_tc_kernel");
}

int step_tc_dispatch(const StepArgs& a, int blocks, cudaStream_t st) {
  if (a.KE == 32) return launch_tc<32>(a, blocks, st);
  if (a.KE == 16) return launch_tc<16>(a, blocks, st);
  set_error("nof_step_fused(amp, tcgen05): unsupported KE=%d", a.KE);
  return NOF_E_UNSUPPORTED;
}

}  // namespace nof
