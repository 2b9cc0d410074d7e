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
//     which serves as K-major A/B for the forward GEMMs, as MN-major B for the dgrad GEMMs (W is never transposed) AND as MN-major A and B
//     for the weight-gradient GEMMs (dY^T X: the point index is the K dimension of both operands);
//   * the weight gradients are tensor-core GEMMs too (M = 64, K = 128 points = 8 instructions per layer) whose accumulators STAY IN TMEM
//     for the whole kernel (232 of the CTA's 256 columns) and are flushed to global memory once, at the end. Each is issued right behind the
//     dgrad MMA of its layer by the same thread; the layer's epilogue waits for it only before it overwrites an operand. Bias gradients and
//     the per-ray column sums of dY3 come out of the same GEMMs: the three B operands carry 8 extra columns holding the ray indicator
//     (column r = 1 for the points of ray r). profiles/tc_issue.cu: one tcgen05.mma costs the issuing thread 46-49 cycles, whatever its shape.
// Ablation on B200 (profiles/README.md): with the weight gradients on mma.sync + ldmatrix (register accumulators, previous version) they
// cost 40 of the kernel's 204 us at C2.
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
#ifndef NOF_TC_GATHER_UNROLL
#define NOF_TC_GATHER_UNROLL 2                   // levels in flight per thread in the gather
#endif
#define NOF_PRAGMA(x) _Pragma(#x)
#define NOF_UNROLL(n) NOF_PRAGMA(unroll n)
constexpr int HB = 64 + 8;                        // row length of a 64-wide activation buffer that carries the ray indicator
constexpr int TM_COLS = 256;                     // TMEM columns per CTA: 64 activations + 232 weight-gradient accumulators, two CTAs per SM
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
  s.x0 = take(PT * (KE + 8) * 2);            // + 8 columns: ray indicator (the B operand of wgrad 1 ends with it)
  s.x1 = take(PT * 64 * 2);
  s.xc = take(PT * (KG + 8) * 2);            // geo features (15 + pad): the only per-sample input of the colour net; + indicator
  s.x3 = take(PT * HB * 2);                  // + indicator; also holds dEnc fp32 [KE][DES] (transposed) at the end of the backward
  s.x4 = take(PT * 64 * 2);
  s.d_o = take(PT * 16 * 2);
  s.out = take(PT * 4 * 4);
  s.zs = take(5 * DES * 4);                  // z, valid flag and u[3] of every point, in the scatter's padded order
  s.rays = take(2 * MAX_R * (int)sizeof(RayT));    // two tiles' rays: the next tile's are staged while this one computes
  s.bar = take(64);                          // [0] TMA staging barrier, [1] forward/dgrad MMA completion, [2] wgrad MMA completion
  s.tmem = take(16);
#ifdef NOF_EXP_STAGE_L0   // ablation: level 0 of the fp16 table (17^3 entries) staged in shared memory by TMA
  s.l0 = take(4920 * 4);
#else
  s.l0 = o;
#endif
  s.total = o;
  return s;
}

// One weight-gradient GEMM on the tensor core: D[64 x N] (+)= A^T B over the tile's 128 points. Both operands are [128 points x cols]
// core-matrix buffers read MN-major (M / N run along the columns, K = the point index along the rows): A's first 64 columns, B's first N.
// RA / RB: row lengths of the two buffers. The accumulator (M = 64: rows 16q .. 16q+15 live in lanes 0..15 of TMEM lane quadrant q) is
// never cleared between tiles: `fresh` only on the CTA's first tile.
template <int N>
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t a_addr, int RA, uint32_t b_addr, int RB, bool fresh) {
  constexpr uint32_t idesc = umma_idesc(64, N, 1, 1);
#ifdef NOF_EXP_NO_WGRAD
  return;
#endif
#pragma unroll
  for (int ks = 0; ks < PT / 16; ++ks) {
    const uint64_t ad = umma_desc(a_addr + ks * 2 * (RA * 16), RA * 16, 128);              // MN-major: LBO = 8-row group stride, SBO = 8-column group stride
    const uint64_t bd = umma_desc(b_addr + ks * 2 * (RB * 16), RB * 16, 128);
    umma_f16(tmem_d, ad, bd, idesc, (ks > 0 || !fresh) ? 1u : 0u);
  }
}

// End of the kernel: one [64 x NC] accumulator block out of TMEM into the global gradient. Thread (quadrant q, lane < 16) owns row
// 16q + lane; warps q and q + 4 split the columns. addr(row, col) returns the destination (nullptr: padding, skip).
template <int NC, class F>
__device__ __forceinline__ void flush_tmem(uint32_t tq, int col0, int warp, int lane, F addr) {
  static_assert(NC % 8 == 0, "flush_tmem");
  const int row = (warp & 3) * 16 + lane;
#pragma unroll 1
  for (int c = (warp >> 2) * 8; c < NC; c += 16) {
    float v[8];
    TmemLd<8>::ld(tq + col0 + c, v);                       // warp-collective: lanes >= 16 read unused TMEM lanes
    if (lane < 16) {
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        float* d0 = addr(row, c + j);
        float* d3 = addr(row, c + j + 3);
        if (d0 && d3 == d0 + 3 && (reinterpret_cast<uintptr_t>(d0) & 15u) == 0u) {          // four columns of a row-major weight: one 16-byte reduction
          if (v[j] != 0.f || v[j + 1] != 0.f || v[j + 2] != 0.f || v[j + 3] != 0.f) red_add_v4(d0, v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float* dst = addr(row, c + j + k);
            if (dst && v[j + k] != 0.f) red_add(dst, v[j + k]);
          }
        }
      }
    }
  }
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
  uint64_t* bar_wg = bar_tma + 2;
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
    mbar_init(bar_wg, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)TM_COLS) : "memory");
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

  // TMEM columns: [0,64) activations / dgrad results; then the weight-gradient accumulators, kept for the whole kernel
  //   D4 [64 out x (64 in | 8 ind)]   D1 [64 out x (KE in | 8 ind)]   D3 [64 out x (16 geo | 8 ind)]   D2^T [64 in x 16 out]   D5^T [64 in x 16 out]
  constexpr int KEB = KE + 8, KGB = KG + 8;
  constexpr int C4 = 64, C1 = C4 + HB, C3 = C1 + KEB, C2 = C3 + KGB, C5 = C2 + 16;
  static_assert(C5 + 16 <= TM_COLS, "TMEM columns");
  float wb2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // this thread's share of d b2 (its 8 columns of dH2, its point of every tile)
  float wb2s = 0.f;                                              // owner threads: d b2[0] (the sdf column)
  float wb5[3] = {0.f, 0.f, 0.f};                                // owner threads: d b5
  float w3v[MAX_V];                                              // lanes < 16 of warps 0-3: row 16 q + lane of dW3[:, views]
  float cr_prev[MAX_R];                                          // the same threads: cumulative per-ray column sums of dY3 up to the previous tile
#pragma unroll
  for (int v = 0; v < MAX_V; ++v) w3v[v] = 0.f;
#pragma unroll
  for (int r = 0; r < MAX_R; ++r) cr_prev[r] = 0.f;
  float loss_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float n_valid_s = 0.f, n_valid_r = 0.f;
  bool overflow = false;
  uint32_t phase = 0, phase_wg = 0;          // parities of bar_mma / bar_wg
  bool fresh = true;                         // this CTA's first tile: the weight-gradient accumulators in TMEM start from zero

  const int Sp = a.Sp, R = a.R, S = a.p.S;
  const int half = tid / PT;                 // gather/scatter: which half of the levels; epilogues: which half of the columns
  const int pt = tid - half * PT;            // = 32*(warp%4) + lane: the TMEM lane this thread may read
  const int rl = pt / Sp, sidx = pt - rl * Sp;
  const int l_beg = half ? (L + 1) >> 1 : 0, l_end = half ? L : (L + 1) >> 1;
  const bool owner = half == 0;
  uint4* Jslot = reinterpret_cast<uint4*>(a.jws) + (size_t)blockIdx.x * MAX_L * PT;      // [level][point]: d enc / d u as 3 x half2 (+ pad) = one 16-byte access
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
  auto wg_commit = [&]() {
#ifndef NOF_EXP_NO_TC
    umma_commit(bar_wg);
#endif
  };
  // wait for the layer's weight-gradient MMAs: their operands may be overwritten, their accumulators read
  auto wg_wait = [&]() {
#ifdef NOF_EXP_NO_TC
    return;
#endif
    tc_ok &= mbar_wait_inl(bar_wg, phase_wg);
    phase_wg ^= 1u;
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
  // ray indicator (column r of the 8 extra B columns = 1 for the points of the tile's ray r): constant for the launch. X0's and XC's copies
  // are written once; X3's is rewritten every tile (the region doubles as the dEnc buffer).
  const uint4 ind_row = make_uint4((pt / Sp) == 0 ? 0x3C00u : ((pt / Sp) == 1 ? 0x3C000000u : 0u), (pt / Sp) == 2 ? 0x3C00u : ((pt / Sp) == 3 ? 0x3C000000u : 0u), 0u, 0u);
  if (owner) {
    *reinterpret_cast<uint4*>(pX0 + cm_off(pt, KE, KEB)) = ind_row;
    *reinterpret_cast<uint4*>(pXC + cm_off(pt, KG, KGB)) = ind_row;
  }
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
      NOF_UNROLL(NOF_TC_GATHER_UNROLL)
      for (int l = l_beg; l < l_end; ++l) {
        float enc[2], J[3][2];
        if (a.p.need_pose_grad) {
#ifdef NOF_EXP_STAGE_L0
          gather_level<true, true>(l == 0 ? (const void*)table0 : a.p.table_f16, lv, l, u, enc, J);
#else
          gather_level<true, true>(a.p.table_f16, lv, l, u, enc, J);
#endif
          Jslot[(size_t)l * PT + pt] = make_uint4(pack_h2(J[0][0], J[0][1]), pack_h2(J[1][0], J[1][1]), pack_h2(J[2][0], J[2][1]), 0u);
        } else {
          gather_level<true, false>(a.p.table_f16, lv, l, u, enc, J);
        }
        *reinterpret_cast<uint32_t*>(pX0 + cm_off(pt, 2 * l, KEB)) = pack_h2(enc[0], enc[1]);
      }
    } else {
      for (int l = l_beg; l < l_end; ++l) *reinterpret_cast<uint32_t*>(pX0 + cm_off(pt, 2 * l, KEB)) = 0u;
    }
    if (owner) for (int j = E; j < KE; j += 2) *reinterpret_cast<uint32_t*>(pX0 + cm_off(pt, j, KEB)) = 0u;
    sync_for_mma();
    // ============ 3. MLP forward: five tcgen05 GEMMs, epilogue = this thread's row, its half of the columns
    // ---- L1: E -> 64, ReLU
    if (tid == 0) issue_gemm<64, KE, 0>(tmem, aX0, KEB, aW1, KE, bar_mma);
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
        else *reinterpret_cast<__half*>(pXC + cm_off(pt, col - 1, KGB)) = hv;
      }
      if (!owner) *reinterpret_cast<__half*>(pXC + cm_off(pt, 15, KGB)) = __float2half_rn(0.f);
    }
    sync_for_mma();
    // ---- L3: geo 15 (+ the ray's view / feature share as a bias) -> 64, ReLU
    if (tid == 0) issue_gemm<64, KG, 0>(tmem, aXC, KGB, aW3, KG, bar_mma);
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
        *reinterpret_cast<uint4*>(pX3 + cm_off(pt, half * 32 + ch * 8, HB)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      if (owner) *reinterpret_cast<uint4*>(pX3 + cm_off(pt, 64, HB)) = ind_row;
    }
    sync_for_mma();
    // ---- L4: 64 -> 64, ReLU
    if (tid == 0) issue_gemm<64, 64, 0>(tmem, aX3, HB, aW4, 64, bar_mma);
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
      wb5[0] += __half2float(__float2half_rn(s0)); wb5[1] += __half2float(__float2half_rn(s1)); wb5[2] += __half2float(__float2half_rn(s2));
      *reinterpret_cast<uint4*>(pDO + cm_off(pt, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
    sync_for_mma();                                         // (S0) dOut visible to the warps and to the tensor core

    // ============ 6. backward: per layer the dgrad MMA, then the weight-gradient MMAs (8 k-steps over the tile's points) behind it in the same
    // in-order pipe; the epilogue needs the first for its values and the second only before it overwrites an operand
    // ---- layer 5
    if (tid == 0) {
      issue_gemm<64, 16, 1>(tmem, aDO, 16, aW5, 64, bar_mma);
      issue_wgrad<16>(tmem + C5, aX4, 64, aDO, 16, fresh);  // D5^T [64 in x 16 out] += X4^T dOut
      wg_commit();
    }
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
      uint32_t o[16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint4 m = *reinterpret_cast<const uint4*>(pX4 + cm_off(pt, half * 32 + ch * 8, 64));
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // packed: round the pair, mask it with (activation > 0), test the exponents
          const uint32_t keep = __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
          o[ch * 4 + j] = pack_h2(v[ch * 8 + 2 * j], v[ch * 8 + 2 * j + 1]) & keep;
          overflow |= ((o[ch * 4 + j] & 0x7C00u) == 0x7C00u) || ((o[ch * 4 + j] & 0x7C000000u) == 0x7C000000u);
        }
      }
      wg_wait();                                            // wgrad 5 has read X4: only now may the operand be overwritten
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(pX4 + cm_off(pt, half * 32 + ch * 8, 64)) = make_uint4(o[ch * 4], o[ch * 4 + 1], o[ch * 4 + 2], o[ch * 4 + 3]);
    }
    sync_for_mma();                                         // dY4 visible
    // ---- layer 4
    if (tid == 0) {
      issue_gemm<64, 64, 1>(tmem, aX4, 64, aW4, 64, bar_mma);
      issue_wgrad<HB>(tmem + C4, aX4, 64, aX3, HB, fresh);   // D4 [64 out x (64 in | ind)] += dY4^T [X3 | ind]
      wg_commit();
    }
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
      uint32_t o[16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint4 m = *reinterpret_cast<const uint4*>(pX3 + cm_off(pt, half * 32 + ch * 8, HB));
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // packed: round the pair, mask it with (activation > 0), test the exponents
          const uint32_t keep = __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
          o[ch * 4 + j] = pack_h2(v[ch * 8 + 2 * j], v[ch * 8 + 2 * j + 1]) & keep;
          overflow |= ((o[ch * 4 + j] & 0x7C00u) == 0x7C00u) || ((o[ch * 4 + j] & 0x7C000000u) == 0x7C000000u);
        }
      }
      wg_wait();                                            // wgrad 4 has read X3: only now may the operand be overwritten
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(pX3 + cm_off(pt, half * 32 + ch * 8, HB)) = make_uint4(o[ch * 4], o[ch * 4 + 1], o[ch * 4 + 2], o[ch * 4 + 3]);
    }
    sync_for_mma();                                         // dY3 visible
    // ---- layer 3: dgrad -> dgeo (the view part needs no per-sample dgrad), wgrad dY3^T XG + the per-ray column sums of dY3
    if (tid == 0) {
      issue_gemm<KG, 64, 1>(tmem, aX3, HB, aW3, KG, bar_mma);
      issue_wgrad<KGB>(tmem + C3, aX3, HB, aXC, KGB, fresh);  // D3 [64 out x (16 geo | ind)] += dY3^T [XC | ind]
      wg_commit();
    }
    mma_wait();
    {
      float v[8];
      TmemLd<8>::ld(trow + half * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = half * 8 + j;
        if (col < 15) {
          overflow |= !(fabsf(v[j]) <= 65504.f);
          const __half hv = __float2half_rn(v[j]);
          *reinterpret_cast<__half*>(pDO + cm_off(pt, 1 + col, 16)) = hv;
          wb2[j] += __half2float(hv);                       // d b2[1 + col]: what the tensor core would sum
        }
      }
      if (owner) {
        const __half hs = __float2half_rn(dsdf_s);
        *reinterpret_cast<__half*>(pDO + cm_off(pt, 0, 16)) = hs;
        wb2s += __half2float(hs);
      }
    }
    wg_wait();                                              // wgrad 3: its indicator columns are the per-ray column sums of dY3
    if (warp < 4) {
      // c_r = sum over the ray's samples of dY3: the accumulator is cumulative over this CTA's tiles, so this tile's share is the difference
      // to the previous read. Feeds d views (end of the tile) and dW3[:, views] += c_r (x) views.
      float c8[8];
      TmemLd<8>::ld(trow + C3 + KG, c8);
      if (lane < 16) {
        const int o = warp * 16 + lane;
#pragma unroll
        for (int r = 0; r < MAX_R; ++r) {
          if (r < R) {
            const float c = c8[r] - cr_prev[r];
            cr_prev[r] = c8[r];
            sRay[r].cr[o] = c;
#pragma unroll
            for (int vv = 0; vv < MAX_V; ++vv)
              if (vv < V) w3v[vv] = fmaf(c, __half2float(__float2half_rn(sRay[r].views[vv])), w3v[vv]);
          }
        }
      }
    }
    sync_for_mma();                                         // dH2 visible
    // ---- layer 2
    if (tid == 0) {
      issue_gemm<64, 16, 1>(tmem, aDO, 16, aW2, 64, bar_mma);
      issue_wgrad<16>(tmem + C2, aX1, 64, aDO, 16, fresh);  // D2^T [64 in x 16 out] += X1^T dH2
      wg_commit();
    }
    mma_wait();
    {
      float v[32];
      TmemLd<32>::ld(trow + half * 32, v);
      uint32_t o[16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint4 m = *reinterpret_cast<const uint4*>(pX1 + cm_off(pt, half * 32 + ch * 8, 64));
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // packed: round the pair, mask it with (activation > 0), test the exponents
          const uint32_t keep = __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
          o[ch * 4 + j] = pack_h2(v[ch * 8 + 2 * j], v[ch * 8 + 2 * j + 1]) & keep;
          overflow |= ((o[ch * 4 + j] & 0x7C00u) == 0x7C00u) || ((o[ch * 4 + j] & 0x7C000000u) == 0x7C000000u);
        }
      }
      wg_wait();                                            // wgrad 2 has read X1: only now may the operand be overwritten
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(pX1 + cm_off(pt, half * 32 + ch * 8, 64)) = make_uint4(o[ch * 4], o[ch * 4 + 1], o[ch * 4 + 2], o[ch * 4 + 3]);
    }
    sync_for_mma();                                         // dY1 visible
    // ---- layer 1: dgrad -> dEnc fp32 (X3 region, dead), wgrad dY1^T X0
    if (tid == 0) {
      issue_gemm<KE, 64, 1>(tmem, aX1, 64, aW1, KE, bar_mma);
      issue_wgrad<KEB>(tmem + C1, aX1, 64, aX0, KEB, fresh);  // D1 [64 out x (KE in | ind)] += dY1^T [X0 | ind]
      wg_commit();
    }
    mma_wait();
    {
      constexpr int NH = KE / 2;                            // columns per thread
      float v[NH];
      TmemLd<NH>::ld(trow + half * NH, v);
      float* dE = reinterpret_cast<float*>(pX3) + des_idx(pt);                 // transposed: [column][DES] — exactly the X3 region (dY3 is dead)
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
                const uint4 jv = Jslot[(size_t)l * PT + q];
                const uint32_t jw[3] = {jv.x, jv.y, jv.z};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  const float2 jj = __half22float2(*reinterpret_cast<const __half2*>(&jw[d]));
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
    wg_wait();                                              // wgrad 1 (long finished): X0 / X1 may be rewritten by the next tile
    fresh = false;
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

  // ============ flush: the weight-gradient accumulators leave TMEM once per CTA (9.1 k reductions, as the register version did)
  if (!fresh) {
    float* G = a.p.grad_mlp;
    const int K3 = V + 15;
    const int o = (warp & 3) * 16 + lane;                   // lanes < 16: the accumulator row this thread reads
    tc_fence_after();
    flush_tmem<64>(trow, C4, warp, lane, [&](int r, int c) { return G + a.po[6] + r * 64 + c; });
    flush_tmem<KE>(trow, C1, warp, lane, [&](int r, int c) { return c < E ? G + a.po[0] + r * E + c : (float*)nullptr; });
    flush_tmem<16>(trow, C3, warp, lane, [&](int r, int c) { return c < 15 ? G + a.po[4] + r * K3 + V + c : (float*)nullptr; });
    flush_tmem<16>(trow, C2, warp, lane, [&](int r, int c) { return G + a.po[2] + c * 64 + r; });                          // transposed: row = input
    flush_tmem<16>(trow, C5, warp, lane, [&](int r, int c) { return c < 3 ? G + a.po[8] + c * 64 + r : (float*)nullptr; });
    if (warp < 4) {                                         // bias gradients = the indicator columns summed over the rays
      float c4[8], c1[8];
      TmemLd<8>::ld(trow + C4 + 64, c4);
      TmemLd<8>::ld(trow + C1 + KE, c1);
      if (lane < 16) {
        float b4 = 0.f, b1 = 0.f, b3 = 0.f;
#pragma unroll
        for (int r = 0; r < MAX_R; ++r)
          if (r < R) { b4 += c4[r]; b1 += c1[r]; b3 += cr_prev[r]; }
        if (b4 != 0.f) red_add(G + a.po[7] + o, b4);
        if (b1 != 0.f) red_add(G + a.po[1] + o, b1);
        if (b3 != 0.f) red_add(G + a.po[5] + o, b3);
#pragma unroll
        for (int vv = 0; vv < MAX_V; ++vv)
          if (vv < V && w3v[vv] != 0.f) red_add(G + a.po[4] + o * K3 + vv, w3v[vv]);
      }
    }
    {                                                       // d b2, d b5: per-thread sums over this CTA's tiles -> warp -> global
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = warp_sum(wb2[j]);
        if (lane == 0 && half * 8 + j < 15 && t != 0.f) red_add(G + a.po[3] + 1 + half * 8 + j, t);
      }
      const float ts = warp_sum(wb2s);
      if (lane == 0 && ts != 0.f) red_add(G + a.po[3], ts);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float t = warp_sum(wb5[c]);
        if (lane == 0 && t != 0.f) red_add(G + a.po[9] + c, t);
      }
    }
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TM_COLS) : "memory");
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
