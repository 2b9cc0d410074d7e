// Fused forward + loss + backward of one Neural-Object-Field train step, AMP policy (cfg amp: true):
// fp16 hash table gathered with 4-byte vector loads, fp16 activations resident in shared memory, the five tiny
// dense layers of NeRFSmall (nerf_helpers.py:243-321) on tensor cores (mma.sync m16n8k16, fp32 accumulate) for
// forward, dgrad and wgrad, weight gradients accumulated in registers across all tiles of a persistent CTA, fp32
// vector reductions (red.global.add.v2.f32) for the grid gradient. Weights and biases are staged into shared
// memory once per CTA with a TMA bulk copy (cp.async.bulk + mbarrier).
//
// Tile = PT points = whole rays; the CTA runs 2*PT threads (two threads per point):
//   * gather / scatter: the two threads of a point split the L levels (half each) -> twice the loads in flight per point
//     and half the dependent round trips per thread;
//   * MLP: warp w owns rows [16w, 16w+16) (one m16 tile) for forward and dgrad, all warps cooperate on wgrad;
//   * 16 resident warps per SM (2 CTAs x 8 warps at PT=128) instead of 8: the kernel is latency-bound (ncu: issue
//     utilisation 22 %, long-scoreboard stalls on the gather), see profiles/README.md.
//
// Replaces, for one batch: run_network + raw2outputs + loss assembly + loss.backward() of the reference
// (nerf_runner.py:1083-1088,1227-1304,1132-1169,679-758; grid.py:34-99; gridencoder.cu:107-365).
#include "nof_step_common.cuh"

#ifndef NOF_GATHER_UNROLL
#define NOF_GATHER_UNROLL 2      // levels of the multires gather in flight per thread (8 x 4-byte loads each)
#endif

namespace nof {

constexpr int kGatherUnroll = NOF_GATHER_UNROLL;

// ------------------------------------------------------------------------------------------------
// tensor-core / ldmatrix / TMA primitives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t r[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t r[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// shared-memory plan (halfs unless noted). Row strides are K+8 halfs: ldmatrix rows stay 16-byte aligned and the 8
// rows of an 8x8 tile fall into distinct 16-byte bank groups.
// ------------------------------------------------------------------------------------------------
constexpr int LD64 = 72;   // stride of 64-wide activations / weights
constexpr int LD32 = 40;   // stride of 32-wide
constexpr int LD16 = 24;   // stride of 16-wide

struct SmemPlan {
  int w1, w2, w3, w4, w5, bias, x0, x1, xc, x3, x4, d_o, out, rays, lv, bar, qh, nbuf, total;   // byte offsets
  int ldx0;                // stride (halfs) of X0 / W1 (KE + 8)
};

__host__ __device__ inline SmemPlan make_plan(int PT, int KE, bool eik = false) {
  SmemPlan s;
  s.ldx0 = KE + 8;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) / 128 * 128; return r; };
  s.w1 = take(64 * s.ldx0 * 2);
  s.w2 = take(16 * LD64 * 2);
  s.w3 = take(64 * LD32 * 2);
  s.w4 = take(64 * LD64 * 2);
  s.w5 = take(16 * LD64 * 2);
  s.bias = take(216 * 4);                    // b1 64, b2 16, b3 64, b4 64, b5 8
  s.x0 = take(PT * s.ldx0 * 2);
  s.x1 = take(PT * LD64 * 2);
  s.xc = take(PT * LD32 * 2);
  s.x3 = take(PT * LD64 * 2);
  s.x4 = take(PT * LD64 * 2);
  s.d_o = take(PT * LD16 * 2);               // dOut (phase 5), later dH2 (phase 2)
  s.out = take(PT * 4 * 4);                  // fp32 [PT][4]
  s.rays = take(MAX_R * (int)sizeof(RayS));
  s.lv = take((int)sizeof(LevelS));
  s.bar = take(64);
  s.qh = take(eik ? PT * s.ldx0 * 2 : 0);      // eikonal: q = d sdf / d enc of every point (fp16), kept until the scatter
  s.nbuf = take(eik ? 2 * PT * 3 * 4 : 0);     // eikonal: the two half-sums of the normal
  s.total = o;
  return s;
}

// ------------------------------------------------------------------------------------------------
// warp-level GEMM pieces on the warp's own 16 rows (one m16 tile)
// ------------------------------------------------------------------------------------------------
// Y[16 x N] = A[16 x K] * W^T, W stored [N][K] (nn.Linear layout).  acc[nt][4]
template <int K, int N>
__device__ __forceinline__ void warp_fwd(const __half* A, int lda, const __half* W, int ldw, float (*acc)[4], int lane) {
#ifdef NOF_EXP_NO_MMA      // profiling experiment only: the kernel without its tensor-core work
  return;
#endif
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, A + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * lda + ks * 16 + (lane >> 4) * 8);
#pragma unroll
    for (int np = 0; np < N / 16; ++np) {      // two n-tiles per ldmatrix.x4
      uint32_t b[4];
      ldsm_x4(b, W + (size_t)(np * 16 + (lane & 7) + (lane >> 4) * 8) * ldw + ks * 16 + ((lane >> 3) & 1) * 8);
      mma16816(acc[np * 2], a, b[0], b[1]);
      mma16816(acc[np * 2 + 1], a, b[2], b[3]);
    }
    if constexpr ((N / 8) % 2 == 1) {          // N == 8: single n-tile (the x4 load reads 16 W rows; rows 8..15 exist, unused)
      uint32_t b[4];
      ldsm_x4(b, W + (size_t)((N / 16) * 16 + (lane & 7) + (lane >> 4) * 8) * ldw + ks * 16 + ((lane >> 3) & 1) * 8);
      mma16816(acc[N / 8 - 1], a, b[0], b[1]);
    }
  }
}

// dX[16 x NI] = dY[16 x KO] * W, W stored [KO][NI] row-major (k = output index of the layer).
template <int KO, int NI>
__device__ __forceinline__ void warp_dgrad(const __half* dY, int ldy, const __half* W, int ldw, float (*acc)[4], int lane) {
#ifdef NOF_EXP_NO_MMA
  return;
#endif
#pragma unroll
  for (int ks = 0; ks < KO / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, dY + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * ldy + ks * 16 + (lane >> 4) * 8);
#pragma unroll
    for (int np = 0; np < NI / 16; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, W + (size_t)(ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldw + np * 16 + (lane >> 4) * 8);
      mma16816(acc[np * 2], a, b[0], b[1]);
      mma16816(acc[np * 2 + 1], a, b[2], b[3]);
    }
  }
}

// One wgrad item: dW[strip*16 .. +16][nt0*8 .. +NTU*8] += dY^T X over all PT rows of the CTA; bias via a ones B-operand.
// (For odd NTU the x4 load of the last pair also reads the 8 columns after the tile: they are inside the row padding.)
template <int NTU>
__device__ __forceinline__ void wgrad_item(const __half* dY, int ldy, const __half* X, int ldx, int PT, int strip, int nt0,
                                           float (*acc)[4], float* bias2, bool do_bias, int lane) {
#ifdef NOF_EXP_NO_MMA
  return;
#endif
  const uint32_t ones = 0x3C003C00u;
  for (int ks = 0; ks < PT / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4_t(a, dY + (size_t)(ks * 16 + (lane & 7) + (lane >> 4) * 8) * ldy + strip * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int np = 0; np < (NTU + 1) / 2; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, X + (size_t)(ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldx + (nt0 + np * 2) * 8 + (lane >> 4) * 8);
      mma16816(acc[np * 2], a, b[0], b[1]);
      if (np * 2 + 1 < NTU) mma16816(acc[np * 2 + 1], a, b[2], b[3]);
    }
    if (do_bias) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(t, a, ones, ones);
      bias2[0] += t[0];
      bias2[1] += t[2];
    }
  }
}

// Even split of one layer's weight gradient (NS strips of 16 output rows x NTL n-tiles of 8 input columns) over the warps of
// the CTA: every warp owns at most ONE item (strip, CNT consecutive n-tiles) per layer, so each wgrad phase keeps all warps
// busy for the same time (the first version gave whole layers to 2-4 warps and the rest waited at the barrier: ncu showed
// 21 % barrier stalls). CNT = smallest divisor of NTL that is >= NS*NTL/NWARP; accumulators stay in registers for the
// whole kernel: (CNT1 + 1 + 2 + 4 + 1) * 4 floats.
template <int NS, int NTL, int NWARP>
struct WSplit {
  static constexpr int ideal = (NS * NTL + NWARP - 1) / NWARP;
  static constexpr int CNT = ideal <= 1 ? 1 : (ideal <= 2 ? (NTL % 2 == 0 ? 2 : NTL) : (NTL % 4 == 0 ? 4 : NTL));
  static constexpr int GROUPS = NTL / CNT;
  static constexpr int ITEMS = NS * GROUPS;
  static_assert(CNT <= 4 && NTL % CNT == 0, "wgrad split");
};

// registers -> global (fp32 atomics) for one item
template <int CNT>
__device__ __forceinline__ void flush_item(float* G, int wofs, int bofs, int ncols, int nrows, int strip, int nt0, const float (*acc)[4],
                                           const float* bias2, bool has_bias, int g8, int t4) {
#pragma unroll
  for (int nt = 0; nt < CNT; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int o = strip * 16 + g8 + h * 8, i = (nt0 + nt) * 8 + 2 * t4 + c;
        const float v = acc[nt][h * 2 + c];
        if (o < nrows && i < ncols && v != 0.f) red_add(G + wofs + (size_t)o * ncols + i, v);
      }
  if (has_bias && t4 == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = strip * 16 + g8 + h * 8;
      if (o < nrows && bias2[h] != 0.f) red_add(G + bofs + o, bias2[h]);
    }
  }
}

// epilogue helpers for one m16 tile ----------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void init_bias(float (*acc)[4], const float* bias, int t4) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    acc[nt][0] = bias[nt * 8 + 2 * t4]; acc[nt][1] = bias[nt * 8 + 2 * t4 + 1];
    acc[nt][2] = acc[nt][0];            acc[nt][3] = acc[nt][1];
  }
}
template <int NT>
__device__ __forceinline__ void zero_acc(float (*acc)[4]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
}
// ReLU + fp16 store of a 16 x 64 tile
__device__ __forceinline__ void store_relu64(__half* Y, int row0, const float (*acc)[4], int g8, int t4) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    __half* y = Y + (size_t)(row0 + g8) * LD64 + nt * 8 + 2 * t4;
    *reinterpret_cast<uint32_t*>(y) = pack_h2(fmaxf(acc[nt][0], 0.f), fmaxf(acc[nt][1], 0.f));
    *reinterpret_cast<uint32_t*>(y + 8 * LD64) = pack_h2(fmaxf(acc[nt][2], 0.f), fmaxf(acc[nt][3], 0.f));
  }
}
// dY = dX * relu'(X) written in place over X (16 x 64 tile); returns true if an fp16 conversion overflowed
__device__ __forceinline__ bool store_masked64(__half* X, int row0, const float (*acc)[4], int g8, int t4) {
  bool ovf = false;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    __half* y = X + (size_t)(row0 + g8) * LD64 + nt * 8 + 2 * t4;
    const __half2 m0 = *reinterpret_cast<__half2*>(y), m1 = *reinterpret_cast<__half2*>(y + 8 * LD64);
    const float v0 = __low2float(m0) > 0.f ? acc[nt][0] : 0.f, v1 = __high2float(m0) > 0.f ? acc[nt][1] : 0.f;
    const float v2 = __low2float(m1) > 0.f ? acc[nt][2] : 0.f, v3 = __high2float(m1) > 0.f ? acc[nt][3] : 0.f;
    ovf |= !(fabsf(v0) <= 65504.f) || !(fabsf(v1) <= 65504.f) || !(fabsf(v2) <= 65504.f) || !(fabsf(v3) <= 65504.f);
    *reinterpret_cast<uint32_t*>(y) = pack_h2(v0, v1);
    *reinterpret_cast<uint32_t*>(y + 8 * LD64) = pack_h2(v2, v3);
  }
  return ovf;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int PT, int KE_, bool EIK>
__global__ void __launch_bounds__(2 * PT, (PT <= 128 && !EIK) ? 2 : 1) step_amp_kernel(const StepArgs a) {
  constexpr int NT = 2 * PT;                 // threads
  constexpr int NWARP = NT / 32;
  constexpr int KE = KE_;
  constexpr int LDX0 = KE + 8;
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemPlan sp = make_plan(PT, KE, EIK);
  __half* sW1 = reinterpret_cast<__half*>(smem + sp.w1);
  __half* sW2 = reinterpret_cast<__half*>(smem + sp.w2);
  __half* sW3 = reinterpret_cast<__half*>(smem + sp.w3);
  __half* sW4 = reinterpret_cast<__half*>(smem + sp.w4);
  __half* sW5 = reinterpret_cast<__half*>(smem + sp.w5);
  float* sB = reinterpret_cast<float*>(smem + sp.bias);
  __half* X0 = reinterpret_cast<__half*>(smem + sp.x0);
  __half* X1 = reinterpret_cast<__half*>(smem + sp.x1);
  __half* XC = reinterpret_cast<__half*>(smem + sp.xc);
  __half* X3 = reinterpret_cast<__half*>(smem + sp.x3);
  __half* X4 = reinterpret_cast<__half*>(smem + sp.x4);
  __half* DO = reinterpret_cast<__half*>(smem + sp.d_o);
  float* sOut = reinterpret_cast<float*>(smem + sp.out);
  RayS* sRay = reinterpret_cast<RayS*>(smem + sp.rays);
  LevelS& lv = *reinterpret_cast<LevelS*>(smem + sp.lv);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sp.bar);
  float* sStage = reinterpret_cast<float*>(smem + sp.x1);     // fp32 staging of the packed params (aliases X1..)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = a.E, V = a.V, L = a.p.L;
  const float scale_ls = a.p.loss_scale ? *a.p.loss_scale : 1.0f;

  // ---- stage parameters: TMA bulk copy of the packed fp32 block into smem, then convert to padded fp16 tiles
  const int n_par = a.po[9] + 3;
  const uint32_t par_bytes = (uint32_t)((n_par * 4 + 15) / 16 * 16);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, par_bytes);
    tma_bulk_g2s(sStage, a.p.mlp, par_bytes, bar);
  }
  init_levels(lv, a);
  // zero the weight tiles (padding rows / columns must be exact zeros)
  for (int i = tid; i < (sp.bias - sp.w1) / 4; i += NT) reinterpret_cast<uint32_t*>(smem + sp.w1)[i] = 0u;
  __syncthreads();
  mbar_wait(bar, 0);
  {
    const float* P = sStage;
    for (int i = tid; i < 64 * E; i += NT) sW1[(i / E) * LDX0 + (i % E)] = __float2half_rn(P[a.po[0] + i]);
    for (int i = tid; i < 16 * 64; i += NT) sW2[(i / 64) * LD64 + (i % 64)] = __float2half_rn(P[a.po[2] + i]);
    const int K3 = V + 15;
    for (int i = tid; i < 64 * K3; i += NT) sW3[(i / K3) * LD32 + (i % K3)] = __float2half_rn(P[a.po[4] + i]);
    for (int i = tid; i < 64 * 64; i += NT) sW4[(i / 64) * LD64 + (i % 64)] = __float2half_rn(P[a.po[6] + i]);
    for (int i = tid; i < 3 * 64; i += NT) sW5[(i / 64) * LD64 + (i % 64)] = __float2half_rn(P[a.po[8] + i]);
    // biases: autocast rounds them to fp16 as well (F.linear casts all three operands)
    for (int i = tid; i < 64; i += NT) sB[i] = __half2float(__float2half_rn(P[a.po[1] + i]));
    for (int i = tid; i < 16; i += NT) sB[64 + i] = __half2float(__float2half_rn(P[a.po[3] + i]));
    for (int i = tid; i < 64; i += NT) sB[80 + i] = __half2float(__float2half_rn(P[a.po[5] + i]));
    for (int i = tid; i < 64; i += NT) sB[144 + i] = __half2float(__float2half_rn(P[a.po[7] + i]));
    for (int i = tid; i < 8; i += NT) sB[208 + i] = (i < 3) ? __half2float(__float2half_rn(P[a.po[9] + i])) : 0.f;
  }
  __syncthreads();

  // weight-gradient accumulators: one item per layer per warp, kept in registers across all tiles of this CTA
  using S1 = WSplit<4, KE / 8, NWARP>;
  using S2 = WSplit<1, 8, NWARP>;
  using S3 = WSplit<4, KC / 8, NWARP>;
  using S4 = WSplit<4, 8, NWARP>;
  using S5 = WSplit<1, 8, NWARP>;
  float wg1[S1::CNT][4], wg2[S2::CNT][4], wg3[S3::CNT][4], wg4[S4::CNT][4], wg5[S5::CNT][4];
  float wb1[2] = {0.f, 0.f}, wb2[2] = {0.f, 0.f}, wb3[2] = {0.f, 0.f}, wb4[2] = {0.f, 0.f}, wb5[2] = {0.f, 0.f};
  zero_acc<S1::CNT>(wg1); zero_acc<S2::CNT>(wg2); zero_acc<S3::CNT>(wg3); zero_acc<S4::CNT>(wg4); zero_acc<S5::CNT>(wg5);
  float loss_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float n_valid_s = 0.f, n_valid_r = 0.f;
  bool overflow = false;
  // eikonal term: its weight over the number of selected samples (counted by eik_count_kernel before this launch),
  // times the loss scale; accumulators of d L_eik / d W2[0,:] (lanes with g8 == 0 own columns nt*8 + 2*t4, +1)
  float eik_loss = 0.f, w2e[8][2];
  int n_sel = 0;
  __half* Qh = reinterpret_cast<__half*>(smem + sp.qh);
  float* nbuf = reinterpret_cast<float*>(smem + sp.nbuf);
  const float eik_c = EIK ? a.p.eikonal_weight / fmaxf((float)*reinterpret_cast<const int*>(static_cast<const char*>(a.wpack) + kWPackBytes - 32), 1.f) : 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) w2e[nt][0] = w2e[nt][1] = 0.f;

  const int Sp = a.Sp, R = a.R, S = a.p.S;
  // point owned by this thread in the per-point phases: both halves of the CTA see the same points
  const int half = tid / PT;                 // 0: levels [0, LH) ; 1: levels [LH, L)
  const int pt = tid - half * PT;
  const int rl = pt / Sp, sidx = pt - rl * Sp;
  const int LH = (L + 1) >> 1;
  const int l_beg = half ? LH : 0, l_end = half ? L : LH;
  const bool owner = half == 0;              // the thread that does the once-per-point work (compositing, seeds)
  __half2* Jslot = reinterpret_cast<__half2*>(a.jws) + (size_t)blockIdx.x * (MAX_L * 3) * PT;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int row0 = warp * 16;                // MLP rows of this warp

  for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
    // ============ 1. ray setup
    if (tid < R) setup_ray(sRay[tid], a, grp * R + tid);
    __syncthreads();
    const RayS& rs = sRay[rl];
    const bool active = rs.active && sidx < S;
    const float z = active ? a.p.z_vals[(size_t)rs.ray * S + sidx] : 0.f;
    float pc[3], x[3], u[3];
    world_point(rs, z, pc, x);
    const bool valid = active && fabsf(x[0]) <= 1.f && fabsf(x[1]) <= 1.f && fabsf(x[2]) <= 1.f;   // nerf_runner.py:1245
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = (x[d] + 1.0f) * 0.5f;                                        // grid.py:160
    const float w_raw = active ? raw_weight(a, z, rs.depth) : 0.f;
    if (owner) {                                            // warp-uniform (PT is a multiple of 32)
      const float ws = warp_sum(w_raw);
      const unsigned anyv = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) {
        if (ws != 0.f) atomicAdd(&sRay[rl].sumw, ws);
        if (anyv) atomicOr(&sRay[rl].anyvalid, 1);
      }
    }
    // ============ 2. colour-net input row (views part; geo filled by L2) and this thread's half of the multires gather
    {
      if (!owner) {
        __half* xc = XC + (size_t)pt * LD32;
#pragma unroll 4
        for (int j = 0; j < KC; ++j) xc[j] = __float2half_rn(j < V ? rs.views[j] : 0.f);
      }
      __half* x0 = X0 + (size_t)pt * LDX0;
      if (valid) {
#pragma unroll kGatherUnroll
        for (int l = l_beg; l < l_end; ++l) {
          float enc[2], J[3][2];
          if (EIK || a.p.need_pose_grad) {
            gather_level<true, true>(a.p.table_f16, lv, l, u, enc, J);
#pragma unroll
            for (int d = 0; d < 3; ++d) Jslot[(size_t)(l * 3 + d) * PT + pt] = __floats2half2_rn(J[d][0], J[d][1]);
          } else {
            gather_level<true, false>(a.p.table_f16, lv, l, u, enc, J);
          }
          *reinterpret_cast<__half2*>(x0 + 2 * l) = __floats2half2_rn(enc[0], enc[1]);
        }
        if (owner) for (int j = E; j < KE; ++j) x0[j] = __float2half_rn(0.f);
      } else {
        for (int j = 2 * l_beg; j < 2 * l_end; j += 2) *reinterpret_cast<uint32_t*>(x0 + j) = 0u;   // nerf_runner.py:1247: zeros for invalid
        if (owner) for (int j = E; j < KE; ++j) x0[j] = __float2half_rn(0.f);
      }
    }
    __syncthreads();                                        // (A) rows are produced by two different warps
    // ============ eikonal term (the INTENDED maths of nerf_runner.py:734-738, 1297-1302, 1342-1345, defined by oracle.eikonal_loss):
    // n = d sdf / d x = 1/2 J^T q with q = W1^T (relu'(h1) * W2[0,:]) the gradient of the sdf w.r.t. the encoding; L = w/C sum (|n| - 1)^2
    // over the samples with sdf < 1 (C of them, counted beforehand). With g = dL/dn: dL/dq = a = 1/2 J g, dL/dW2[0,:] = relu'(h1) * (W1 a),
    // dL/dW1 = (relu'(h1) * W2[0,:]) (x) a, and the table receives 1/2 q_c g_d scale_l dw_k/df_d per corner (scatter). x is detached.
    float eg[3] = {0.f, 0.f, 0.f};                          // g of this thread's point (both threads of a point hold it)
    auto eikonal_block = [&]() {
      __half* XR = X3;                                      // relu'(h1) * W2[0,:]  (X3 / X4 are free until layers 3 / 4 of the forward)
      __half* XA = X4;                                      // a, row stride LDX0
      {                                                     // e1: own 16 rows, lane = (row, half of the 64 columns)
        const int r = row0 + (lane >> 1), c0 = (lane & 1) * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const __half2 m = *reinterpret_cast<const __half2*>(X1 + (size_t)r * LD64 + c0 + j);
          const __half2 w = *reinterpret_cast<const __half2*>(sW2 + c0 + j);
          const uint32_t keep = __hgt2_mask(m, __float2half2_rn(0.f));
          *reinterpret_cast<uint32_t*>(XR + (size_t)r * LD64 + c0 + j) = *reinterpret_cast<const uint32_t*>(&w) & keep;
        }
      }
      __syncwarp();
      {                                                     // e2: q = r W1 on the tensor cores, stored as fp16
        constexpr int NT1 = KE / 8;
        float acc[NT1][4];
        zero_acc<NT1>(acc);
        warp_dgrad<64, KE>(XR + (size_t)row0 * LD64, LD64, sW1, LDX0, acc, lane);
#pragma unroll
        for (int nt = 0; nt < NT1; ++nt) {
          __half* y = Qh + (size_t)(row0 + g8) * LDX0 + nt * 8 + 2 * t4;
          *reinterpret_cast<uint32_t*>(y) = pack_h2(acc[nt][0], acc[nt][1]);
          *reinterpret_cast<uint32_t*>(y + 8 * LDX0) = pack_h2(acc[nt][2], acc[nt][3]);
        }
      }
      __syncthreads();                                      // (E1) q of every point visible to the point's two threads
      {                                                     // e3: this thread's levels of n = 1/2 J^T q
        float nh[3] = {0.f, 0.f, 0.f};
        if (valid) {
          for (int l = l_beg; l < l_end; ++l) {
            const float2 q = __half22float2(*reinterpret_cast<const __half2*>(Qh + (size_t)pt * LDX0 + 2 * l));
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float2 jj = __half22float2(Jslot[(size_t)(l * 3 + d) * PT + pt]);
              nh[d] = fmaf(q.x, jj.x, fmaf(q.y, jj.y, nh[d]));
            }
          }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) nbuf[(size_t)(half * PT + pt) * 3 + d] = 0.5f * nh[d];
      }
      __syncthreads();                                      // (E2)
      {                                                     // e4: |n|, selection (sdf < 1; out-of-bounds samples have sdf = 0, n = 0), g = dL/dn
        float n[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) n[d] = nbuf[(size_t)pt * 3 + d] + nbuf[(size_t)(PT + pt) * 3 + d];
        const float sdf = sOut[pt * 4 + 3];
        if (active) {
          if (!valid) {
            if (owner) eik_loss += eik_c;                   // (0 - 1)^2
          } else if (sdf < 1.f) {
            const float nn = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (owner) eik_loss += eik_c * (nn - 1.f) * (nn - 1.f);
            const float coef = nn > 0.f ? eik_c * 2.f * (nn - 1.f) / nn * scale_ls : 0.f;      // the norm's subgradient at 0 is 0 (torch)
#pragma unroll
            for (int d = 0; d < 3; ++d) eg[d] = coef * n[d];
          }
        }
        // e5: a = 1/2 J g for this thread's levels -> XA (fp16 operand of the next two GEMMs)
        __half* xa = XA + (size_t)pt * LDX0;
        for (int l = l_beg; l < l_end; ++l) {
          float a0 = 0.f, a1 = 0.f;
          if (valid && (eg[0] != 0.f || eg[1] != 0.f || eg[2] != 0.f)) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float2 jj = __half22float2(Jslot[(size_t)(l * 3 + d) * PT + pt]);
              a0 = fmaf(jj.x, eg[d], a0);
              a1 = fmaf(jj.y, eg[d], a1);
            }
            a0 *= 0.5f; a1 *= 0.5f;
            overflow |= !(fabsf(a0) <= 65504.f) || !(fabsf(a1) <= 65504.f);
          }
          *reinterpret_cast<uint32_t*>(xa + 2 * l) = pack_h2(a0, a1);
        }
        if (owner) for (int j = E; j < KE; ++j) xa[j] = __float2half_rn(0.f);
      }
      __syncthreads();                                      // (E3) a of every row complete
      {                                                     // e6: t = a W1^T on own rows; dW2[0,:] += column sums of relu'(h1) * t
        float acc[8][4];
        zero_acc<8>(acc);
        warp_fwd<KE, 64>(XA + (size_t)row0 * LDX0, LDX0, sW1, LDX0, acc, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const __half2 m0 = *reinterpret_cast<const __half2*>(X1 + (size_t)(row0 + g8) * LD64 + nt * 8 + 2 * t4);
          const __half2 m1 = *reinterpret_cast<const __half2*>(X1 + (size_t)(row0 + g8 + 8) * LD64 + nt * 8 + 2 * t4);
          float s0 = (__low2float(m0) > 0.f ? acc[nt][0] : 0.f) + (__low2float(m1) > 0.f ? acc[nt][2] : 0.f);
          float s1 = (__high2float(m0) > 0.f ? acc[nt][1] : 0.f) + (__high2float(m1) > 0.f ? acc[nt][3] : 0.f);
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          }
          w2e[nt][0] += s0;
          w2e[nt][1] += s1;
        }
      }
      // e7: dW1 += r^T a over all rows of the tile (same split as the layer-1 weight gradient, no bias part)
      if (warp < S1::ITEMS) {
        float nob[2] = {0.f, 0.f};
        wgrad_item<S1::CNT>(XR, LD64, XA, LDX0, PT, warp / S1::GROUPS, (warp % S1::GROUPS) * S1::CNT, wg1, nob, false, lane);
      }
      __syncthreads();                                      // (E4) X3 / X4 are free again for layers 3 / 4
    };
    // ============ 3. MLP forward on the warp's own 16 rows
    {
      float acc[8][4];
      // ---- L1: E -> 64, ReLU
      init_bias<8>(acc, sB, t4);
      warp_fwd<KE, 64>(X0 + (size_t)row0 * LDX0, LDX0, sW1, LDX0, acc, lane);
      store_relu64(X1, row0, acc, g8, t4);
      __syncwarp();
      // ---- L2: 64 -> 16 (sdf | geo 15), no activation
      float acc2[2][4];
      init_bias<2>(acc2, sB + 64, t4);
      warp_fwd<64, 16>(X1 + (size_t)row0 * LD64, LD64, sW2, LD64, acc2, lane);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = row0 + g8 + h * 8;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col = nt * 8 + 2 * t4 + c;
            const __half hv = __float2half_rn(acc2[nt][h * 2 + c]);
            if (col == 0) sOut[r * 4 + 3] = __half2float(hv);              // sdf (nerf_helpers.py:312)
            else XC[(size_t)r * LD32 + V + col - 1] = hv;                   // geo feature -> colour-net input (:316)
          }
        }
      __syncwarp();
      if constexpr (EIK) {
        if (a.count_only) {                                 // counting pass: the sdf of every sample is all it needs (same arithmetic as the real pass)
          __syncthreads();
          if (owner && active && (!valid || sOut[pt * 4 + 3] < 1.f)) n_sel += 1;
          __syncthreads();                                  // sOut / ray state free for the next tile
          continue;
        }
        eikonal_block();
      }
      // ---- L3: (V+15 padded 32) -> 64, ReLU
      init_bias<8>(acc, sB + 80, t4);
      warp_fwd<KC, 64>(XC + (size_t)row0 * LD32, LD32, sW3, LD32, acc, lane);
      store_relu64(X3, row0, acc, g8, t4);
      __syncwarp();
      // ---- L4: 64 -> 64, ReLU
      init_bias<8>(acc, sB + 144, t4);
      warp_fwd<64, 64>(X3 + (size_t)row0 * LD64, LD64, sW4, LD64, acc, lane);
      store_relu64(X4, row0, acc, g8, t4);
      __syncwarp();
      // ---- L5: 64 -> 3 (padded 8)
      float acc5[1][4];
      init_bias<1>(acc5, sB + 208, t4);
      warp_fwd<64, 8>(X4 + (size_t)row0 * LD64, LD64, sW5, LD64, acc5, lane);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 2 * t4 + c;
          if (col < 3) sOut[(row0 + g8 + h * 8) * 4 + col] = __half2float(__float2half_rn(acc5[0][h * 2 + c]));
        }
    }
    __syncthreads();                                        // (B) sumw / anyvalid complete, sOut rows visible
    // ============ 4. compositing (nerf_runner.py:1163-1167) — once per point (owner threads)
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
    }
    __syncthreads();                                        // (C) rgb_map complete
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
        loss_acc[1] += a.p.rgb_weight * e * ray_w * a.inv_N3;                     // nerf_runner.py:700-701
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
      __half* dr = DO + (size_t)pt * LD16;
      *reinterpret_cast<uint32_t*>(dr) = pack_h2(s0, s1);
      *reinterpret_cast<uint32_t*>(dr + 2) = pack_h2(s2, 0.f);
#pragma unroll
      for (int j = 4; j < 16; j += 2) *reinterpret_cast<uint32_t*>(dr + j) = 0u;
    }
    __syncthreads();                                        // (S0) dOut rows visible to every warp

    // ============ 6. backward through the MLP
    // ---- layer 5: wgrad (all rows) + dgrad (own rows) -> dY4 = dX4 * relu'(X4), in place
    if (warp < S5::ITEMS)
      wgrad_item<S5::CNT>(DO, LD16, X4, LD64, PT, 0, (warp % S5::GROUPS) * S5::CNT, wg5, wb5, (warp % S5::GROUPS) == 0, lane);
    {
      float acc[8][4];
      zero_acc<8>(acc);
      warp_dgrad<16, 64>(DO + (size_t)row0 * LD16, LD16, sW5, LD64, acc, lane);
      __syncthreads();                                      // every warp finished reading X4 (wgrad5)
      overflow |= store_masked64(X4, row0, acc, g8, t4);
    }
    __syncthreads();                                        // dY4 visible
    // ---- layer 4
    if (warp < S4::ITEMS)
      wgrad_item<S4::CNT>(X4, LD64, X3, LD64, PT, warp / S4::GROUPS, (warp % S4::GROUPS) * S4::CNT, wg4, wb4, (warp % S4::GROUPS) == 0, lane);
    {
      float acc[8][4];
      zero_acc<8>(acc);
      warp_dgrad<64, 64>(X4 + (size_t)row0 * LD64, LD64, sW4, LD64, acc, lane);
      __syncthreads();                                      // wgrad4 finished reading X3
      overflow |= store_masked64(X3, row0, acc, g8, t4);
    }
    __syncthreads();                                        // dY3 visible
    // ---- layer 3: wgrad (dY3^T XC), dgrad -> [dviews | dgeo]
    if (warp < S3::ITEMS)
      wgrad_item<S3::CNT>(X3, LD64, XC, LD32, PT, warp / S3::GROUPS, (warp % S3::GROUPS) * S3::CNT, wg3, wb3, (warp % S3::GROUPS) == 0, lane);
    {
      float acc[4][4];
      zero_acc<4>(acc);
      warp_dgrad<64, KC>(X3 + (size_t)row0 * LD64, LD64, sW3, LD32, acc, lane);
      // dviews: sum over the 16 rows of this warp (samples of one ray) — warp-level reduction, then one shared
      // atomic per column per warp.
      const int ray_of_rows = row0 / Sp;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = nt * 8 + 2 * t4 + c;
          float v = acc[nt][c] + acc[nt][2 + c];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g8 == 0 && col < V && v != 0.f) atomicAdd(&sRay[ray_of_rows].dviews[col], v);
        }
      // dH2 = [dsdf | dgeo] -> DO buffer (dOut is dead: wgrad5/dgrad5 completed before the barriers above)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col = nt * 8 + 2 * t4 + c;
            if (col >= V && col < V + 15) {
              const float v = acc[nt][h * 2 + c];
              overflow |= !(fabsf(v) <= 65504.f);
              DO[(size_t)(row0 + g8 + h * 8) * LD16 + 1 + (col - V)] = __float2half_rn(v);
            }
          }
      if (owner) DO[(size_t)pt * LD16] = __float2half_rn(dsdf_s);
    }
    __syncthreads();                                        // dH2 visible
    // ---- layer 2
    if (warp < S2::ITEMS)
      wgrad_item<S2::CNT>(DO, LD16, X1, LD64, PT, 0, (warp % S2::GROUPS) * S2::CNT, wg2, wb2, (warp % S2::GROUPS) == 0, lane);
    {
      float acc[8][4];
      zero_acc<8>(acc);
      warp_dgrad<16, 64>(DO + (size_t)row0 * LD16, LD16, sW2, LD64, acc, lane);
      __syncthreads();                                      // wgrad2 finished reading X1
      overflow |= store_masked64(X1, row0, acc, g8, t4);
    }
    __syncthreads();                                        // dY1 visible
    // ---- layer 1: wgrad (dY1^T X0), dgrad -> dEnc (fp32, scaled) into the X3 region (dead)
    {
      constexpr int NT1 = KE / 8;
      if (warp < S1::ITEMS)
        wgrad_item<S1::CNT>(X1, LD64, X0, LDX0, PT, warp / S1::GROUPS, (warp % S1::GROUPS) * S1::CNT, wg1, wb1, (warp % S1::GROUPS) == 0, lane);
      float acc[NT1][4];
      zero_acc<NT1>(acc);
      warp_dgrad<64, KE>(X1 + (size_t)row0 * LD64, LD64, sW1, LDX0, acc, lane);
      float* dE = reinterpret_cast<float*>(X3);             // [PT][36] fp32 (144-byte rows)
#pragma unroll
      for (int nt = 0; nt < NT1; ++nt) {
        float* y = dE + (size_t)(row0 + g8) * 36 + nt * 8 + 2 * t4;
        *reinterpret_cast<float2*>(y) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(y + 8 * 36) = make_float2(acc[nt][2], acc[nt][3]);
      }
    }
    __syncthreads();                                        // dEnc rows visible to the two threads of each point
    // ============ 7. grid-gradient scatter + pose Jacobian (path A: positions) for this thread's half of the levels
    {
      const float* dE = reinterpret_cast<const float*>(X3) + (size_t)pt * 36;
      float gx[3] = {0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll 2
        for (int l = l_beg; l < l_end; ++l) {
          const float2 g = *reinterpret_cast<const float2*>(dE + 2 * l);
          if constexpr (EIK) {
            const float2 q = __half22float2(*reinterpret_cast<const __half2*>(Qh + (size_t)pt * LDX0 + 2 * l));
            if (g.x != 0.f || g.y != 0.f || eg[0] != 0.f || eg[1] != 0.f || eg[2] != 0.f) scatter_level_eik(a.p.grad_table, lv, l, u, g.x, g.y, q.x, q.y, eg);
          } else
          if (g.x != 0.f || g.y != 0.f) scatter_level(a.p.grad_table, lv, l, u, g.x, g.y);
          if (a.p.need_pose_grad) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float2 j = __half22float2(Jslot[(size_t)(l * 3 + d) * PT + pt]);
              gx[d] = fmaf(g.x, j.x, fmaf(g.y, j.y, gx[d]));
            }
          }
        }
      }
      if (a.p.need_pose_grad) {
        // dL/dx = 0.5 * dL/du ; dL/dR[i][j] += gx[i]*pc[j] ; dL/dt[i] += gx[i]   (x = R pc + t)
        // x = R (dir z) + t  =>  dL/dR[i][j] = dir[j] * sum_p gi z ,  dL/dt[i] = sum_p gi   (gi = 0.5 gx[i]: u = (x+1)/2):
        // six warp sums instead of twelve (all lanes of a warp belong to one ray).
        float st[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float gi = 0.5f * gx[i];
          st[i] = warp_sum(gi);
          st[3 + i] = warp_sum(gi * z);
        }
        if (lane == 0 && rs.active && rs.frame != 0) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float v = st[3 + i] * rs.dir[j];
              if (v != 0.f) red_add(a.p.grad_tf + (size_t)rs.frame * 12 + i * 4 + j, v);
            }
            if (st[i] != 0.f) red_add(a.p.grad_tf + (size_t)rs.frame * 12 + i * 4 + 3, st[i]);
          }
        }
      }
    }
    __syncthreads();                                        // dviews complete; rays/buffers free for the next tile
    // ---- path B: view directions (per ray): dviews -> dfeat, dSH -> d(dir_w) -> dR
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

  // ============ flush: weight gradients (registers -> global, fp32 atomics), losses, flags
  {
    float* G = a.p.grad_mlp;
    const int K3 = V + 15;
    if (warp < S1::ITEMS)
      flush_item<S1::CNT>(G, a.po[0], a.po[1], E, 64, warp / S1::GROUPS, (warp % S1::GROUPS) * S1::CNT, wg1, wb1, (warp % S1::GROUPS) == 0, g8, t4);
    if (warp < S2::ITEMS)
      flush_item<S2::CNT>(G, a.po[2], a.po[3], 64, 16, 0, (warp % S2::GROUPS) * S2::CNT, wg2, wb2, (warp % S2::GROUPS) == 0, g8, t4);
    if (warp < S3::ITEMS)
      flush_item<S3::CNT>(G, a.po[4], a.po[5], K3, 64, warp / S3::GROUPS, (warp % S3::GROUPS) * S3::CNT, wg3, wb3, (warp % S3::GROUPS) == 0, g8, t4);
    if (warp < S4::ITEMS)
      flush_item<S4::CNT>(G, a.po[6], a.po[7], 64, 64, warp / S4::GROUPS, (warp % S4::GROUPS) * S4::CNT, wg4, wb4, (warp % S4::GROUPS) == 0, g8, t4);
    if (warp < S5::ITEMS)
      flush_item<S5::CNT>(G, a.po[8], a.po[9], 64, 3, 0, (warp % S5::GROUPS) * S5::CNT, wg5, wb5, (warp % S5::GROUPS) == 0, g8, t4);
  }
  {
    if constexpr (EIK) {
      if (a.count_only) {
        for (int o = 16; o > 0; o >>= 1) n_sel += __shfl_xor_sync(0xffffffffu, n_sel, o);
        if (lane == 0 && n_sel) atomicAdd(reinterpret_cast<int*>(static_cast<char*>(a.wpack) + kWPackBytes - 32), n_sel);
        return;
      }
      if (g8 == 0) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          if (w2e[nt][0] != 0.f) red_add(a.p.grad_mlp + a.po[2] + nt * 8 + 2 * t4, w2e[nt][0]);
          if (w2e[nt][1] != 0.f) red_add(a.p.grad_mlp + a.po[2] + nt * 8 + 2 * t4 + 1, w2e[nt][1]);
        }
      }
    }
    loss_acc[0] = loss_acc[1] + loss_acc[2] + loss_acc[3] + loss_acc[4] + eik_loss;
    float vals[8] = {loss_acc[0], loss_acc[1], loss_acc[2], loss_acc[3], loss_acc[4], n_valid_s, n_valid_r, eik_loss};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = warp_sum(vals[i]);
      if (lane == 0 && v != 0.f) red_add(a.p.losses + i, v);
    }
    const unsigned ov = __ballot_sync(0xffffffffu, overflow);
    if (lane == 0 && ov && a.p.found_inf) atomicExch(a.p.found_inf, 1);
  }
}

// ------------------------------------------------------------------------------------------------
size_t step_amp_smem(int PT, int KE, bool eik) { return (size_t)make_plan(PT, KE, eik).total; }

template <int PT, int KE, bool EIK>
static int launch_amp(const StepArgs& a, int blocks, cudaStream_t st) {
  const size_t smem = step_amp_smem(PT, KE, EIK);
  // function attributes are per device: set on every launch (a cheap host-side call) rather than once per process
  cudaFuncSetAttribute(step_amp_kernel<PT, KE, EIK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  step_amp_kernel<PT, KE, EIK><<<blocks, 2 * PT, smem, st>>>(a);
  return check_launch("step_amp_kernel");
}

// NW = points per tile / 32 (4, 6 or 8)
int step_amp_dispatch(const StepArgs& a, int NW, int blocks, cudaStream_t st) {
  const bool eik = a.p.eikonal_weight > 0.f;
#define NOF_AMP_CASE(nw)                                                          \
  case nw:                                                                        \
    if (a.KE == 32) return eik ? launch_amp<nw * 32, 32, true>(a, blocks, st) : launch_amp<nw * 32, 32, false>(a, blocks, st);   \
    if (a.KE == 16) return eik ? launch_amp<nw * 32, 16, true>(a, blocks, st) : launch_amp<nw * 32, 16, false>(a, blocks, st);   \
    break;
  switch (NW) {
    NOF_AMP_CASE(4)
    NOF_AMP_CASE(6)
    NOF_AMP_CASE(8)
    default: break;
  }
#undef NOF_AMP_CASE
  set_error("nof_step_fused(amp): unsupported tile NW=%d KE=%d", NW, a.KE);
  return NOF_E_UNSUPPORTED;
}

}  // namespace nof
