// Fused forward + loss + backward of one train step, AMP policy, tcgen05 edition (tile PT = 128 points, S <= 128).
//
// Same structure as nof_step_amp.cu (two threads per point, gather / compositing / seeds / scatter identical), but the
// ten chained GEMMs of the MLP forward and dgrad run on the 5th-generation tensor cores:
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
#include "nof_step_common.cuh"

namespace nof {
namespace tc {

// ------------------------------------------------------------------------------------------------ primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4_t(uint32_t r[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
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
// Bounded wait: never hangs the GPU. Returns false on timeout (the caller raises the device error flag).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t phase) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(phase)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type=0 [61,64)).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         (1ull << 46);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=f16 (0), a_major bit15, b_major bit16,
// N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N> struct TmemLd;
template <> struct TmemLd<8> {
  static __device__ __forceinline__ void ld(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
  }
};
template <> struct TmemLd<16> {
  static __device__ __forceinline__ void ld(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
  }
};
template <> struct TmemLd<32> {
  static __device__ __forceinline__ void ld(uint32_t taddr, float* v) {
    TmemLd<16>::ld(taddr, v);
    TmemLd<16>::ld(taddr + 16, v + 16);
  }
};

// ------------------------------------------------------------------------------------------------ core-matrix layout
// element (r, k) of an [R x K] fp16 matrix: 8 rows x 8 columns (16 bytes per row) form one contiguous 128-byte core matrix;
// core matrices are ordered k-chunk fastest: byte offset = (r/8)*(K*16) + (k/8)*128 + (r%8)*16 + (k%8)*2.
__device__ __forceinline__ uint32_t cm_off(int r, int k, int K) { return (uint32_t)((r >> 3) * (K * 16) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2); }

constexpr int PT = 128, NT = 256, NWARP = 8;
constexpr int DES = PT + PT / 8;                 // padded point stride of the transposed dEnc / z arrays: index pt + pt/8
__device__ __forceinline__ int des_idx(int pt) { return pt + (pt >> 3); }

struct Plan {
  int w1, w2, w3, w4, w5, bias, x0, x1, xc, x3, x4, d_o, out, zs, rays, lv, bar, tmem, total;
};
__host__ __device__ inline Plan make_plan(int KE) {
  Plan s;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) / 128 * 128; return r; };
  s.w1 = take(64 * KE * 2);
  s.w2 = take(16 * 64 * 2);
  s.w3 = take(64 * KC * 2);
  s.w4 = take(64 * 64 * 2);
  s.w5 = take(16 * 64 * 2);
  s.bias = take(216 * 4);
  s.x0 = take(PT * KE * 2);
  s.x1 = take(PT * 64 * 2);
  s.xc = take(PT * KC * 2);
  s.x3 = take(PT * 64 * 2);                 // x3|x4 also hold dEnc fp32 [KE][DES] (transposed) at the end of the backward
  s.x4 = take(PT * 64 * 2);
  s.d_o = take(PT * 16 * 2);
  s.out = take(PT * 4 * 4);
  s.zs = take(5 * DES * 4);                  // z, valid flag and u[3] of every point, in the scatter's padded order
  s.rays = take(MAX_R * (int)sizeof(RayS));
  s.lv = take((int)sizeof(LevelS));
  s.bar = take(64);                          // [0] TMA staging barrier, [1] MMA completion barrier
  s.tmem = take(16);
  s.total = o;
  return s;
}

// wgrad on mma.sync reading core-matrix buffers (same balanced split as nof_step_amp.cu): dW[strip*16..+16][nt0*8..] += dY^T X
template <int NTU>
__device__ __forceinline__ void wgrad_item(uint32_t dY, int Ky, uint32_t X, int Kx, int strip, int nt0, float (*acc)[4], float* bias2,
                                           bool do_bias, int lane) {
#ifdef NOF_EXP_NO_WGRAD
  return;
#endif
  const uint32_t ones = 0x3C003C00u;
  const int pa = (lane & 7) + (lane >> 4) * 8, oa = strip * 16 + ((lane >> 3) & 1) * 8;       // A: rows p, cols o (dY^T)
  const int pb = (lane & 7) + ((lane >> 3) & 1) * 8, ib = (lane >> 4) * 8;                      // B: rows p, cols i
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
                                           const float* bias2, bool has_bias, int g8, int t4) {
#pragma unroll
  for (int nt = 0; nt < CNT; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = strip * 16 + g8 + h * 8, i = (nt0 + nt) * 8 + 2 * t4;
      const float v0 = acc[nt][h * 2], v1 = acc[nt][h * 2 + 1];
      if (o >= nrows || i >= ncols) continue;
      const size_t e = (size_t)wofs + (size_t)o * ncols + i;
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

// ------------------------------------------------------------------------------------------------ operand packing
// fp32 packed parameters (nof_mlp_param_offsets order) -> the kernel's shared-memory image [W1 W2 W3 W4 W5 | biases]: fp16
// weights in core-matrix order, zero padding, biases rounded through fp16 like torch autocast. One tiny kernel per step; the
// 296 step CTAs then fetch the 21.5 KB image with one TMA bulk copy each instead of converting 9.6k weights each.
template <int KE>
__global__ void __launch_bounds__(256) pack_mlp_kernel(const StepArgs a) {
  const Plan sp = make_plan(KE);
  unsigned char* out = static_cast<unsigned char*>(a.wpack);      // image offset 0 == plan offset sp.w1 (== 0)
  const int gid = blockIdx.x * 256 + threadIdx.x;
  if (gid == 0) *reinterpret_cast<int*>(out + kWPackBytes - 16) = 0;                 // tile ticket of the step kernel
  if (gid < 8) a.p.losses[gid] = 0.f;                                                // per-step results start from zero
  if (a.p.grad_tf) for (int i = gid; i < a.p.F * 12; i += gridDim.x * 256) a.p.grad_tf[i] = 0.f;
  const float* P = a.p.mlp;
  const int o = gid * 4;                                          // one 32-bit word of the image per thread: gather form, no zero pass
  if (o < sp.bias) {
    int base, rows, kreal, kpad, po;
    if (o >= sp.w5) { base = sp.w5; rows = 3; kreal = 64; kpad = 64; po = a.po[8]; }
    else if (o >= sp.w4) { base = sp.w4; rows = 64; kreal = 64; kpad = 64; po = a.po[6]; }
    else if (o >= sp.w3) { base = sp.w3; rows = 64; kreal = a.V + 15; kpad = KC; po = a.po[4]; }
    else if (o >= sp.w2) { base = sp.w2; rows = 16; kreal = 64; kpad = 64; po = a.po[2]; }
    else { base = sp.w1; rows = 64; kreal = a.E; kpad = KE; po = a.po[0]; }
    const int rel = o - base, rg = rel / (kpad * 16), rem = rel % (kpad * 16);
    const int n = rg * 8 + (rem % 128) / 16, k = (rem / 128) * 8 + (rem % 16) / 2;          // inverse of cm_off
    const float v0 = (n < rows && k < kreal) ? P[po + n * kreal + k] : 0.f;
    const float v1 = (n < rows && k + 1 < kreal) ? P[po + n * kreal + k + 1] : 0.f;
    *reinterpret_cast<uint32_t*>(out + o) = pack_h2(v0, v1);
  } else if (o < sp.x0) {
    const int j = (o - sp.bias) / 4;
    float v = 0.f;
    if (j < 64) v = P[a.po[1] + j];
    else if (j < 80) v = P[a.po[3] + j - 64];
    else if (j < 144) v = P[a.po[5] + j - 80];
    else if (j < 208) v = P[a.po[7] + j - 144];
    else if (j < 211) v = P[a.po[9] + j - 208];
    *reinterpret_cast<float*>(out + o) = __half2float(__float2half_rn(v));
  }
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
  RayS* sRay = reinterpret_cast<RayS*>(smem + sp.rays);
  LevelS& lv = *reinterpret_cast<LevelS*>(smem + sp.lv);
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + sp.bar);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + sp.tmem);
  int* s_next = reinterpret_cast<int*>(s_tmem + 1);
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
  const uint32_t pack_bytes = (uint32_t)(sp.x0 - sp.w1);
  if (tid == 0) {
    mbar_expect_tx(bar_tma, pack_bytes);
    tma_bulk_g2s(smem + sp.w1, a.wpack, pack_bytes, bar_tma);
  }
  init_levels(lv, a);
  bool tc_ok = mbar_wait(bar_tma, 0);
  fence_async_smem();
  __syncthreads();

  using S1 = WSplit<4, KE / 8>;
  using S2 = WSplit<1, 8>;
  using S3 = WSplit<4, KC / 8>;
  using S4 = WSplit<4, 8>;
  using S5 = WSplit<1, 8>;
  float wg1[S1::CNT][4], wg2[S2::CNT][4], wg3[S3::CNT][4], wg4[S4::CNT][4], wg5[S5::CNT][4];
  float wb1[2] = {0.f, 0.f}, wb2[2] = {0.f, 0.f}, wb3[2] = {0.f, 0.f}, wb4[2] = {0.f, 0.f}, wb5[2] = {0.f, 0.f};
  zero_acc<S1::CNT>(wg1); zero_acc<S2::CNT>(wg2); zero_acc<S3::CNT>(wg3); zero_acc<S4::CNT>(wg4); zero_acc<S5::CNT>(wg5);
  float loss_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float n_valid_s = 0.f, n_valid_r = 0.f;
  bool overflow = false;
  uint32_t phase = 0;                        // parity of bar_mma

  const int Sp = a.Sp, R = a.R, S = a.p.S;
  const int half = tid / PT;                 // gather/scatter: which half of the levels; epilogues: which half of the columns
  const int pt = tid - half * PT;            // = 32*(warp%4) + lane: the TMEM lane this thread may read
  const int rl = pt / Sp, sidx = pt - rl * Sp;
  const int LH = (L + 1) >> 1;
  const int l_beg = half ? LH : 0, l_end = half ? L : LH;
  const bool owner = half == 0;
  __half2* Jslot = reinterpret_cast<__half2*>(a.jws) + (size_t)blockIdx.x * (MAX_L * 3) * PT;
  const int g8 = lane >> 2, t4 = lane & 3;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);     // TMEM address of this warp's lane quadrant

  // wait for the MMA generation, make TMEM readable
  auto mma_wait = [&]() {
#ifdef NOF_EXP_NO_TC
    return;
#endif
    tc_ok &= mbar_wait(bar_mma, phase);
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
  for (int grp = blockIdx.x; grp < a.n_groups;) {
    // ============ 1. ray setup
    if (tid < R) setup_ray(sRay[tid], a, grp * R + tid);
    if (tid == NT - 1) *s_next = (int)gridDim.x + atomicAdd(tile_ticket, 1);
    __syncthreads();
    grp = *s_next;                                          // the NEXT tile (this one's rays are already staged)
    const RayS& rs = sRay[rl];
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
    // ============ 2. colour-net input row (views; geo comes from L2) and this thread's half of the gather
    if (!owner) {
#pragma unroll
      for (int ch = 0; ch < KC / 8; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = ch * 8 + 2 * j;
          w[j] = pack_h2(k < V ? rs.views[k] : 0.f, k + 1 < V ? rs.views[k + 1] : 0.f);
        }
        *reinterpret_cast<uint4*>(pXC + cm_off(pt, ch * 8, KC)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (valid) {
#pragma unroll 2
      for (int l = l_beg; l < l_end; ++l) {
        float enc[2], J[3][2];
        if (a.p.need_pose_grad) {
          gather_level<true, true>(a.p.table_f16, lv, l, u, enc, J);
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
        else *reinterpret_cast<__half*>(pXC + cm_off(pt, V + col - 1, KC)) = hv;
      }
    }
    sync_for_mma();
    // ---- L3: (V+15 padded 32) -> 64, ReLU
    if (tid == 0) issue_gemm<64, KC, 0>(tmem, aXC, KC, aW3, KC, bar_mma);
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
          w[j] = pack_h2(fmaxf(v[ch * 8 + 2 * j] + sB[80 + c], 0.f), fmaxf(v[ch * 8 + 2 * j + 1] + sB[80 + c + 1], 0.f));
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
    tc_fence_before();
    __syncthreads();                                        // (B) sumw / anyvalid complete, sOut rows visible
    // ============ 4. compositing — once per point (owner threads)
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
        const __half2* mh = reinterpret_cast<const __half2*>(&m);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = __low2float(mh[j]) > 0.f ? v[ch * 8 + 2 * j] : 0.f, v1 = __high2float(mh[j]) > 0.f ? v[ch * 8 + 2 * j + 1] : 0.f;
          overflow |= !(fabsf(v0) <= 65504.f) || !(fabsf(v1) <= 65504.f);
          o[j] = pack_h2(v0, v1);
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
        const __half2* mh = reinterpret_cast<const __half2*>(&m);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = __low2float(mh[j]) > 0.f ? v[ch * 8 + 2 * j] : 0.f, v1 = __high2float(mh[j]) > 0.f ? v[ch * 8 + 2 * j + 1] : 0.f;
          overflow |= !(fabsf(v0) <= 65504.f) || !(fabsf(v1) <= 65504.f);
          o[j] = pack_h2(v0, v1);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    sync_for_mma();                                         // dY3 visible
    // ---- layer 3: dgrad -> [dviews | dgeo | pad], wgrad dY3^T XC
    if (tid == 0) issue_gemm<KC, 64, 1>(tmem, aX3, 64, aW3, KC, bar_mma);
    if (warp < S3::ITEMS)
      wgrad_item<S3::CNT>(aX3, 64, aXC, KC, warp / S3::GROUPS, (warp % S3::GROUPS) * S3::CNT, wg3, wb3, (warp % S3::GROUPS) == 0, lane);
    mma_wait();
    {
      float v[16];
      TmemLd<16>::ld(trow + half * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = half * 16 + j;
        if (col < V) {                                       // warp-uniform: all 32 rows of the warp belong to one ray
          const float s = warp_sum(v[j]);
          if (lane == 0 && s != 0.f) atomicAdd(&sRay[rl].dviews[col], s);
        } else if (col < V + 15) {
          overflow |= !(fabsf(v[j]) <= 65504.f);
          *reinterpret_cast<__half*>(pDO + cm_off(pt, 1 + (col - V), 16)) = __float2half_rn(v[j]);
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
        const __half2* mh = reinterpret_cast<const __half2*>(&m);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = __low2float(mh[j]) > 0.f ? v[ch * 8 + 2 * j] : 0.f, v1 = __high2float(mh[j]) > 0.f ? v[ch * 8 + 2 * j + 1] : 0.f;
          overflow |= !(fabsf(v0) <= 65504.f) || !(fabsf(v1) <= 65504.f);
          o[j] = pack_h2(v0, v1);
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
      const int sg = lane & 15, l = 2 * warp + (lane >> 4);
      const int p0 = 8 * sg;
      const RayS& r8 = sRay[p0 / Sp];
      float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (l < L) {
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
  static_assert((size_t)(64 * KE * 2 + 16 * 64 * 2 + 64 * KC * 2 + 64 * 64 * 2 + 16 * 64 * 2 + 216 * 4 + 6 * 128 + 16) <= kWPackBytes, "wpack too small");
  tc::pack_mlp_kernel<KE><<<(tc::make_plan(KE).x0 / 4 + 255) / 256, 256, 0, st>>>(a);
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
