// C-ABI glue: error reporting, device info, packed-parameter layout, nof_step_fused dispatch, SDF-only query.
#include <stdarg.h>

#include <algorithm>

#include "nof_step_common.cuh"

#ifndef NOF_AMP_IMPL_DEFAULT
#define NOF_AMP_IMPL_DEFAULT 3
#endif

namespace nof {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();   // clear the sticky-less error so the next call starts clean
    return NOF_E_LAUNCH;
  }
  return NOF_OK;
}

size_t step_amp_smem(int T, int KE, bool eik);
size_t step_f32_smem(int E, int V, int grp_pts);
int step_amp_dispatch(const StepArgs& a, int NW, int blocks, cudaStream_t st);
int step_f32_dispatch(const StepArgs& a, int blocks, cudaStream_t st);
size_t step_tc_smem(int KE);
int step_tc_dispatch(const StepArgs& a, int blocks, cudaStream_t st);
size_t step_ws_smem(int KE);
size_t step_ws_jscratch(int blocks);
bool step_ws_tiling(int S, int* Sp_out, int* R_out);
int step_ws_dispatch(const StepArgs& a, int blocks, cudaStream_t st);

// The AMP step has three implementations, all behind nof_step_fused:
//   1 "tc"  : tcgen05 tile kernel, every warp walks the phases of one 128-point tile together, 2 CTAs/SM (nof_step_tc.cu; S <= 128)
//   0 "mma" : the same structure on mma.sync with 128/192/256-point tiles (nof_step_amp.cu; S <= 256)
//   2 "ws"  : warp-specialised streaming pipeline on tcgen05, rays may span tiles (nof_step_ws.cu; S <= 384)
//   3 "auto": the fastest measured on B200 for the given S (profiles/README.md): tc for S <= 128, mma for S <= 256, ws above (default).
// Forcing 0/1/2 (nof_set_amp_impl, env NOF_AMP_IMPL=mma|tc|ws) keeps the three as in-process cross-checks of one another.
static int g_amp_impl = -1;
static int amp_impl() {
  if (g_amp_impl < 0) {
    const char* e = getenv("NOF_AMP_IMPL");
    g_amp_impl = (e && strcmp(e, "mma") == 0) ? 0 : (e && strcmp(e, "tc") == 0) ? 1 : (e && strcmp(e, "ws") == 0) ? 2 : NOF_AMP_IMPL_DEFAULT;
  }
  return g_amp_impl;
}
// implementation actually used for S samples per ray under the current setting
static int amp_impl_for(int S) {
  const int m = amp_impl();
  if (m == 2 || (m == 3 && S > 256)) return 2;
  if (m == 0 || S > 128) return S <= 256 ? 0 : 2;          // tc carries S <= 128 only; mma S <= 256; beyond that only ws exists
  return 1;
}

static void mlp_offsets(int E, int V, int32_t o[10], size_t* total) {
  const int sizes[10] = {64 * E, 64, 16 * 64, 16, 64 * (V + 15), 64, 64 * 64, 64, 3 * 64, 3};
  int acc = 0;
  for (int i = 0; i < 10; ++i) { o[i] = acc; acc += sizes[i]; }
  if (total) *total = (size_t)acc;
}

struct Tiling { int NW, Sp, R, n_groups, blocks; };

// AMP: one CTA = R whole rays, thread = sample; supported CTA sizes 128/192/256/320 threads.
static bool amp_tiling(const NofStep* p, int sms, Tiling* t) {
  const int S = p->S;
  int T;
  // 656 B of shared memory per point + 25 KB of weights: 256 points (193 KB) is the largest tile that fits 227 KB
  if (S <= 128) T = 128; else if (S <= 192) T = 192; else if (S <= 256) T = 256; else return false;
  int Sp = T, R = 1;
  if (S <= 32) { Sp = 32; R = 4; } else if (S <= 64) { Sp = 64; R = 2; }
  t->NW = T / 32; t->Sp = Sp; t->R = R;
  t->n_groups = (p->N + R - 1) / R;
  const int per_sm = (t->NW <= 4) ? 2 : 1;
  t->blocks = std::max(1, std::min(t->n_groups, sms * per_sm));
  return true;
}
// fp32: CTA = 128 threads walking a ray group in sub-tiles of 128 points.
static bool f32_tiling(const NofStep* p, int sms, Tiling* t) {
  const int Sp = (p->S + 31) / 32 * 32;
  int R = 1;
  while ((R * Sp) % 128 != 0) R *= 2;
  if (R > MAX_R || R * Sp > 1024) return false;
  t->NW = 4; t->Sp = Sp; t->R = R;
  t->n_groups = (p->N + R - 1) / R;
  t->blocks = std::max(1, std::min(t->n_groups, sms));
  return true;
}

}  // namespace nof

using namespace nof;

extern "C" int nof_version(void) { return NOF_VERSION; }
extern "C" int nof_set_amp_impl(int impl) {
  const int old = nof::amp_impl();
  nof::g_amp_impl = impl < 0 ? 0 : (impl > 3 ? 3 : impl);
  return old;
}
extern "C" const char* nof_last_error(void) { return g_err; }

extern "C" int nof_device_info(int* sm_count, int* max_smem_optin) {
  static int cached_dev = -1, cached_sm = 0, cached_smem = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("nof_device_info: no CUDA device"); cudaGetLastError(); return NOF_E_LAUNCH; }
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&cached_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cached_dev = dev;
  }
  if (sm_count) *sm_count = cached_sm;
  if (max_smem_optin) *max_smem_optin = cached_smem;
  return NOF_OK;
}

extern "C" size_t nof_mlp_param_count(int E, int V) {
  int32_t o[10];
  size_t total;
  mlp_offsets(E, V, o, &total);
  return (total + 3) / 4 * 4;     // padded to 16 bytes: the kernels stage the block with one TMA bulk copy
}
extern "C" int nof_mlp_param_offsets(int E, int V, int32_t offsets_out[10]) {
  NOF_REQUIRE(offsets_out && E >= 2 && V >= 9, "nof_mlp_param_offsets: bad arguments");
  mlp_offsets(E, V, offsets_out, nullptr);
  return NOF_OK;
}

static int validate_step(const NofStep* p, const char* fn) {
  NOF_REQUIRE(p, "%s: null NofStep", fn);
  NOF_REQUIRE(p->C == 2, "%s: C=%d (only feature_grid_dim 2 is built)", fn, p->C);
  NOF_REQUIRE(p->L >= 1 && p->L <= MAX_L, "%s: L=%d out of range 1..%d", fn, p->L, MAX_L);
  NOF_REQUIRE(p->ff >= 0 && p->ff <= 8, "%s: ff=%d out of range 0..8", fn, p->ff);
  NOF_REQUIRE(p->offsets && p->mlp, "%s: null model pointer", fn);
  NOF_REQUIRE(p->amp ? p->table_f16 != nullptr : p->table_f32 != nullptr, "%s: table pointer for amp=%d is null", fn, p->amp);
  NOF_REQUIRE(p->ff == 0 || p->feat, "%s: ff>0 needs feat", fn);
  return NOF_OK;
}

extern "C" size_t nof_step_workspace_bytes(const NofStep* p) {
  if (!p) return 0;
  int sms = 148;
  nof_device_info(&sms, nullptr);
  Tiling t;
  if (p->amp) {                                             // large enough for every AMP implementation (nof_set_amp_impl may switch)
    size_t j = 0;
    int Sp, R;
    if (step_ws_tiling(p->S, &Sp, &R)) j = step_ws_jscratch(std::max(1, std::min((p->N + R - 1) / R, sms)));
    if (amp_tiling(p, sms, &t)) j = std::max(j, (size_t)t.blocks * MAX_L * 3 * (t.NW * 32) * 4);
    return j ? kWPackBytes + j + 256 : 0;
  }
  if (!f32_tiling(p, sms, &t)) return 0;
  return kWPackBytes + (size_t)t.blocks * MAX_L * 3 * 128 * 8 + 256;
}

// Per-step results start from zero (the tcgen05 path does this inside its operand-packing kernel instead).
static void zero_step_outputs(const NofStep* p, cudaStream_t st) {
  cudaMemsetAsync(p->losses, 0, 8 * sizeof(float), st);
  if (p->grad_tf) cudaMemsetAsync(p->grad_tf, 0, (size_t)p->F * 12 * sizeof(float), st);
}

extern "C" int nof_step_fused(const NofStep* p, nof_stream_t stream) {
  int rc = validate_step(p, "nof_step_fused");
  if (rc) return rc;
  NOF_REQUIRE(p->N >= 0 && p->S >= 1, "nof_step_fused: bad sizes N=%d S=%d", p->N, p->S);
  NOF_REQUIRE(p->rays && p->tf && p->z_vals, "nof_step_fused: null batch pointer");
  NOF_REQUIRE(p->grad_table && p->grad_mlp && p->losses, "nof_step_fused: null output pointer");
  NOF_REQUIRE((reinterpret_cast<uintptr_t>(p->grad_table) & 15u) == 0, "nof_step_fused: grad_table must be 16-byte aligned");
  NOF_REQUIRE(!p->need_pose_grad || p->grad_tf, "nof_step_fused: need_pose_grad without grad_tf");
  NOF_REQUIRE(p->workspace, "nof_step_fused: null workspace (see nof_step_workspace_bytes)");
  NOF_REQUIRE(p->ray_dim >= 10, "nof_step_fused: ray_dim=%d", p->ray_dim);
  NOF_REQUIRE(p->F >= 1, "nof_step_fused: F=%d", p->F);
  if (p->N == 0) return NOF_OK;
  int sms = 148, smem_max = 0;
  nof_device_info(&sms, &smem_max);
  StepArgs a;
  a.p = *p;
  a.E = p->L * p->C;
  a.V = p->ff + 9;
  a.KE = (a.E + 15) / 16 * 16;
  mlp_offsets(a.E, a.V, a.po, nullptr);
  a.inv_N3 = 1.0f / (3.0f * (float)p->N);
  a.inv_NS = 1.0f / ((float)p->N * (float)p->S);
  a.inv_NS3 = a.inv_NS / 3.0f;
  a.count_only = 0;
  a.wpack = p->workspace;                                   // [0, kWPackBytes): packed fp16 MLP operands (tcgen05 path)
  a.jws = static_cast<char*>(p->workspace) + kWPackBytes;   // then the per-CTA Jacobian scratch
  Tiling t;
  const bool eik = p->eikonal_weight > 0.f;
  NOF_REQUIRE(!eik || (p->amp && p->S <= 256), "nof_step_fused: eikonal_weight > 0 is built for amp: true and S <= 256 (S=%d amp=%d)", p->S, p->amp);
  if (eik) {
    // the term is a mean over the selected samples of the WHOLE batch: they are counted first, by a forward-only pass of the tile kernel
    int* cnt = reinterpret_cast<int*>(static_cast<char*>(a.wpack) + kWPackBytes - 32);
    cudaMemsetAsync(cnt, 0, sizeof(int), as_stream(stream));
  }
  if (p->amp && !eik && amp_impl_for(p->S) == 2) {
    int Sp, R;
    NOF_REQUIRE(step_ws_tiling(p->S, &Sp, &R), "nof_step_fused(amp): S=%d > 384 samples per ray not supported (use amp: false)", p->S);
    a.R = R; a.Sp = Sp; a.n_groups = (p->N + R - 1) / R;
    NOF_REQUIRE(step_ws_smem(a.KE) <= (size_t)smem_max, "nof_step_fused(amp, ws): needs %zu B shared memory, device allows %d", step_ws_smem(a.KE), smem_max);
    return step_ws_dispatch(a, std::max(1, std::min(a.n_groups, sms)), as_stream(stream));
  }
  if (p->amp) {
    NOF_REQUIRE(amp_tiling(p, sms, &t), "nof_step_fused(amp): S=%d > 256 samples per ray not supported by the AMP tile (use amp: false)", p->S);
    a.R = t.R; a.Sp = t.Sp; a.n_groups = t.n_groups;
    if (t.NW == 4 && !eik && amp_impl_for(p->S) == 1) {
      NOF_REQUIRE(step_tc_smem(a.KE) <= (size_t)smem_max, "nof_step_fused(amp, tcgen05): needs %zu B shared memory, device allows %d",
                  step_tc_smem(a.KE), smem_max);
      return step_tc_dispatch(a, t.blocks, as_stream(stream));
    }
    NOF_REQUIRE(step_amp_smem(t.NW * 32, a.KE, eik) <= (size_t)smem_max, "nof_step_fused(amp): needs %zu B shared memory, device allows %d",
                step_amp_smem(t.NW * 32, a.KE, eik), smem_max);
    zero_step_outputs(p, as_stream(stream));
    if (eik) {                                              // counting pass: same kernel, forward up to the sdf only
      a.count_only = 1;
      rc = step_amp_dispatch(a, t.NW, t.blocks, as_stream(stream));
      if (rc) return rc;
      a.count_only = 0;
    }
    return step_amp_dispatch(a, t.NW, t.blocks, as_stream(stream));
  }
  NOF_REQUIRE(f32_tiling(p, sms, &t), "nof_step_fused(fp32): S=%d not supported (R*ceil32(S) must be <= 1024 with R<=4)", p->S);
  a.R = t.R; a.Sp = t.Sp; a.n_groups = t.n_groups;
  NOF_REQUIRE(step_f32_smem(a.E, a.V, a.R * a.Sp) <= (size_t)smem_max, "nof_step_fused(fp32): shared memory %zu > %d",
              step_f32_smem(a.E, a.V, a.R * a.Sp), smem_max);
  zero_step_outputs(p, as_stream(stream));
  return step_f32_dispatch(a, t.blocks, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// SDF-only query (mesh extraction): encode + sigma_net, thread = point, weights broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
namespace nof {
template <bool HALF>
__global__ void __launch_bounds__(256) query_sdf_kernel(const StepArgs a, const float* __restrict__ xin, float* __restrict__ sdf, int64_t P) {
  extern __shared__ __align__(16) float sq[];
  __shared__ LevelS lv;
  const int E = a.E;
  float* sW1 = sq;                 // [64][E]
  float* sB1 = sW1 + 64 * E;       // [64]
  float* sW2 = sB1 + 64;           // [64] (row 0 of W2)
  float b2;
  for (int i = threadIdx.x; i < 64 * E; i += blockDim.x) { float w = a.p.mlp[a.po[0] + i]; sW1[i] = HALF ? __half2float(__float2half_rn(w)) : w; }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    float b = a.p.mlp[a.po[1] + i], w2 = a.p.mlp[a.po[2] + i];
    sB1[i] = HALF ? __half2float(__float2half_rn(b)) : b;
    sW2[i] = HALF ? __half2float(__float2half_rn(w2)) : w2;
  }
  b2 = a.p.mlp[a.po[3]];
  if (HALF) b2 = __half2float(__float2half_rn(b2));
  init_levels(lv, a);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float u[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = (fminf(fmaxf(xin[i * 3 + d], -1.f), 1.f) + 1.f) * 0.5f;   // nerf_runner.py:1314, grid.py:160
    float enc[32];
#pragma unroll
    for (int l = 0; l < MAX_L; ++l) {
      float e[2] = {0.f, 0.f}, J[3][2];
      if (l < a.p.L) gather_level<HALF, false>(HALF ? a.p.table_f16 : (const void*)a.p.table_f32, lv, l, u, e, J);
      enc[2 * l] = HALF ? __half2float(__float2half_rn(e[0])) : e[0];
      enc[2 * l + 1] = HALF ? __half2float(__float2half_rn(e[1])) : e[1];
    }
    float out = b2;
    for (int o = 0; o < 64; ++o) {
      float acc = sB1[o];
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < E) acc = fmaf(enc[k], sW1[o * E + k], acc);
      acc = fmaxf(acc, 0.f);
      if (HALF) acc = __half2float(__float2half_rn(acc));
      out = fmaf(acc, sW2[o], out);
    }
    sdf[i] = HALF ? __half2float(__float2half_rn(out)) : out;
  }
}
}  // namespace nof

extern "C" int nof_query_sdf(const NofStep* model, const float* x, float* sdf, int64_t P, nof_stream_t stream) {
  int rc = validate_step(model, "nof_query_sdf");
  if (rc) return rc;
  NOF_REQUIRE(x && sdf && P >= 0, "nof_query_sdf: bad arguments");
  if (P == 0) return NOF_OK;
  StepArgs a;
  a.p = *model;
  a.E = model->L * model->C;
  a.V = model->ff + 9;
  a.KE = (a.E + 15) / 16 * 16;
  mlp_offsets(a.E, a.V, a.po, nullptr);
  a.R = a.Sp = a.n_groups = 0;
  a.count_only = 0;
  a.inv_N3 = a.inv_NS = a.inv_NS3 = 0.f;
  int sms = 148;
  nof_device_info(&sms, nullptr);
  const size_t smem = (size_t)(64 * a.E + 128) * 4;
  const int blocks = (int)std::min<int64_t>((P + 255) / 256, (int64_t)sms * 8);
  if (model->amp) query_sdf_kernel<true><<<blocks, 256, smem, as_stream(stream)>>>(a, x, sdf, P);
  else query_sdf_kernel<false><<<blocks, 256, smem, as_stream(stream)>>>(a, x, sdf, P);
  return check_launch("query_sdf_kernel");
}
