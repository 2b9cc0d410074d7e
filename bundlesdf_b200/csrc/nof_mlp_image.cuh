// The MLP operand image shared by the tcgen05 step kernels (nof_step_tc.cu, nof_step_ws.cu): what every CTA fetches with ONE TMA bulk copy.
//   [W1 | W2 | W3[:, geo] | W4 | W5]   fp16, canonical core-matrix order (nof_tc_prims.cuh cm_off), zero padded
//   biases                               fp32 values rounded through fp16 (autocast casts them): b1 64 | b2 16 | b3 64 | b4 64 | b5 8 | 64 zeros
//   W3[:, views]                         fp16 row-major [64][VPAD]: the colour net's first layer acts on the per-ray inputs (SH of the view
//                                        direction, frame feature) ONCE per ray; its result enters the tile GEMM as a per-ray bias
//   level geometry                       LevelS (scale, resolution, table offset ... per hash-grid level; device exp2f, gridencoder.cu:155)
// One tiny kernel per step builds it from the fp32 parameter block and also zeroes the per-step outputs and the tile ticket.
#pragma once
#include "nof_step_common.cuh"
#include "nof_tc_prims.cuh"

namespace nof {
namespace img {
using namespace prim;

constexpr int KG = 16;                        // colour-net GEMM input: geo(15) + pad
constexpr int VPAD = 18;                      // row length of the fp16 W3[:, views] block (V <= 17)

struct ImagePlan { int w1, w2, w3g, w4, w5, bias, w3v, lv, bytes; };
__host__ __device__ inline ImagePlan make_image_plan(int KE) {
  ImagePlan s;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) / 128 * 128; return r; };
  s.w1 = take(64 * KE * 2);
  s.w2 = take(16 * 64 * 2);
  s.w3g = take(64 * KG * 2);
  s.w4 = take(64 * 64 * 2);
  s.w5 = take(16 * 64 * 2);
  s.bias = take(280 * 4);
  s.w3v = take(64 * VPAD * 2);
  s.lv = take((int)sizeof(LevelS));
  s.bytes = o;
  return s;
}

// fp32 packed parameters -> [W1 | W2 | W3[:, geo] | W4 | W5 (core-matrix fp16, zero padded) | biases (fp32, fp16-rounded) |
// W3[:, views] fp16 row-major [64][VPAD] | level geometry]. One small kernel per step; also zeroes the per-step outputs and the tile ticket.
template <int KE>
__global__ void __launch_bounds__(256) pack_image_kernel(const StepArgs a) {
  const ImagePlan sp = make_image_plan(KE);
  unsigned char* out = static_cast<unsigned char*>(a.wpack);
  const int gid = blockIdx.x * 256 + threadIdx.x;
  if (gid == 0) *reinterpret_cast<int*>(out + kWPackBytes - 16) = 0;                 // tile ticket of the step kernel
  if (gid < 8) a.p.losses[gid] = 0.f;                                                // per-step results start from zero
  if (a.p.grad_tf) for (int i = gid; i < a.p.F * 12; i += gridDim.x * 256) a.p.grad_tf[i] = 0.f;
  const float* P = a.p.mlp;
  const int V = a.V, K3 = V + 15;
  const int o = gid * 4;                                          // one 32-bit word of the image per thread
  if (o < sp.bias) {
    int base, rows, kreal, kpad, po, kofs = 0, ld;
    if (o >= sp.w5) { base = sp.w5; rows = 3; kreal = 64; kpad = 64; po = a.po[8]; ld = 64; }
    else if (o >= sp.w4) { base = sp.w4; rows = 64; kreal = 64; kpad = 64; po = a.po[6]; ld = 64; }
    else if (o >= sp.w3g) { base = sp.w3g; rows = 64; kreal = 15; kpad = KG; po = a.po[4]; kofs = V; ld = K3; }
    else if (o >= sp.w2) { base = sp.w2; rows = 16; kreal = 64; kpad = 64; po = a.po[2]; ld = 64; }
    else { base = sp.w1; rows = 64; kreal = a.E; kpad = KE; po = a.po[0]; ld = a.E; }
    const int rel = o - base, rg = rel / (kpad * 16), rem = rel % (kpad * 16);
    const int n = rg * 8 + (rem % 128) / 16, k = (rem / 128) * 8 + (rem % 16) / 2;          // inverse of cm_off
    const float v0 = (n < rows && k < kreal) ? P[po + n * ld + kofs + k] : 0.f;
    const float v1 = (n < rows && k + 1 < kreal) ? P[po + n * ld + kofs + k + 1] : 0.f;
    *reinterpret_cast<uint32_t*>(out + o) = pack_h2(v0, v1);
  } else if (o < sp.w3v) {
    const int j = (o - sp.bias) / 4;
    float v = 0.f;
    if (j < 64) v = P[a.po[1] + j];
    else if (j < 80) v = P[a.po[3] + j - 64];
    else if (j < 144) v = P[a.po[5] + j - 80];
    else if (j < 208) v = P[a.po[7] + j - 144];
    else if (j < 211) v = P[a.po[9] + j - 208];
    *reinterpret_cast<float*>(out + o) = __half2float(__float2half_rn(v));
  } else if (o < sp.lv) {
    const int e = (o - sp.w3v) / 2, n = e / VPAD, k = e % VPAD;
    const float v0 = (n < 64 && k < V) ? P[a.po[4] + n * K3 + k] : 0.f;
    const float v1 = (n < 64 && k + 1 < V) ? P[a.po[4] + n * K3 + k + 1] : 0.f;
    *reinterpret_cast<uint32_t*>(out + o) = pack_h2(v0, v1);
  } else if (o < sp.bytes) {
    const int l = (o - sp.lv) / 4;                               // thread l (< MAX_L) fills column l of every LevelS array
    if (l < MAX_L) {
      LevelS* lv = reinterpret_cast<LevelS*>(out + sp.lv);
      if (l < a.p.L) {
        LevelGeom g = level_geom3(l, a.p.S_log2, a.p.H, a.p.offsets);
        lv->scale[l] = g.scale; lv->res1[l] = g.resolution + 1u; lv->hsize[l] = g.hashmap_size; lv->off[l] = g.offset; lv->dense[l] = g.dense;
        lv->hmask[l] = (g.hashmap_size & (g.hashmap_size - 1)) == 0 ? g.hashmap_size - 1 : 0u;
      } else {
        lv->scale[l] = 0.f; lv->res1[l] = 1u; lv->hsize[l] = 1u; lv->off[l] = 0u; lv->dense[l] = 1u; lv->hmask[l] = 0u;
      }
    }
  }
}

}  // namespace img
}  // namespace nof
