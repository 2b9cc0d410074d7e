// Iso-surface extraction for NerfRunner.extract_mesh (reference nerf_runner.py:1387-1404 calls skimage.measure.marching_cubes on
// the host; SURVEY.md §8(f)-2 asks for it on the GPU). Marching TETRAHEDRA over the Kuhn split of every grid cell (6 tetrahedra
// around the main diagonal): no case tables, no ambiguous faces — neighbouring cells split their shared face along the same
// diagonal, so the surface is a closed 2-manifold wherever it does not leave the grid. Every vertex lies on one grid edge
// (cube edge, face diagonal or main diagonal) and is interpolated from the edge's lower-index end point, so the cells that
// share the edge produce bit-identical positions and the same 64-bit key (lower point index * 8 + edge direction): the host
// side welds vertices with one torch.unique over the keys. Two passes (count -> exclusive scan on the host side -> emit) keep
// the output order deterministic. Triangles are oriented towards increasing field values (outward for an SDF).
#include "nof_common.cuh"

namespace nof {

__constant__ int kCorner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
__constant__ int kTet[6][4] = {{0, 5, 1, 6}, {0, 1, 2, 6}, {0, 2, 3, 6}, {0, 3, 7, 6}, {0, 7, 4, 6}, {0, 4, 5, 6}};

__device__ __forceinline__ int tet_triangles(const float v[4], float iso, int* mask_out) {
  int mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) mask |= (v[i] < iso) ? (1 << i) : 0;
  *mask_out = mask;
  const int n_in = __popc(mask);
  return (n_in == 1 || n_in == 3) ? 1 : (n_in == 2 ? 2 : 0);
}

__global__ void mtet_count_kernel(const float* __restrict__ f, int nx, int ny, int nz, float iso, int32_t* __restrict__ counts) {
  const int64_t cells = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int k = (int)(c % (nz - 1)), j = (int)((c / (nz - 1)) % (ny - 1)), i = (int)(c / ((int64_t)(nz - 1) * (ny - 1)));
  float cv[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) cv[q] = __ldg(f + ((int64_t)(i + kCorner[q][0]) * ny + (j + kCorner[q][1])) * nz + (k + kCorner[q][2]));
  int total = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const float v[4] = {cv[kTet[t][0]], cv[kTet[t][1]], cv[kTet[t][2]], cv[kTet[t][3]]};
    int mask;
    total += tet_triangles(v, iso, &mask);
  }
  counts[c] = total;
}

struct MVert { float p[3]; int64_t key; };

// vertex on the grid edge between cube corners qa and qb of cell (i,j,k): interpolate from the lower-index end point
__device__ __forceinline__ MVert edge_vertex(int i, int j, int k, int qa, int qb, float va, float vb, float iso, int ny, int nz) {
  const int ax = i + kCorner[qa][0], ay = j + kCorner[qa][1], az = k + kCorner[qa][2];
  const int bx = i + kCorner[qb][0], by = j + kCorner[qb][1], bz = k + kCorner[qb][2];
  const int64_t ia = ((int64_t)ax * ny + ay) * nz + az, ib = ((int64_t)bx * ny + by) * nz + bz;
  const bool a_lo = ia < ib;
  const int lx = a_lo ? ax : bx, ly = a_lo ? ay : by, lz = a_lo ? az : bz;
  const int hx = a_lo ? bx : ax, hy = a_lo ? by : ay, hz = a_lo ? bz : az;
  const float vl = a_lo ? va : vb, vh = a_lo ? vb : va;
  const float t = __fdiv_rn(__fsub_rn(iso, vl), __fsub_rn(vh, vl));        // in [0,1]: vl and vh are on different sides of iso
  MVert r;
  r.p[0] = __fmaf_rn(t, (float)(hx - lx), (float)lx);
  r.p[1] = __fmaf_rn(t, (float)(hy - ly), (float)ly);
  r.p[2] = __fmaf_rn(t, (float)(hz - lz), (float)lz);
  r.key = (a_lo ? ia : ib) * 8 + ((hx - lx) * 4 + (hy - ly) * 2 + (hz - lz));
  return r;
}

__device__ __forceinline__ void put_triangle(float* __restrict__ verts, int64_t* __restrict__ keys, int64_t tri, const MVert& a, const MVert& b,
                                             const MVert& c, bool flip) {
  const MVert& p1 = flip ? c : b;
  const MVert& p2 = flip ? b : c;
  float* o = verts + tri * 9;
  o[0] = a.p[0]; o[1] = a.p[1]; o[2] = a.p[2];
  o[3] = p1.p[0]; o[4] = p1.p[1]; o[5] = p1.p[2];
  o[6] = p2.p[0]; o[7] = p2.p[1]; o[8] = p2.p[2];
  keys[tri * 3] = a.key; keys[tri * 3 + 1] = p1.key; keys[tri * 3 + 2] = p2.key;
}

__global__ void mtet_emit_kernel(const float* __restrict__ f, int nx, int ny, int nz, float iso, const int64_t* __restrict__ offsets,
                                 float* __restrict__ verts, int64_t* __restrict__ keys) {
  const int64_t cells = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int k = (int)(c % (nz - 1)), j = (int)((c / (nz - 1)) % (ny - 1)), i = (int)(c / ((int64_t)(nz - 1) * (ny - 1)));
  float cv[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) cv[q] = __ldg(f + ((int64_t)(i + kCorner[q][0]) * ny + (j + kCorner[q][1])) * nz + (k + kCorner[q][2]));
  int64_t tri = offsets[c];
  for (int t = 0; t < 6; ++t) {
    int q[4];
    float v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) { q[m] = kTet[t][m]; v[m] = cv[q[m]]; }
    int mask;
    const int nt = tet_triangles(v, iso, &mask);
    if (nt == 0) continue;
    // inside / outside vertex lists (positions inside the tet, 0..3)
    int in[3], out[3], ni = 0, no = 0;
    for (int m = 0; m < 4; ++m) {
      const bool inside = (mask >> m) & 1;
      if (inside) in[ni++] = m; else out[no++] = m;
    }
    // Winding from the tetrahedron's orientation, not from the triangle's geometry (a triangle whose vertices coincide with grid
    // points has no normal): with the vertex list L = (a | b c d) used below and D = det(b-a, c-a, d-a) over the integer corner
    // coordinates, the triangle (P_ab, P_ac, P_ad) has its normal towards the far face iff D > 0 (N.(u+v+w) = D (tc td + tb tc + tb td)),
    // and the quad (P_ac, P_ad, P_bd, P_bc) of the 2+2 case has N.(inside -> outside) = D/4 with L = (a b | c d).
    int L4[4];
    if (ni == 1) { L4[0] = in[0]; L4[1] = out[0]; L4[2] = out[1]; L4[3] = out[2]; }
    else if (ni == 3) { L4[0] = out[0]; L4[1] = in[0]; L4[2] = in[1]; L4[3] = in[2]; }
    else { L4[0] = in[0]; L4[1] = in[1]; L4[2] = out[0]; L4[3] = out[1]; }
    int e[3][3];
    for (int r = 0; r < 3; ++r)
      for (int d = 0; d < 3; ++d) e[r][d] = kCorner[q[L4[r + 1]]][d] - kCorner[q[L4[0]]][d];
    const int D = e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1]) - e[0][1] * (e[1][0] * e[2][2] - e[1][2] * e[2][0]) +
                  e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0]);
    const bool flip = (ni == 3) ? (D > 0) : (D < 0);
    auto ev = [&](int a, int b) { return edge_vertex(i, j, k, q[a], q[b], v[a], v[b], iso, ny, nz); };
    if (ni == 1) {
      put_triangle(verts, keys, tri++, ev(in[0], out[0]), ev(in[0], out[1]), ev(in[0], out[2]), flip);
    } else if (ni == 3) {
      put_triangle(verts, keys, tri++, ev(out[0], in[0]), ev(out[0], in[1]), ev(out[0], in[2]), flip);
    } else {                                                // 2 + 2: a quad, split along the (in0,out0)-(in1,out1) diagonal
      const MVert p00 = ev(in[0], out[0]), p01 = ev(in[0], out[1]), p11 = ev(in[1], out[1]), p10 = ev(in[1], out[0]);
      put_triangle(verts, keys, tri++, p00, p01, p11, flip);
      put_triangle(verts, keys, tri++, p00, p11, p10, flip);
    }
  }
}

// Nearest-neighbour distance test against a point cloud binned into a uniform grid of cell size >= radius (cloud points sorted by
// cell, cell_start = exclusive prefix of the per-cell counts): a query is within `radius` of the cloud iff one of the 27 cells around
// it holds such a point. Replaces the host cKDTree query of the ray-pool denoise (nerf_runner.py:178-195). Thread = query.
__global__ void __launch_bounds__(256) cloud_within_kernel(const float* __restrict__ query, int64_t Q, const float* __restrict__ cloud,
                                                           const int32_t* __restrict__ cell_start, float lo, float inv_cell, int n, float r2,
                                                           uint8_t* __restrict__ within) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  const float qx = query[i * 3 + 0], qy = query[i * 3 + 1], qz = query[i * 3 + 2];
  const int cx = (int)floorf((qx - lo) * inv_cell), cy = (int)floorf((qy - lo) * inv_cell), cz = (int)floorf((qz - lo) * inv_cell);
  bool hit = false;
  for (int dx = -1; dx <= 1 && !hit; ++dx) {
    const int x = cx + dx;
    if (x < 0 || x >= n) continue;
    for (int dy = -1; dy <= 1 && !hit; ++dy) {
      const int y = cy + dy;
      if (y < 0 || y >= n) continue;
      const int z0 = max(cz - 1, 0), z1 = min(cz + 1, n - 1);          // cells along z are contiguous in the sorted cloud
      if (z0 > z1) continue;
      const int c0 = (x * n + y) * n + z0, c1 = (x * n + y) * n + z1;
      for (int k = cell_start[c0]; k < cell_start[c1 + 1]; ++k) {
        const float ex = cloud[k * 3 + 0] - qx, ey = cloud[k * 3 + 1] - qy, ez = cloud[k * 3 + 2] - qz;
        if (ex * ex + ey * ey + ez * ez <= r2) { hit = true; break; }
      }
    }
  }
  within[i] = hit ? 1 : 0;
}

}  // namespace nof

using namespace nof;

extern "C" int nof_cloud_within_radius(const float* query, int64_t Q, const float* cloud_sorted, const int32_t* cell_start, float lo, float cell,
                                       int n, float radius, uint8_t* within, nof_stream_t stream) {
  NOF_REQUIRE(query && cloud_sorted && cell_start && within, "nof_cloud_within_radius: null pointer");
  NOF_REQUIRE(Q >= 0 && n >= 1 && n <= 1024 && cell >= radius && radius > 0.f, "nof_cloud_within_radius: bad grid (n=%d cell=%g radius=%g)", n, cell, radius);
  if (Q == 0) return NOF_OK;
  cloud_within_kernel<<<(unsigned)div_up<int64_t>(Q, 256), 256, 0, as_stream(stream)>>>(query, Q, cloud_sorted, cell_start, lo, 1.0f / cell, n,
                                                                                          radius * radius, within);
  return check_launch("cloud_within_kernel");
}

extern "C" int nof_marching_tets_count(const float* field, int nx, int ny, int nz, float iso, int32_t* counts, nof_stream_t stream) {
  NOF_REQUIRE(field && counts, "nof_marching_tets_count: null pointer");
  NOF_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "nof_marching_tets_count: grid %dx%dx%d too small", nx, ny, nz);
  const int64_t cells = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
  NOF_REQUIRE(cells <= (int64_t)0x7fffffff * 256, "nof_marching_tets_count: grid too large");
  mtet_count_kernel<<<(unsigned)div_up<int64_t>(cells, 256), 256, 0, as_stream(stream)>>>(field, nx, ny, nz, iso, counts);
  return check_launch("mtet_count_kernel");
}

extern "C" int nof_marching_tets_emit(const float* field, int nx, int ny, int nz, float iso, const int64_t* offsets, float* verts,
                                      int64_t* keys, nof_stream_t stream) {
  NOF_REQUIRE(field && offsets && verts && keys, "nof_marching_tets_emit: null pointer");
  NOF_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "nof_marching_tets_emit: grid %dx%dx%d too small", nx, ny, nz);
  const int64_t cells = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
  mtet_emit_kernel<<<(unsigned)div_up<int64_t>(cells, 256), 256, 0, as_stream(stream)>>>(field, nx, ny, nz, iso, offsets, verts, keys);
  return check_launch("mtet_emit_kernel");
}
