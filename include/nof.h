/*
 * nof.h — C ABI of libnof_sm100.so: the B200-native (sm_100a) Neural-Object-Field training hot path.
 *
 * This is the drop-in boundary for BundleSDF's NeRF training step. Every entry point replaces a native
 * interface (or a run of PyTorch ops) of the reference; the reference file:line each one replaces is cited.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers + sizes, NO torch / ATen types.
 *   - every data pointer is a DEVICE pointer owned by the caller; nothing is allocated or freed inside.
 *   - asynchronous on `stream` (a cudaStream_t passed as void*; pass the framework's CURRENT stream —
 *     the reference launches on the legacy default stream, common.cu:121,163 / gridencoder.cu:373).
 *   - no host synchronisation inside, thread-safe per stream, device = the caller's current device.
 *   - return 0 on success, a negative NOF_E_* code otherwise; nof_last_error() gives the message
 *     (thread-local). Launch failures are reported via cudaPeekAtLastError.
 *   - kernels never spin or trap on bad data: they clamp and set a device-side error flag (int32) the
 *     caller may poll (the reference's sampler prints and spins forever, common.cu:66-71,87-92).
 */
#ifndef NOF_H_
#define NOF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NOF_VERSION 100

enum {
  NOF_OK = 0,
  NOF_E_INVALID = -1,     /* bad argument (null pointer, unsupported size / template) */
  NOF_E_LAUNCH = -2,      /* CUDA launch / runtime error */
  NOF_E_UNSUPPORTED = -3  /* valid in the reference but not built here (e.g. D=5) */
};

enum { NOF_F32 = 0, NOF_F16 = 1 };

typedef void* nof_stream_t; /* cudaStream_t */

int nof_version(void);
const char* nof_last_error(void);
/* Number of SMs / max opt-in shared memory of the current device (for grid sizing by the host). */
int nof_device_info(int* sm_count, int* max_smem_optin);

/* =====================================================================================================
 * 1. Op-level entry points: 1:1 with the reference's pybind modules `gridencoder` and `common`.
 * ===================================================================================================== */

/* Replaces gridencoder.grid_encode_forward  (mycuda/torch_ngp_grid_encoder/gridencoder.h:23,
 * gridencoder.cu:447-470; kernel_grid :107-246).
 *   inputs      [B,D] fp32 in [0,1]
 *   embeddings  [sO,C] dtype (NOF_F32 | NOF_F16)
 *   offsets     [L+1] int32 (entries)
 *   outputs     [L,B,C] dtype          (level-major, exactly like the reference; the caller permutes)
 *   dy_dx       [B,L,D,C] dtype, written iff calc_grad_inputs
 *   S = log2(per_level_scale) as fp32, H = base resolution, gridtype 0=hash 1=tiled.
 * D in {2,3}, C in {1,2,4,8} are built (the hot path uses D=3,C=2); others -> NOF_E_UNSUPPORTED. */
int nof_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                            int calc_grad_inputs, void* dy_dx, uint32_t gridtype, int align_corners,
                            int dtype, nof_stream_t stream);

/* Replaces gridencoder.grid_encode_backward (gridencoder.h:24, gridencoder.cu:472-502; kernels :250-365).
 *   grad            [L,B,C] dtype
 *   grad_embeddings [sO,C] dtype, ACCUMULATED into (caller zero-fills, grid.py:86)
 *   grad_inputs     [B,D] dtype, written iff calc_grad_inputs (from dy_dx) */
int nof_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                             void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                             uint32_t H, int calc_grad_inputs, const void* dy_dx, void* grad_inputs,
                             uint32_t gridtype, int align_corners, int dtype, nof_stream_t stream);

/* Per-level geometry exactly as the device evaluates it (gridencoder.cu:155: scale = exp2f(level*S)*H - 1). CUDA's exp2f
 * differs from a host libm by an ulp at some levels, so test oracles take the scales from here. scales_out: device [L]. */
int nof_grid_level_scales(float S, uint32_t H, int L, float* scales_out, nof_stream_t stream);

/* Replaces common.sampleRaysUniformOccupiedVoxels (mycuda/common.h:28, common.cu:41-125).
 *   z_in_out [N,I,2], z_sampled [N,S], z_vals [N,S] (written where the ray has intervals, else untouched).
 *   err_flag: optional device int32, set to 1 where the reference would print+spin (we clamp). */
int nof_sample_rays_uniform_occupied_voxels(const float* z_in_out, const float* z_sampled, float* z_vals,
                                            int N, int I, int S, int32_t* err_flag, nof_stream_t stream);

/* Replaces common.postprocessOctreeRayTracing (mycuda/common.h:29, common.cu:129-167).
 *   ray_index [M] int64, depth_in_out [M,2] fp32, unique_ids [U] int64, start_poss [U] int64
 *   out [N_rays, max_intersections, 2] fp32: zero-filled INSIDE (the reference allocates at::zeros),
 *   on the CURRENT device (the reference hard-codes cuda:0, common.cu:158). */
int nof_postprocess_octree_ray_tracing(const int64_t* ray_index, const float* depth_in_out,
                                       const int64_t* unique_ids, const int64_t* start_poss, int M, int U,
                                       int max_intersections, int N_rays, float* out, nof_stream_t stream);

/* =====================================================================================================
 * 2. Fused hot path. One train step = nof_gather_rays -> nof_pose_forward -> nof_ray_march ->
 *    nof_step_fused (forward + losses + backward in ONE kernel) -> nof_pose_backward -> nof_adam_step.
 * ===================================================================================================== */

/* Replaces DataLoader.__next__'s index gather rays[ids] (nerf_runner.py:97-107).
 *   pool [R,ray_dim] fp32, ids [N] int64 (a slice of the epoch permutation), batch [N,ray_dim] fp32. */
int nof_gather_rays(const float* pool, const int64_t* ids, float* batch, int N, int ray_dim, nof_stream_t stream);

/* Replaces PoseArray.get_matrices + `tf = dT @ c2w` (nerf_helpers.py:143-154, nerf_runner.py:1051-1053),
 * for ALL frames at once. pose_data [F,6] (NULL -> tf = c2w), c2w [F,4,4] fp32 row-major,
 * tf [F,12] = rows 0..2 of the 4x4. max_trans already multiplied by sc_factor (nerf_runner.py:240). */
int nof_pose_forward(const float* pose_data, const float* c2w, float* tf, int F, float max_trans,
                     float max_rot_deg, nof_stream_t stream);
/* Backward of the above: grad_tf [F,12] -> grad_pose [F,6] (ACCUMULATED into; row 0 gets 0:
 * frame 0 is forced to identity, nerf_helpers.py:151-153). grad_scale multiplies the result (1/loss_scale). */
int nof_pose_backward(const float* pose_data, const float* c2w, const float* grad_tf, float* grad_pose, int F,
                      float max_trans, float max_rot_deg, const float* loss_scale_or_null, nof_stream_t stream);

/* Everything of a train step that precedes the ray march, in one launch: DataLoader.__next__ (nerf_runner.py:97-107) at a
 * DEVICE-resident cursor + PoseArray.get_matrices / `dT @ c2w` for all frames (see nof_gather_rays / nof_pose_forward) + the
 * counters. With the cursor in device memory a captured CUDA graph can hold many consecutive steps (NerfRunner.train_steps). */
typedef struct {
  const float* pool;        /* [R,ray_dim] ray pool, or NULL: no gather (the caller filled `batch`) */
  const int64_t* ids;       /* [n_ids] epoch permutation */
  int64_t n_ids;
  float* batch;             /* [N,ray_dim] out */
  int N, ray_dim;
  int64_t* cursor;          /* device scalar: index into `ids` of this step's first ray; advanced by N (the host keeps whole
                               batches inside an epoch and rewrites the cursor when it reshuffles) */
  const float* pose_data;   /* [F,6] or NULL */
  const float* c2w;         /* [F,4,4] */
  float* tf;                /* [F,12] out */
  int F;
  float max_trans, max_rot_deg;
  uint64_t* tick;           /* optional device counter, +1 per call (NofMarchCfg.offset_ptr: the sampler's RNG stream position) */
  int32_t* done;            /* device int32, zero-initialised: CTA completion ticket (required when cursor or tick is given) */
  /* truncation schedule on the device (cfg trunc_decay_type linear|exp, nerf_runner.py:663-676) so that the annealed truncation does not
   * force one launch sequence per step: trunc_out[0] = trunc_table[min(*gstep, trunc_len-1)], then *gstep += 1 (by the last block).
   * trunc_out is what NofMarchCfg.trunc_ptr / NofStep.trunc_ptr of the same step point to. All four NULL/0: constant truncation. */
  const float* trunc_table; /* [trunc_len] device floats: truncation (normalised units) of step 0, 1, ... */
  int32_t trunc_len;
  int64_t* gstep;           /* device int64: the step index this launch belongs to */
  float* trunc_out;         /* device float */
} NofPrologue;
int nof_step_prologue(const NofPrologue* p, nof_stream_t stream);

typedef struct {
  int N;                /* rays in the batch */
  int ray_dim;          /* floats per ray row: 12 = dir3 rgb3 depth mask frame type near far (nerf_runner.py:259-300) */
  int S_occ;            /* cfg N_samples: stratified over the occupied-voxel length (nerf_runner.py:979-1011) */
  int S_depth;          /* cfg N_samples_around_depth (nerf_runner.py:1063-1081) */
  int level;            /* occupancy grid is (2^level)^3 over [-1,1]^3 (nerf_runner.py:1058-1059) */
  int I_max;            /* capacity of the per-ray interval list (<= 3*2^level) */
  float trunc;          /* get_truncation() (nerf_runner.py:663-676), already * sc_factor */
  float near_sc, far_sc;/* cfg near/far * sc_factor */
  float neg_trunc_ratio;
  int perturb;          /* 1: stratified jitter */
  uint64_t seed, offset;/* Philox counter RNG (used when t_rand == NULL && perturb) */
  const uint64_t* offset_ptr; /* optional DEVICE counter added to `offset` (keeps the launch arguments static under CUDA graphs) */
  const float* trunc_ptr;     /* optional DEVICE scalar overriding `trunc` (NofPrologue.trunc_out) */
} NofMarchCfg;

/* Replaces OctreeManager.ray_trace (Utils.py:443-475: kaolin unbatched_raytrace + unique_consecutive +
 * common.postprocessOctreeRayTracing) AND sample_rays_uniform_occupied_voxels (nerf_runner.py:979-1011 +
 * common.sampleRaysUniformOccupiedVoxels) AND the around-depth sampler (nerf_runner.py:1063-1081).
 *   rays     [N,ray_dim] batch rows
 *   tf       [F,12] per-frame transforms (nof_pose_forward)
 *   occ_bits dense occupancy bitmask, bit index = (ix*n + iy)*n + iz, n = 2^level, 32 cells per word
 *   t_rand   [N, S_occ+S_depth] uniform randoms in [0,1) or NULL (Philox inside)
 *   z_vals   [N, S_occ+S_depth]  (out; NOT sorted, like the reference :1080)
 *   intervals_out  optional [N,I_max,2] travel-time intervals (= ray_trace's ray_depths_in_out), or NULL
 *   err_flag optional device int32 */
int nof_ray_march(const NofMarchCfg* cfg, const float* rays, const float* tf, const uint32_t* occ_bits,
                  const float* t_rand, float* z_vals, float* intervals_out, int32_t* err_flag,
                  nof_stream_t stream);

/* Packed MLP parameter block (fp32), NeRFSmall of nerf_runner.py:221 / nerf_helpers.py:243-321:
 *   W1[64,E] b1[64] W2[16,64] b2[16] W3[64,V+15] b3[64] W4[64,64] b4[64] W5[3,64] b5[3]
 * with E = L*C, V = ff + 9, each block row-major [out,in] (nn.Linear layout), packed back to back. */
size_t nof_mlp_param_count(int E, int V);
/* element offsets of the 10 blocks inside the packed buffer */
int nof_mlp_param_offsets(int E, int V, int32_t offsets_out[10]);

typedef struct {
  /* sizes */
  int N, S;             /* rays, samples per ray (S_occ + S_depth) */
  int L, C;             /* grid levels, features per level (C must be 2) */
  int F, ff;            /* frames, per-frame feature channels (FeatureArray, nerf_helpers.py:108-124) */
  int ray_dim;
  int amp;              /* 1: fp16 table + fp16 tensor-core MLP (cfg amp: true); 0: fp32 everywhere */
  /* grid (grid.py:107-148) */
  float S_log2; int H;  /* log2(per_level_scale), base_res */
  const int32_t* offsets;       /* [L+1] */
  const float* table_f32;       /* [sO,C] master table (used when amp==0) */
  const void* table_f16;        /* [sO,C] fp16 shadow (used when amp==1; nof_adam_step refreshes it) */
  /* model */
  const float* mlp;             /* packed block, see above */
  const float* feat;            /* [F,ff] or NULL */
  /* batch */
  const float* rays;            /* [N,ray_dim] */
  const float* tf;              /* [F,12] */
  const float* z_vals;          /* [N,S] */
  /* loss configuration (config.yml:58-93; nerf_runner.py:679-752) */
  float trunc, near_sc, far_sc, sdf_lambda, neg_trunc_ratio;
  float rgb_weight, fs_weight, empty_weight, trunc_weight, fs_sdf, fs_rgb_weight, first_frame_weight;
  const float* loss_scale;      /* device scalar (GradScaler, nerf_runner.py:159) or NULL = 1 */
  int need_pose_grad;           /* cfg optimize_poses */
  /* outputs. The parameter gradients (grad_table, grad_mlp, grad_feat) are ACCUMULATED into with atomics: the caller zero-fills
   * them once and nof_adam_step re-zeros them. grad_tf and losses are per-step results: the call zeroes them itself. */
  float* grad_table;            /* [sO,C] fp32, scaled by loss_scale */
  float* grad_mlp;              /* packed like `mlp`, scaled by loss_scale */
  float* grad_tf;               /* [F,12], scaled */
  float* grad_feat;             /* [F,ff] or NULL, scaled (reg term added by the host side) */
  float* losses;                /* [8]: loss, rgb, fs, sdf, fs_rgb, n_valid_samples, n_valid_rays, eikonal (unscaled) */
  int32_t* found_inf;           /* device flag, set when an fp16 conversion overflowed (amp) */
  /* optional debug / parity taps (NULL to skip) */
  float* rgb_map;               /* [N,3] */
  float* raw;                   /* [N,S,4] rgb logits + sdf */
  uint8_t* valid_samples;       /* [N,S] */
  float* weights;               /* [N,S] compositing weights */
  /* workspace from nof_step_workspace_bytes */
  void* workspace;
  /* eikonal regulariser (cfg eikonal_weight; nerf_runner.py:734-738 with the normals of :1342-1345 — the reference's own train_loop
   * cannot run it: it renders with get_normals=False, :686): eikonal_weight * mean over {sdf < 1} of (|d sdf / d x| - 1)^2, x detached.
   * > 0 needs amp == 1 and S <= 256 (built in the mma.sync tile kernel); the value of the term lands in losses[7]. */
  float eikonal_weight;
  const float* trunc_ptr;       /* optional DEVICE scalar overriding `trunc` (NofPrologue.trunc_out): annealed truncation under CUDA graphs */
} NofStep;

size_t nof_step_workspace_bytes(const NofStep* p);
/* Replaces render_rays' network part + raw2outputs + the loss assembly + loss.backward()
 * (nerf_runner.py:1083-1088, 1227-1304, 1132-1169, 679-758; grid.py:34-99; nerf_helpers.py:305-321,367-399).
 * Forward + backward are ONE kernel launch; no [P,*] intermediate is written to HBM. */
int nof_step_fused(const NofStep* p, nof_stream_t stream);
/* The AMP step has three interchangeable implementations of the same arithmetic: 1 = tcgen05/TMEM tile kernel (S <= 128), 0 = warp-level
 * mma.sync tiles (S <= 256), 2 = warp-specialised streaming pipeline on tcgen05 whose rays may span tiles (S <= 384: the reference's
 * config.yml 128+64 and run_custom.py 64+256 sample counts); 3 = automatic (default): the fastest measured for the given S. Same results
 * within fp16 rounding; forcing one (env NOF_AMP_IMPL=mma|tc|ws at load) lets them be cross-checked. A forced implementation that cannot
 * carry S falls through to the next one that can. Returns the old value. */
int nof_set_amp_impl(int impl);

typedef struct {
  float* param; float* grad; float* exp_avg; float* exp_avg_sq;
  void* shadow_f16;     /* optional fp16 copy refreshed in the same pass (grid.py:50-51 cast), or NULL */
  size_t n;
  float lr;
  const float* lr_ptr;  /* optional DEVICE scalar overriding `lr` (lr schedule without re-capturing a CUDA graph) */
} NofAdamSeg;

/* Replaces optimizer.zero_grad() + GradScaler.unscale/inf-check/step/update + torch.optim.Adam.step
 * (nerf_runner.py:492-504, 756-761): exact dense Adam (betas, eps, no weight decay), single pass:
 * read g,m,v,p -> write p,m,v,(fp16 shadow) and zero g. Skips the update when *found_inf != 0.
 *   step: device int32[8]: [0] = number of Adam updates applied so far (incremented inside unless the step is skipped);
 *         [1..7] = owned by the library (fp32 bias corrections cached for update [3], CTA completion counter [4],
 *         [5] = number of skipped updates so far, [7] = sticky error flag: set when found_inf carried the value 2, i.e. the
 *         step kernel reported invalid results (a bounded tensor-core wait timed out) rather than an fp16 overflow);
 *         zero-initialise all eight. NULL = no step counter (bias corrections recomputed, bookkeeping in a second launch).
 *   scale_state: device float[2] = {loss_scale, growth_tracker} updated like GradScaler (init 65536,
 *                x2 every 2000 clean steps, x0.5 on inf) or NULL when amp is off.
 *   tick: optional device uint64 incremented on EVERY call (feeds NofMarchCfg.offset_ptr). */
int nof_adam_step(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, int32_t* step,
                  float* scale_state, int32_t* found_inf, uint64_t* tick, nof_stream_t stream);
/* The same pass split in two, for callers that update groups of segments in separate launches (e.g. the big table on a side
 * stream, overlapped with the next step's ray marching — NerfRunner cfg 'defer_table_update'): nof_adam_update applies the update
 * to `segs` reading step / scale_state / found_inf without modifying them; nof_adam_finish does the bookkeeping of
 * nof_adam_step (step count, cached bias corrections, GradScaler growth/backoff, found_inf reset, tick) once all updates of the
 * step have been issued. nof_adam_step == nof_adam_update on all segments + nof_adam_finish, in one launch. */
int nof_adam_update(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, const int32_t* step,
                    const float* scale_state, const int32_t* found_inf, nof_stream_t stream);
int nof_adam_finish(int32_t* step, float* scale_state, int32_t* found_inf, uint64_t* tick, float beta1, float beta2,
                    nof_stream_t stream);
/* One optimizer step spread over SEVERAL launches that may run on different streams in any order (e.g. the table segment next to the
 * pose backward, the small segments after it): each launch updates its segments like nof_adam_update, and the launch whose last block
 * retires last does the bookkeeping of nof_adam_finish — no extra launch on the step's critical path. total_tiles = the sum of
 * nof_adam_tile_count() over all launches of the step; every one of them passes the same value, `step` (not NULL) holds the counter. */
int nof_adam_tile_count(const NofAdamSeg* segs, int n_segs);
int nof_adam_update_shared(const NofAdamSeg* segs, int n_segs, float beta1, float beta2, float eps, int32_t* step,
                           float* scale_state, int32_t* found_inf, uint64_t* tick, int total_tiles, nof_stream_t stream);

/* SDF-only inference for mesh extraction (run_network_density, nerf_runner.py:1307-1347 with
 * NeRFSmall.forward_sdf nerf_helpers.py:296-302): x [P,3] in [-1,1] (clipped inside) -> sdf [P]. */
int nof_query_sdf(const NofStep* model, const float* x, float* sdf, int64_t P, nof_stream_t stream);

/* Replaces the host cKDTree nearest-neighbour query of the ray-pool denoise (nerf_runner.py:178-195: rays whose back-projected point is
 * farther than 2 cm from the octree cloud are marked uncertain): within[i] = 1 iff some cloud point lies within `radius` of query[i].
 *   cloud_sorted [M,3] the cloud sorted by grid cell (cell of p = floor((p - lo) / cell) per axis, linear index (x*n + y)*n + z),
 *   cell_start [n^3 + 1] int32 exclusive prefix sum of the points per cell; cell >= radius (so the 27 cells around a query suffice). */
int nof_cloud_within_radius(const float* query, int64_t Q, const float* cloud_sorted, const int32_t* cell_start, float lo, float cell,
                            int n, float radius, uint8_t* within, nof_stream_t stream);

/* Iso-surface of a dense scalar grid for NerfRunner.extract_mesh (nerf_runner.py:1387-1404 calls skimage.measure.marching_cubes
 * on the host). Marching tetrahedra over the Kuhn split of each cell (closed 2-manifold, no case tables); two passes:
 *   count: field [nx,ny,nz] (C order) -> counts [(nx-1)(ny-1)(nz-1)] triangles per cell (int32);
 *   emit:  offsets = exclusive prefix sum of counts (int64) -> verts [T,3,3] (grid-index coordinates, fp32) and
 *          keys [T,3] (int64: id of the grid edge the vertex lies on = lower end-point index * 8 + direction code); equal keys
 *          carry bit-identical positions, so vertices are welded by `unique(keys)`. Triangles face increasing field values. */
int nof_marching_tets_count(const float* field, int nx, int ny, int nz, float iso, int32_t* counts, nof_stream_t stream);
int nof_marching_tets_emit(const float* field, int nx, int ny, int nz, float iso, const int64_t* offsets, float* verts,
                           int64_t* keys, nof_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NOF_H_ */
