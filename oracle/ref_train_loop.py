"""TEST INFRASTRUCTURE (oracle/): the reference's OWN `NerfRunner.train_loop` (nerf_runner.py:679-852) executed verbatim on a B200.

SURVEY.md 8(c) "strongest available GPU oracle" / 8(d) "second comparison column": the reference's unmodified Python
(oracle/_ref/py/: nerf_runner.py, nerf_helpers.py, Utils.py, mycuda/torch_ngp_grid_encoder/grid.py, staged there by oracle/build_ref.py
because /root/reference does not exist on the GPU box) runs on top of the reference's own CUDA extensions compiled from its sources
(oracle/_ref/gridencoder/gridencoder_ref.so, oracle/_ref/common/common_ref.so) under the stub modules of tests/golden/ref_shims.py.
What is NOT the reference's: kaolin's octree (absent, third-party, unpinned) is replaced by the product's DDA-backed OctreeManager,
which honours the same `ray_trace` contract (Utils.py:443-475), and pytorch3d's se3_exp_map by the oracle's closed form.

Two uses, both test-side only (never imported by bundlesdf_b200/):
  * `time_reference(c, name)`  — bench.py's cpu_baseline leg: rays/s of the reference's train_loop on the same workload, same GPU.
  * `golden_step(...)`         — tests/golden/make_golden_step.py: one train_loop on a small seeded scene, every tensor the parity
                                  test compares (inputs, z_vals, rgb, loss, gradients) captured from INSIDE the reference's own call.
"""
import importlib
import importlib.util
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.path.join(HERE, '_ref')


def _load_ext(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_MODS = None


def reference_modules():
    """(nerf_helpers, nerf_runner, Utils) of the reference, importable on the GPU box; raises when oracle/_ref is not built."""
    global _MODS
    if _MODS is not None:
        return _MODS
    ge_so = os.path.join(REF, 'gridencoder', 'gridencoder_ref.so')
    cm_so = os.path.join(REF, 'common', 'common_ref.so')
    if not (os.path.exists(ge_so) and os.path.exists(cm_so)):
        raise RuntimeError('oracle/_ref extensions are not built (python oracle/build_ref.py in the container that mounts /root/reference)')
    ge = _load_ext('gridencoder_ref', ge_so)
    cm = _load_ext('common_ref', cm_so)
    sys.path.insert(0, os.path.join(REPO, 'tests', 'golden'))
    import ref_shims
    ref_py = '/root/reference' if os.path.isdir('/root/reference') else os.path.join(REF, 'py')
    # the reference's get_embedder does `from mycuda.torch_ngp_grid_encoder.grid import GridEncoder`: give that dotted name a module
    sys.modules['gridencoder'] = ge
    gdir = os.path.join(ref_py, 'mycuda', 'torch_ngp_grid_encoder')
    spec = importlib.util.spec_from_file_location('mycuda.torch_ngp_grid_encoder.grid', os.path.join(gdir, 'grid.py'))
    grid_mod = importlib.util.module_from_spec(spec)
    pkg2 = types.ModuleType('mycuda.torch_ngp_grid_encoder')
    pkg2.__path__ = [gdir]
    sys.modules['mycuda.torch_ngp_grid_encoder'] = pkg2
    sys.modules['mycuda.torch_ngp_grid_encoder.grid'] = grid_mod
    nh, nr, U = ref_shims.import_reference(ref_py, mycuda_common=cm, mycuda_gridencoder=ge)
    sys.modules['mycuda'].torch_ngp_grid_encoder = pkg2
    sys.modules['mycuda.torch_ngp_grid_encoder'] = pkg2
    sys.modules['mycuda.torch_ngp_grid_encoder.grid'] = grid_mod
    spec.loader.exec_module(grid_mod)
    pkg2.grid = grid_mod
    _MODS = (nh, nr, U)
    return _MODS


def build_reference_runner(ours, init_scale=65536.0):
    """A reference NerfRunner that shares `ours`' data (cfg, ray pool, occupancy, initial poses) and owns its own models/optimizer,
    built by the reference's create_nerf / create_optimizer. __init__ is bypassed (it needs open3d / kaolin / cv2 windows)."""
    nh, nr, U = reference_modules()
    ref = object.__new__(nr.NerfRunner)
    ref.cfg = dict(ours.cfg)
    ref.cfg.setdefault('tv_loss_weight', 0)
    ref._run = None
    ref.images, ref.depths, ref.masks, ref.poses = ours.images, ours.depths, ours.masks, ours.poses
    ref.octree_m = ours.octree_m                     # same ray_trace contract as Utils.OctreeManager (see bundlesdf_b200/occupancy.py)
    ref.c2w_array = ours.c2w_array.clone()
    ref.rays = ours.rays
    ref.global_step = 0
    ref.N_iters = ref.cfg['n_step'] + 1
    ref.H, ref.W, ref.K = ours.H, ours.W, ours.K
    ref.ray_dir_slice, ref.ray_rgb_slice, ref.ray_depth_slice, ref.ray_mask_slice = [0, 1, 2], [3, 4, 5], 6, 7
    ref.ray_frame_id_slice, ref.ray_type_slice, ref.ray_near_slice, ref.ray_far_slice = 8, 9, 10, 11
    ref.create_nerf()
    ref.create_optimizer()
    ref.amp_scaler = torch.cuda.amp.GradScaler(init_scale=init_scale, enabled=bool(ref.cfg['amp']))   # nerf_runner.py:159 (default 65536)
    ref.data_loader = nr.DataLoader(rays=ref.rays, batch_size=ref.cfg['N_rand'])
    return ref


def time_reference(c, name, steps=10, warmup=3):
    """rays/s of the reference's train_loop on workload `c` (a bench.CONFIGS entry) on this GPU."""
    sys.path.insert(0, REPO)
    import bench
    dev = torch.device('cuda', torch.cuda.current_device())
    ours, _ = bench.build_runner(c, seed=0, device=dev, eager=True)
    ref = build_reference_runner(ours)
    N = c['N']
    def one():
        batch = next(ref.data_loader)
        # torch >= 2 refuses to index the CPU id tensor with the CUDA mask of render_rays (nerf_runner.py:1075; torch 1.11, which the
        # reference pins, accepted it): hand the ids over on the batch's device. Nothing else of the loop is touched.
        ref.data_loader.batch_ray_ids = ref.data_loader.batch_ray_ids.to(batch.device)
        ref.train_loop(batch)
        ref.global_step += 1
    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {'value': N * steps / dt, 'unit': 'rays/s', 'ms_per_step': 1e3 * dt / steps, 'steps': steps, 'warmup': warmup, 'kind': 'reference-cuda',
            'what': "the reference's own NerfRunner.train_loop (nerf_runner.py:679-852) + its own gridencoder / common CUDA extensions compiled for "
                    "sm_100a (oracle/_ref), PyTorch eager, on this B200; kaolin's ray trace replaced by the product's DDA (third-party, absent)"}


def golden_step(ours, batch, seed=0, init_scale=1024.0):
    """One verbatim reference train_loop on `batch` with the reference's own freshly initialised models; returns a dict of numpy arrays
    with everything the step-level parity test needs. The gradients are read between backward() and the optimizer step.
    init_scale: GradScaler scale of the AMP run. With the reference's default 65536 the very first step overflows in fp16 (observed on
    the B200: sigma_net.2.bias.grad[0] = inf — the reference accumulates bias gradients in fp16 — and GradScaler skips the step); the
    golden uses 1024 so that every gradient of the step is finite and comparable."""
    ref = build_reference_runner(ours, init_scale=init_scale)
    torch.manual_seed(seed)
    cap = {}
    N = batch.shape[0]
    ref.data_loader = types.SimpleNamespace(batch_ray_ids=torch.arange(N, device=batch.device))
    real_render = ref.render

    def render(*a, **k):
        out = real_render(*a, **k)
        cap['rgb'] = out[0].detach().float().cpu().numpy()
        ex = out[1]
        cap['z_vals'] = ex['z_vals'].detach().float().cpu().numpy()
        cap['raw'] = ex['raw'].detach().float().cpu().numpy()
        cap['weights'] = ex['weights'].detach().float().cpu().numpy()
        cap['valid_samples'] = ex['valid_samples'].detach().cpu().numpy().astype(np.uint8)
        return out
    ref.render = render
    real_scale, real_step = ref.amp_scaler.scale, ref.amp_scaler.step

    def scale(loss):
        cap['loss'] = np.float32(loss.detach().float().item())
        return real_scale(loss)

    def step(opt):
        s = ref.amp_scaler.get_scale() if ref.amp_scaler.is_enabled() else 1.0
        cap['loss_scale'] = np.float32(s)
        cap['grad_embeddings'] = (ref.models['embed_fn'].embeddings.grad.detach().float() / s).cpu().numpy()
        for k, p in ref.models['model'].named_parameters():
            cap['grad_' + k] = (p.grad.detach().float() / s).cpu().numpy()
        if ref.models['pose_array'] is not None:
            cap['grad_pose'] = (ref.models['pose_array'].data.grad.detach().float() / s).cpu().numpy()
        return real_step(opt)
    ref.amp_scaler.scale, ref.amp_scaler.step = scale, step
    # parameters BEFORE the step
    cap['embeddings'] = ref.models['embed_fn'].embeddings.detach().float().cpu().numpy()
    cap['offsets'] = ref.models['embed_fn'].offsets.detach().cpu().numpy()
    cap['per_level_scale'] = np.float64(ref.models['embed_fn'].per_level_scale)
    for k, v in ref.models['model'].state_dict().items():
        cap['param_' + k] = v.detach().float().cpu().numpy()
    if ref.models['pose_array'] is not None:
        with torch.no_grad():
            ref.models['pose_array'].data.normal_(0, 0.1)            # non-trivial pose corrections (the reference initialises zeros)
        cap['pose_data'] = ref.models['pose_array'].data.detach().float().cpu().numpy()
    ref.train_loop(batch)
    torch.cuda.synchronize()
    cap['batch'] = batch.detach().float().cpu().numpy()
    cap['c2w'] = ref.c2w_array.detach().float().cpu().numpy()
    return cap
