"""Thin functional wrappers over the fused C-ABI entry points (include/nof.h §2). Tensors in, tensors out; all launches
go to torch's current stream; nothing here synchronises with the host."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import NofAdamSeg, NofMarchCfg, NofPrologue, NofStep


def mlp_param_layout(E, V):
    """(padded element count, offsets[10]) of the packed MLP block W1 b1 W2 b2 W3 b3 W4 b4 W5 b5."""
    lib = _lib.load()
    offs = (C.c_int32 * 10)()
    _lib.check(lib.nof_mlp_param_offsets(E, V, offs), 'nof_mlp_param_offsets')
    return int(lib.nof_mlp_param_count(E, V)), [int(o) for o in offs]


MLP_KEYS = ['sigma_net.0.weight', 'sigma_net.0.bias', 'sigma_net.2.weight', 'sigma_net.2.bias', 'color_net.0.weight',
            'color_net.0.bias', 'color_net.2.weight', 'color_net.2.bias', 'color_net.4.weight', 'color_net.4.bias']


def mlp_shapes(E, V):
    return [(64, E), (64,), (16, 64), (16,), (64, V + 15), (64,), (64, 64), (64,), (3, 64), (3,)]


def pack_occupancy(occ):
    """bool [n,n,n] (indexed [x,y,z]) -> int32 tensor of bit words, bit = (ix*n+iy)*n+iz (host side, numpy)."""
    flat = np.asarray(occ, dtype=bool).reshape(-1)
    pad = (-len(flat)) % 32
    if pad:
        flat = np.concatenate([flat, np.zeros(pad, bool)])
    words = np.packbits(flat.reshape(-1, 32)[:, ::-1], axis=1).view('>u4').astype(np.uint32).reshape(-1)
    return torch.from_numpy(words.view(np.int32).copy())


def grid_level_scales(S, H, L, device='cuda'):
    """Per-level `scale` as the device computes it (exp2f), float32 tensor [L]."""
    lib = _lib.load()
    out = torch.empty(L, device=device, dtype=torch.float32)
    _lib.check(lib.nof_grid_level_scales(float(S), int(H), int(L), _lib.ptr(out), _lib.stream()), 'nof_grid_level_scales')
    return out


def pose_forward(pose_data, c2w, max_trans, max_rot_deg, out=None):
    lib = _lib.load()
    F = c2w.shape[0]
    tf = out if out is not None else torch.empty(F, 12, device=c2w.device, dtype=torch.float32)
    _lib.check(lib.nof_pose_forward(_lib.ptr(pose_data), _lib.ptr(c2w), _lib.ptr(tf), F, float(max_trans), float(max_rot_deg),
                                    _lib.stream()), 'nof_pose_forward')
    return tf


def pose_backward(pose_data, c2w, grad_tf, grad_pose, max_trans, max_rot_deg, loss_scale=None):
    lib = _lib.load()
    _lib.check(lib.nof_pose_backward(_lib.ptr(pose_data), _lib.ptr(c2w), _lib.ptr(grad_tf), _lib.ptr(grad_pose), c2w.shape[0],
                                     float(max_trans), float(max_rot_deg), _lib.ptr(loss_scale), _lib.stream()), 'nof_pose_backward')
    return grad_pose


def step_prologue(pose_data, c2w, tf, max_trans, max_rot_deg, pool=None, ids=None, batch=None, cursor=None, tick=None, done=None,
                  trunc_table=None, gstep=None, trunc_out=None):
    """nof_step_prologue: [batch gather at the device cursor] + pose correction of all frames + counter bumps, one launch."""
    lib = _lib.load()
    F = c2w.shape[0]
    p = NofPrologue(_lib.ptr(pool, pinned_ok=True), _lib.ptr(ids), int(ids.shape[0]) if ids is not None else 0, _lib.ptr(batch),
                    int(batch.shape[0]) if batch is not None else 0, int(batch.shape[1]) if batch is not None else 0, _lib.ptr(cursor),
                    _lib.ptr(pose_data), _lib.ptr(c2w), _lib.ptr(tf), F, float(max_trans), float(max_rot_deg), _lib.ptr(tick), _lib.ptr(done),
                    _lib.ptr(trunc_table), int(trunc_table.shape[0]) if trunc_table is not None else 0, _lib.ptr(gstep), _lib.ptr(trunc_out))
    _lib.check(lib.nof_step_prologue(C.byref(p), _lib.stream()), 'nof_step_prologue')
    return tf


def gather_rays(pool, ids, out=None):
    lib = _lib.load()
    N, D = ids.shape[0], pool.shape[1]
    batch = out if out is not None else torch.empty(N, D, device=pool.device, dtype=torch.float32)
    _lib.check(lib.nof_gather_rays(_lib.ptr(pool, pinned_ok=True), _lib.ptr(ids), _lib.ptr(batch, pinned_ok=True), N, D, _lib.stream()), 'nof_gather_rays')
    return batch


def ray_march(rays, tf, occ_bits, level, S_occ, S_depth, trunc, near_sc, far_sc, neg_trunc_ratio, t_rand=None, perturb=True,
              seed=0, offset=0, I_max=None, z_vals=None, want_intervals=False, err_flag=None, offset_ptr=None, trunc_ptr=None):
    lib = _lib.load()
    N, D = rays.shape
    if I_max is None:
        I_max = 3 * (1 << level)
    S = S_occ + S_depth
    if z_vals is None:
        z_vals = torch.empty(N, S, device=rays.device, dtype=torch.float32)
    inter = torch.empty(N, I_max, 2, device=rays.device, dtype=torch.float32) if want_intervals else None
    cfg = NofMarchCfg(N, D, S_occ, S_depth, level, I_max, float(trunc), float(near_sc), float(far_sc), float(neg_trunc_ratio),
                      int(bool(perturb)), int(seed), int(offset), _lib.ptr(offset_ptr), _lib.ptr(trunc_ptr))
    _lib.check(lib.nof_ray_march(C.byref(cfg), _lib.ptr(rays), _lib.ptr(tf), _lib.ptr(occ_bits), _lib.ptr(t_rand), _lib.ptr(z_vals),
                                 _lib.ptr(inter), _lib.ptr(err_flag), _lib.stream()), 'nof_ray_march')
    return (z_vals, inter) if want_intervals else z_vals


class StepBuffers:
    """Owns the NofStep argument block and keeps every tensor it points to alive."""

    def __init__(self):
        self.s = NofStep()
        self.keep = {}

    def set(self, **tensors):
        for k, t in tensors.items():
            self.keep[k] = t
            setattr(self.s, k, _lib.ptr(t))

    def set_scalars(self, **kw):
        for k, v in kw.items():
            setattr(self.s, k, v)

    def workspace_bytes(self):
        return int(_lib.load().nof_step_workspace_bytes(C.byref(self.s)))

    def launch(self):
        _lib.check(_lib.load().nof_step_fused(C.byref(self.s), _lib.stream()), 'nof_step_fused')


LOSS_CFG_KEYS = ['sdf_lambda', 'neg_trunc_ratio', 'rgb_weight', 'fs_weight', 'empty_weight', 'trunc_weight', 'fs_sdf',
                 'fs_rgb_weight', 'first_frame_weight', 'eikonal_weight']


def fill_step_cfg(sb, cfg, trunc):
    sc = cfg['sc_factor']
    sb.set_scalars(trunc=float(trunc), near_sc=float(cfg['near'] * sc), far_sc=float(cfg['far'] * sc))
    sb.set_scalars(**{k: float(cfg.get(k, 0)) for k in LOSS_CFG_KEYS})


def _adam_segs(segs):
    arr = (NofAdamSeg * len(segs))()
    for i, s in enumerate(segs):
        n = s['param'].numel()
        arr[i] = NofAdamSeg(_lib.ptr(s['param']), _lib.ptr(s['grad']), _lib.ptr(s['exp_avg']), _lib.ptr(s['exp_avg_sq']),
                            _lib.ptr(s.get('shadow_f16')), n, float(s['lr']), s.get('lr_ptr'))
    return arr


def _check_step_buf(step):
    if step is not None and (step.numel() < 8 or step.dtype != torch.int32):
        raise _lib.NofError('nof_adam_*: `step` must be a device int32[8] tensor (see include/nof.h)')


def adam_step(segs, beta1, beta2, eps, step, scale_state=None, found_inf=None, tick=None):
    """segs: list of dict(param, grad, exp_avg, exp_avg_sq, shadow_f16|None, lr). step: device int32 tensor [8] (include/nof.h:
    [0] update count, the rest library scratch) or None."""
    _check_step_buf(step)
    lib = _lib.load()
    arr = _adam_segs(segs)
    _lib.check(lib.nof_adam_step(arr, len(segs), float(beta1), float(beta2), float(eps), _lib.ptr(step), _lib.ptr(scale_state),
                                 _lib.ptr(found_inf), _lib.ptr(tick), _lib.stream()), 'nof_adam_step')


def adam_update(segs, beta1, beta2, eps, step, scale_state=None, found_inf=None):
    """nof_adam_update: the update of `segs` only (reads step / scale / found_inf, modifies none of them)."""
    _check_step_buf(step)
    lib = _lib.load()
    arr = _adam_segs(segs)
    _lib.check(lib.nof_adam_update(arr, len(segs), float(beta1), float(beta2), float(eps), _lib.ptr(step), _lib.ptr(scale_state),
                                   _lib.ptr(found_inf), _lib.stream()), 'nof_adam_update')


def adam_tile_count(segs):
    """nof_adam_tile_count: thread blocks one update launch of `segs` takes (the unit nof_adam_update_shared counts in)."""
    n = _lib.load().nof_adam_tile_count(_adam_segs(segs), len(segs))
    if n < 0:
        _lib.check(n, 'nof_adam_tile_count')
    return n


def adam_update_shared(segs, beta1, beta2, eps, step, scale_state, found_inf, tick, total_tiles):
    """nof_adam_update_shared: like adam_update; the launch (of those sharing `total_tiles`) that retires last does the step's bookkeeping."""
    _check_step_buf(step)
    arr = _adam_segs(segs)
    _lib.check(_lib.load().nof_adam_update_shared(arr, len(segs), float(beta1), float(beta2), float(eps), _lib.ptr(step), _lib.ptr(scale_state),
                                                  _lib.ptr(found_inf), _lib.ptr(tick), int(total_tiles), _lib.stream()), 'nof_adam_update_shared')


def adam_finish(beta1, beta2, step, scale_state=None, found_inf=None, tick=None):
    """nof_adam_finish: the bookkeeping of one optimizer step after all nof_adam_update launches of that step."""
    _check_step_buf(step)
    _lib.check(_lib.load().nof_adam_finish(_lib.ptr(step), _lib.ptr(scale_state), _lib.ptr(found_inf), _lib.ptr(tick), float(beta1),
                                           float(beta2), _lib.stream()), 'nof_adam_finish')


def query_sdf(sb, x, out=None):
    lib = _lib.load()
    P = x.shape[0]
    sdf = out if out is not None else torch.empty(P, device=x.device, dtype=torch.float32)
    _lib.check(lib.nof_query_sdf(C.byref(sb.s), _lib.ptr(x), _lib.ptr(sdf), P, _lib.stream()), 'nof_query_sdf')
    return sdf


def cloud_within_radius(query, cloud, radius):
    """bool [Q]: is some point of `cloud` [M,3] within `radius` of query[i]? Uniform-grid binning on the device (nof_cloud_within_radius)."""
    lib = _lib.load()
    query = query.contiguous().float()
    cloud = cloud.contiguous().float()
    if len(cloud) == 0 or len(query) == 0:
        return torch.zeros(len(query), dtype=torch.bool, device=query.device)
    lo = float(torch.minimum(cloud.min(), query.min()).item()) - 1e-3
    hi = float(torch.maximum(cloud.max(), query.max()).item()) + 1e-3
    n = max(1, min(256, int((hi - lo) / radius)))            # cell >= radius
    cell = (hi - lo) / n
    ci = torch.clamp(torch.floor((cloud - lo) / cell).long(), 0, n - 1)
    key = (ci[:, 0] * n + ci[:, 1]) * n + ci[:, 2]
    order = torch.argsort(key)
    counts = torch.bincount(key, minlength=n * n * n)
    cell_start = torch.zeros(n * n * n + 1, dtype=torch.int32, device=cloud.device)
    cell_start[1:] = torch.cumsum(counts, 0).int()
    within = torch.empty(len(query), dtype=torch.uint8, device=query.device)
    _lib.check(lib.nof_cloud_within_radius(_lib.ptr(query), len(query), _lib.ptr(cloud[order].contiguous()), _lib.ptr(cell_start), lo, float(cell), n,
                                           float(radius), _lib.ptr(within), _lib.stream()), 'nof_cloud_within_radius')
    return within.bool()


def marching_tets(field, iso=0.0):
    """Iso-surface of a dense [nx,ny,nz] fp32 CUDA grid (include/nof.h nof_marching_tets_*). Returns (vertices [V,3] fp32 in
    grid-index coordinates, faces [F,3] int64), vertices welded, triangles facing increasing field values."""
    lib = _lib.load()
    field = field.contiguous().float()
    nx, ny, nz = (int(x) for x in field.shape)
    cells = (nx - 1) * (ny - 1) * (nz - 1)
    counts = torch.empty(cells, dtype=torch.int32, device=field.device)
    _lib.check(lib.nof_marching_tets_count(_lib.ptr(field), nx, ny, nz, float(iso), _lib.ptr(counts), _lib.stream()), 'nof_marching_tets_count')
    csum = torch.cumsum(counts, 0, dtype=torch.int64)
    T = int(csum[-1].item()) if cells > 0 else 0
    if T == 0:
        return torch.zeros(0, 3, device=field.device), torch.zeros(0, 3, dtype=torch.int64, device=field.device)
    offsets = (csum - counts).contiguous()
    verts = torch.empty(T, 3, 3, device=field.device, dtype=torch.float32)
    keys = torch.empty(T, 3, dtype=torch.int64, device=field.device)
    _lib.check(lib.nof_marching_tets_emit(_lib.ptr(field), nx, ny, nz, float(iso), _lib.ptr(offsets), _lib.ptr(verts), _lib.ptr(keys),
                                          _lib.stream()), 'nof_marching_tets_emit')
    uniq, inv = torch.unique(keys.reshape(-1), return_inverse=True)
    vertices = torch.empty(len(uniq), 3, device=field.device, dtype=torch.float32)
    vertices[inv] = verts.reshape(-1, 3)                     # equal keys carry bit-identical positions
    faces = inv.reshape(-1, 3)
    keep = (faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])
    return vertices, faces[keep].contiguous()
