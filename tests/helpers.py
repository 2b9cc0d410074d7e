"""Test-side scene builder (seeded, small) shared by the CPU and GPU tests. Uses the oracle and the synthetic generator;
the batch rows follow the reference layout dir(3) rgb(3) depth mask frame type near far (nerf_runner.py:259-300)."""
import numpy as np
import torch

from bundlesdf_b200 import synthetic as syn
from oracle import nof_oracle as O


def make_cfg(L=4, finest=128, log2T=14, S_occ=32, S_depth=32, ff=0, **over):
    cfg = syn.default_cfg(num_levels=L, finest_res=finest, log2_hashmap_size=log2T, N_samples=S_occ, N_samples_around_depth=S_depth,
                          frame_features=ff)
    cfg.update(over)
    return cfg


def make_scene(n_frames=4, N=64, seed=0, cfg=None, H=120, W=160, pose_scale=0.15, invalid_frac=0.0, type1_frac=0.0):
    """Returns dict: cfg, batch [N,12] torch, c2w [F,4,4] torch, occ (np bool), level, params (oracle dict), seq."""
    cfg = cfg or make_cfg()
    seq = syn.make_sequence(n_frames, H=H, W=W, seed=seed)
    cfg = dict(cfg)
    cfg['sc_factor'] = seq['sc_factor']
    cfg['translation'] = seq['translation'].tolist()
    rng = np.random.default_rng(seed + 1)
    K = seq['K']
    rows = []
    for f in range(n_frames):
        vs, us = np.nonzero(seq['masks'][f, ..., 0])
        sel = rng.choice(len(vs), size=min(len(vs), max(1, N // n_frames + 4)), replace=False)
        vs, us = vs[sel], us[sel]
        dirs = np.stack([(us - K[0, 2]) / K[0, 0], -(vs - K[1, 2]) / K[1, 1], -np.ones_like(us, dtype=np.float64)], -1)   # nerf_helpers.py:358-363
        rgb = seq['images'][f, vs, us]
        depth = seq['depths'][f, vs, us, 0]
        r = np.zeros((len(vs), 12), np.float32)
        r[:, 0:3] = dirs; r[:, 3:6] = rgb; r[:, 6] = depth; r[:, 7] = 1; r[:, 8] = f; r[:, 9] = 0; r[:, 10] = 0.5; r[:, 11] = 8.0
        rows.append(r)
    rows = np.concatenate(rows, 0)
    rows = rows[rng.permutation(len(rows))[:N]]
    n_inv = int(round(invalid_frac * N))
    if n_inv:
        rows[:n_inv, 6] = syn.BAD_DEPTH * cfg['sc_factor']
    n_t1 = int(round(type1_frac * N))
    if n_t1:
        rows[n_inv:n_inv + n_t1, 9] = 1
    batch = torch.from_numpy(rows)
    c2w = torch.from_numpy(seq['poses']).float()
    occ, level = O.build_occupancy(seq['pcd_normalized'], cfg)
    # model
    offsets, pls = O.grid_offsets(cfg['num_levels'], cfg['base_res'], cfg['finest_res'], cfg['log2_hashmap_size'])
    g = torch.Generator().manual_seed(seed + 2)
    C = cfg['feature_grid_dim']
    E = cfg['num_levels'] * C
    V = 9 + cfg['frame_features']
    params = {'embeddings': (torch.rand(int(offsets[-1]), C, generator=g) * 2 - 1) * 0.3,   # larger than the 1e-4 init: exercises the maths
              'offsets': offsets, 'S': float(np.log2(pls)), 'H': cfg['base_res']}
    params.update(O.init_mlp(E, V, seed=seed + 3))
    params['pose_data'] = torch.randn(n_frames, 6, generator=g) * pose_scale if cfg['optimize_poses'] else None
    params['feature_data'] = torch.randn(n_frames, cfg['frame_features'], generator=g) if cfg['frame_features'] > 0 else None
    return dict(cfg=cfg, batch=batch, c2w=c2w, occ=occ, level=level, params=params, seq=seq, E=E, V=V)


def pack_mlp(params, E, V, count, offs):
    from bundlesdf_b200.ops import MLP_KEYS
    flat = torch.zeros(count, dtype=torch.float32)
    for k, o in zip(MLP_KEYS, offs):
        t = params[k].detach().float().reshape(-1)
        flat[o:o + t.numel()] = t
    return flat


def unpack_mlp(flat, E, V, offs):
    from bundlesdf_b200.ops import MLP_KEYS, mlp_shapes
    out = {}
    for k, o, shp in zip(MLP_KEYS, offs, mlp_shapes(E, V)):
        n = int(np.prod(shp))
        out[k] = flat[o:o + n].reshape(shp)
    return out


def run_fused_step(scene, amp, t_rand, dev='cuda', loss_scale=None, march=True, z_vals=None):
    """Drive pose_forward -> ray_march -> step_fused -> pose_backward through the C ABI for one batch."""
    from bundlesdf_b200 import ops
    cfg, P = scene['cfg'], scene['params']
    E, V = scene['E'], scene['V']
    N = scene['batch'].shape[0]
    S_occ, S_d = cfg['N_samples'], cfg['N_samples_around_depth']
    S = S_occ + S_d
    F = scene['c2w'].shape[0]
    sc = cfg['sc_factor']
    trunc = O.get_truncation(cfg, 0)
    batch = scene['batch'].to(dev).contiguous()
    c2w = scene['c2w'].to(dev).contiguous()
    pose = P['pose_data'].detach().to(dev).contiguous() if P.get('pose_data') is not None else None
    tf = ops.pose_forward(pose, c2w, cfg['max_trans'] * sc, cfg['max_rot'])
    occ_bits = ops.pack_occupancy(scene['occ']).to(dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    inter = None
    if z_vals is None:
        tr = None if t_rand is None else torch.as_tensor(t_rand, dtype=torch.float32).to(dev).contiguous()
        z_vals, inter = ops.ray_march(batch, tf, occ_bits, scene['level'], S_occ, S_d, trunc, cfg['near'] * sc, cfg['far'] * sc,
                                      cfg['neg_trunc_ratio'], t_rand=tr, perturb=t_rand is not None, want_intervals=True, err_flag=err)
    else:
        z_vals = z_vals.to(dev).contiguous()
    count, offs = ops.mlp_param_layout(E, V)
    mlp = pack_mlp(P, E, V, count, offs).to(dev)
    emb = P['embeddings'].detach().to(dev).contiguous()
    emb16 = emb.half()
    sb = ops.StepBuffers()
    out = dict(grad_table=torch.zeros_like(emb), grad_mlp=torch.zeros(count, device=dev), grad_tf=torch.zeros(F, 12, device=dev),
               losses=torch.zeros(8, device=dev), found_inf=torch.zeros(1, dtype=torch.int32, device=dev),
               rgb_map=torch.zeros(N, 3, device=dev), raw=torch.zeros(N, S, 4, device=dev),
               valid_samples=torch.zeros(N, S, dtype=torch.uint8, device=dev), weights=torch.zeros(N, S, device=dev))
    feat = P['feature_data'].detach().to(dev).contiguous() if P.get('feature_data') is not None else None
    if feat is not None:
        out['grad_feat'] = torch.zeros_like(feat)
    ls = None
    if loss_scale is not None:
        ls = torch.tensor([float(loss_scale)], device=dev)
    sb.set_scalars(N=N, S=S, L=cfg['num_levels'], C=2, F=F, ff=cfg['frame_features'], ray_dim=12, amp=int(amp),
                   S_log2=float(P['S']), H=int(P['H']), need_pose_grad=int(pose is not None))
    ops.fill_step_cfg(sb, cfg, trunc)
    sb.set(offsets=torch.from_numpy(P['offsets']).to(dev), table_f32=emb, table_f16=emb16, mlp=mlp, feat=feat, rays=batch, tf=tf,
           z_vals=z_vals, loss_scale=ls, **out)
    ws = torch.zeros(max(sb.workspace_bytes(), 256), dtype=torch.uint8, device=dev)
    sb.set(workspace=ws)
    sb.launch()
    res = {k: v for k, v in out.items()}
    res['grad_mlp_named'] = unpack_mlp(out['grad_mlp'], E, V, offs)
    if pose is not None:
        gp = torch.zeros(F, 6, device=dev)
        ops.pose_backward(pose, c2w, out['grad_tf'], gp, cfg['max_trans'] * sc, cfg['max_rot'], ls)
        res['grad_pose'] = gp
    torch.cuda.synchronize()
    res.update(z_vals=z_vals, intervals=inter, tf=tf, march_err=int(err.item()), sb=sb)
    return res
