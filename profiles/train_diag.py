"""Diagnostic: loss terms over a training round on a small synthetic sequence (not a test; printed as JSON lines).
  exp A: 8 frames 320x240, 2048 rays x (64+64), L=16 T=2^19, pose noise, 500 steps  (reference default n_step)
  exp B: same but ray pool restricted to 4096 rays, 512 rays/step, constant lr — mirrors the CPU-oracle experiment
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bundlesdf_b200 import synthetic as syn
from bundlesdf_b200.nerf_runner import DataLoader, NerfRunner


def run(tag, amp, N, steps, pool=None, noise=True, T=19, H=240, W=320, n_step=500):
    seq = syn.make_sequence(8, H=H, W=W, device='cuda', seed=3, pose_noise=noise)
    cfg = syn.default_cfg(N_rand=N, N_samples=64, N_samples_around_depth=64, num_levels=16, finest_res=256, log2_hashmap_size=T, amp=amp,
                          sc_factor=seq['sc_factor'], translation=seq['translation'].tolist(), n_step=n_step)
    r = NerfRunner(cfg, seq['images'], seq['depths'], seq['masks'], None, seq['poses'], seq['K'], build_octree_pcd=syn.PointCloud(seq['pcd_normalized']))
    fr = r.rays[:, 8].long()
    print(json.dumps({'tag': tag, 'pool': int(r.rays.shape[0]), 'rays_per_frame': torch.bincount(fr, minlength=8).tolist(),
                      'gt_rgb_mean': r.rays[:, 3:6].mean(0).tolist(), 'gt_rgb_var': r.rays[:, 3:6].var(0).tolist()}))
    if pool:
        idx = torch.randperm(r.rays.shape[0], device='cuda')[:pool]
        r.rays = r.rays[idx].contiguous()
        r.data_loader = DataLoader(r.rays, N)
    for it in range(steps + 1):
        r.train_loop(next(r.data_loader)); r.global_step += 1
        if it % max(steps // 10, 1) == 0:
            m = r.get_metrics()
            print(json.dumps({'tag': tag, 'amp': amp, 'it': it, **{k: round(m[k], 5) for k in ('loss', 'rgb_loss', 'fs_loss', 'sdf_loss', 'valid_rays')}}))
    P = r.models['pose_array'].get_matrices(np.arange(8)).cpu().numpy() @ seq['poses']
    err = np.linalg.norm(P[:, :3, 3] - seq['poses_gt'][:, :3, 3], axis=-1) / seq['sc_factor']
    err0 = np.linalg.norm(seq['poses'][:, :3, 3] - seq['poses_gt'][:, :3, 3], axis=-1) / seq['sc_factor']
    print(json.dumps({'tag': tag, 'pose_trans_err_m_before': err0.round(4).tolist(), 'after': err.round(4).tolist()}))


run('A_amp', True, 2048, 500)
run('B_pool4096', True, 512, 300, pool=4096, n_step=100000)
run('C_nonoise', True, 2048, 300, noise=False)
