"""Diagnostic: loss terms over a 500-step round (reference default n_step) on a small synthetic sequence, both policies."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bundlesdf_b200 import synthetic as syn
from bundlesdf_b200.nerf_runner import NerfRunner
for amp in (True, False):
    seq = syn.make_sequence(8, H=240, W=320, device='cuda', seed=3, pose_noise=True)
    cfg = syn.default_cfg(N_rand=2048, N_samples=64, N_samples_around_depth=64, num_levels=16, finest_res=256, log2_hashmap_size=19, amp=amp,
                          sc_factor=seq['sc_factor'], translation=seq['translation'].tolist(), n_step=500)
    r = NerfRunner(cfg, seq['images'], seq['depths'], seq['masks'], None, seq['poses'], seq['K'], build_octree_pcd=syn.PointCloud(seq['pcd_normalized']))
    for it in range(501):
        r.train_loop(next(r.data_loader)); r.global_step += 1
        if it % 50 == 0:
            m = r.get_metrics()
            print(json.dumps({'amp': amp, 'it': it, **{k: round(m[k], 5) for k in ('loss', 'rgb_loss', 'fs_loss', 'sdf_loss', 'valid_samples', 'valid_rays')}, 'scale': r.amp_scaler.get_scale()}))
    import numpy as np
    P = r.models['pose_array'].get_matrices(np.arange(8)).cpu().numpy() @ seq['poses']
    err = np.linalg.norm(P[:, :3, 3] - seq['poses_gt'][:, :3, 3], axis=-1) / seq['sc_factor']
    err0 = np.linalg.norm(seq['poses'][:, :3, 3] - seq['poses_gt'][:, :3, 3], axis=-1) / seq['sc_factor']
    print(json.dumps({'amp': amp, 'pose_trans_err_m_before': err0.round(4).tolist(), 'after': err.round(4).tolist()}))
