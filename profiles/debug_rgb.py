import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundlesdf_b200 import synthetic as syn
from bundlesdf_b200.nerf_runner import NerfRunner
seq = syn.make_sequence(8, H=240, W=320, device='cuda', seed=3, pose_noise=False)
cfg = syn.default_cfg(N_rand=2048, N_samples=64, N_samples_around_depth=64, num_levels=16, finest_res=256, log2_hashmap_size=19, amp=False,
                      sc_factor=seq['sc_factor'], translation=seq['translation'].tolist(), n_step=500)
r = NerfRunner(cfg, seq['images'], seq['depths'], seq['masks'], None, seq['poses'], seq['K'], build_octree_pcd=syn.PointCloud(seq['pcd_normalized']))
batch = next(r.data_loader)
rgb, ex = r.render(batch, depth=batch[:, 6], perturb=True)
w = ex['weights']; raw = ex['raw']; z = ex['z_vals']; d = batch[:, 6:7]
tr = r.get_truncation()
print(json.dumps({'trunc': tr, 'sc': cfg['sc_factor'], 'rgb_map_mean': rgb.mean(0).tolist(), 'gt_mean': batch[:, 3:6].mean(0).tolist(),
                  'wsum_mean': w.sum(-1).mean().item(), 'wsum_min': w.sum(-1).min().item(), 'frac_wsum_lt_0.99': (w.sum(-1) < 0.99).float().mean().item(),
                  'valid_frac': ex['valid_samples'].float().mean().item(), 'logit_mean': raw[..., :3].mean().item(), 'logit_std': raw[..., :3].std().item(),
                  'sdf_mean': raw[..., 3].mean().item(), 'in_band_frac': (((z - d) <= tr) & ((z - d) >= -tr)).float().mean().item(),
                  'depth_minmax': [d.min().item(), d.max().item()], 'z_minmax': [z.min().item(), z.max().item()],
                  'frame_hist': torch.bincount(batch[:, 8].long(), minlength=8).tolist(),
                  'mse': ((rgb - batch[:, 3:6]) ** 2).mean().item()}))
m = None
r.train_loop(batch); print(json.dumps(r.get_metrics()))
