"""Step-level golden vectors from the reference's OWN train_loop (nerf_runner.py:679-852) executed on a B200 on top of its own compiled
CUDA extensions (oracle/ref_train_loop.py; oracle/_ref built by oracle/build_ref.py). Run on the GPU box:

    python tests/golden/make_golden_step.py        ->  tests/golden/ref_gpu_train_step_amp{0,1}.npz   (via gpurun_out/golden/)

Each file holds the inputs (batch, c2w, pose_data, the reference's freshly initialised table and MLP), the z_vals the reference sampled,
and what its train_loop produced: rgb_map, raw, weights, valid_samples, the loss and every gradient (read between backward and the
optimizer step, divided by the GradScaler scale). tests/test_gpu_golden.py feeds the same inputs to nof_step_fused."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))


def main():
    import ref_train_loop
    from bundlesdf_b200 import synthetic as syn
    from bundlesdf_b200.nerf_runner import NerfRunner
    out_dir = os.path.join(REPO, 'gpurun_out', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for amp in (0, 1):
        seq = syn.make_sequence(5, H=120, W=160, device='cuda', seed=3, pose_noise=True)
        cfg = syn.default_cfg(N_rand=192, N_samples=64, N_samples_around_depth=64, num_levels=16, finest_res=256, log2_hashmap_size=12, amp=bool(amp),
                              sc_factor=seq['sc_factor'], translation=seq['translation'].tolist(), n_step=60, chunk=99999999999, netchunk=6553600)
        ours = NerfRunner(cfg, seq['images'], seq['depths'], seq['masks'], None, seq['poses'], seq['K'], build_octree_pcd=syn.PointCloud(seq['pcd_normalized']))
        batch = next(ours.data_loader)
        cap = ref_train_loop.golden_step(ours, batch, seed=7)
        cap['sc_factor'] = np.float64(seq['sc_factor'])
        cap['occ'] = ours.octree_m.occ.cpu().numpy()
        cap['level'] = np.int64(ours.octree_m.level)
        path = os.path.join(out_dir, f'ref_gpu_train_step_amp{amp}.npz')
        np.savez_compressed(path, **cap)
        print(path, {k: getattr(v, 'shape', v) for k, v in cap.items() if not k.startswith('param_') and not k.startswith('grad_')}, 'loss', cap['loss'])


if __name__ == '__main__':
    main()
