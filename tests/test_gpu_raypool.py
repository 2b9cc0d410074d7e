"""GPU: the device-side ray-pool construction against the REFERENCE's own functions, imported from its Python under the
stub modules of tests/golden/ref_shims.py (oracle/_ref/py on the GPU box, /root/reference in the build container):
  NerfRunner.make_frame_rays (nerf_runner.py:246-316) incl. compute_near_far_and_filter_rays (:39-65) and the cv2 mask dilation,
  and the octree-cloud denoise of __init__ (:178-195, inline there: restated with the same scipy cKDTree call).
The occupancy trace inside make_frame_rays goes through the product's OctreeManager in both (kaolin is absent)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _reference():
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import ref_shims
    ref_py = '/root/reference' if os.path.isdir('/root/reference') else os.path.join(REPO, 'oracle', '_ref', 'py')
    if not os.path.exists(os.path.join(ref_py, 'nerf_runner.py')):
        pytest.skip('reference Python not staged (oracle/build_ref.py stage_py)')
    return ref_shims.import_reference(ref_py)


def _runner(denoise):
    from bundlesdf_b200 import synthetic as syn
    from bundlesdf_b200.nerf_runner import NerfRunner
    seq = syn.make_sequence(4, H=120, W=160, device='cuda', seed=5, pose_noise=True)
    cfg = syn.default_cfg(N_rand=128, N_samples=32, N_samples_around_depth=32, num_levels=4, finest_res=128, log2_hashmap_size=12,
                          sc_factor=seq['sc_factor'], translation=seq['translation'].tolist(), denoise_depth_use_octree_cloud=denoise)
    r = NerfRunner(cfg, seq['images'], seq['depths'], seq['masks'], None, seq['poses'], seq['K'], build_octree_pcd=syn.PointCloud(seq['pcd_normalized']))
    return r, seq


def test_make_frame_rays_matches_the_reference():
    nh, nr, U = _reference()
    ours, seq = _runner(False)
    fake = types.SimpleNamespace(masks=ours.masks, images=ours.images, depths=ours.depths, poses=np.asarray(ours.poses), K=ours.K, H=ours.H, W=ours.W,
                                 cfg=ours.cfg, occ_masks=None, normal_maps=None, octree_m=ours.octree_m)
    for fid in (0, 2):
        want = nr.NerfRunner.make_frame_rays(fake, fid)                        # numpy float64 [R, 12]
        got = ours.make_frame_rays(fid).cpu().numpy()
        assert got.shape == want.shape and got.shape[0] > 500, (got.shape, want.shape)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)            # fp32 on the device vs fp64 numpy; same rows in the same order


def test_octree_cloud_denoise_matches_ckdtree():
    from scipy.spatial import cKDTree
    ours, seq = _runner(False)
    sc = ours.cfg['sc_factor']
    rays = torch.cat([ours.make_frame_rays(i) for i in range(4)], 0)
    # push a tenth of the depths off the surface so that the filter has something to remove
    g = torch.Generator(device='cpu').manual_seed(1)
    off = torch.rand(len(rays), generator=g).to(rays.device) < 0.1
    rays[off, 6] += 0.05 * sc
    got = ours._denoise_rays(rays.clone()).cpu().numpy()
    # the reference's lines (nerf_runner.py:178-195) on the same rows
    r = rays.cpu().numpy().astype(np.float64)
    mask = (r[:, 7] > 0) & (r[:, 6] <= ours.cfg['far'] * sc)
    pts = r[mask][:, 0:3] * r[mask][:, 6:7]
    fid = r[mask][:, 8].astype(int)
    P = np.asarray(ours.poses)
    pts_w = (P[fid] @ np.concatenate([pts, np.ones((len(pts), 1))], 1)[..., None])[:, :3, 0]
    d, _ = cKDTree(ours.build_octree_pts).query(pts_w, k=1)
    bad = d > 0.02 * sc
    margin = np.abs(d - 0.02 * sc) < 1e-5                                         # fp32-vs-fp64 ties
    keep = np.ones(len(r), bool)
    keep[np.arange(len(r))[mask][bad]] = False
    want = r[keep]
    assert bad.sum() > 50 and abs(len(got) - len(want)) <= int(margin.sum())
    if len(got) == len(want):
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
