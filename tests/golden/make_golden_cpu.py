"""Generate golden vectors by executing the REFERENCE's own Python (nerf_helpers.py / nerf_runner.py /
Utils.py imported from /root/reference under ref_shims) on CPU. Run in the build container only:

    python tests/golden/make_golden_cpu.py      ->  tests/golden/ref_py_*.npz

Each fixture stores the seeded inputs and the reference's outputs; tests/test_oracle_golden.py checks
oracle/nof_oracle.py against them, and the -m gpu tests check the CUDA path against the same files.
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402


def ref_cfg():
    cfg = yaml.safe_load(open('/root/reference/config.yml'))
    cfg['sc_factor'] = 4.2
    cfg['translation'] = [0.0, 0.0, 0.0]
    return cfg


def main():
    nh, nr, U = ref_shims.import_reference()
    torch.set_num_threads(4)
    cfg = ref_cfg()
    sv = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})

    # ---- 1. SHEncoder degree 3 (nerf_helpers.py:67-105)
    g = torch.Generator().manual_seed(1)
    d = torch.randn(257, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    sh = nh.SHEncoder(degree=3)(d)
    sv('ref_py_sh.npz', dirs=d, sh=sh)

    # ---- 2. NeRFSmall (nerf_helpers.py:243-321), both encoder widths and with frame features
    for tag, E, V in [('L16', 32, 9), ('L4', 8, 9), ('L16ff2', 32, 11)]:
        torch.manual_seed(7)
        m = nh.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                         input_ch=E, input_ch_views=V)
        x = torch.randn(513, E + V, generator=g) * 0.5
        x.requires_grad_(True)
        y = m(x)
        sdf_only = m.forward_sdf(x[:, :E])
        gy = torch.randn(y.shape, generator=g)
        (y * gy).sum().backward()
        kw = {('p_' + k): v for k, v in m.state_dict().items()}
        kw.update({('g_' + k): p.grad for k, p in m.named_parameters()})
        sv(f'ref_py_mlp_{tag}.npz', x=x, y=y, sdf=sdf_only, gy=gy, gx=x.grad, **kw)

    # ---- 3. PoseArray.get_matrices (nerf_helpers.py:143-154); se3_exp_map is the shimmed restatement
    torch.manual_seed(3)
    pa = nh.PoseArray(6, max_trans=cfg['max_trans'] * cfg['sc_factor'], max_rot=cfg['max_rot'])
    with torch.no_grad():
        pa.data.copy_(torch.randn(6, 6, generator=g) * 0.3)
        pa.data[2, 3:] = 0.0                      # exercises the clamped small-angle branch
        pa.data[3, 3:] *= 1e-3
    ids = torch.tensor([0, 3, 5, 1, 2, 0, 4])
    Ts = pa.get_matrices(ids)
    sv('ref_py_pose.npz', data=pa.data, ids=ids, Ts=Ts, max_trans=cfg['max_trans'] * cfg['sc_factor'], max_rot=cfg['max_rot'])

    # ---- 4. sample_rays_uniform (nerf_runner.py:67-87), perturb on and off
    near = torch.rand(33, 1, generator=g) * 0.2
    far = near + torch.rand(33, 1, generator=g) * 0.5 + 0.01
    z0 = nr.sample_rays_uniform(64, near, far, lindisp=False, perturb=False)
    torch.manual_seed(11)
    z1 = nr.sample_rays_uniform(64, near, far, lindisp=False, perturb=True)
    torch.manual_seed(11)
    t_rand = torch.rand(33, 64)
    sv('ref_py_sample_uniform.npz', near=near, far=far, z_noperturb=z0, z_perturb=z1, t_rand=t_rand)

    # ---- 5. get_sdf_loss / get_masks (nerf_helpers.py:367-399)
    N, S = 48, 40
    sc = cfg['sc_factor']
    trunc = cfg['trunc'] * sc
    target_d = torch.rand(N, generator=g) * 2.0 + 1.0
    target_d[::7] = 99 * sc                                         # BAD_DEPTH rays (Utils.py:34)
    z_vals = target_d.clamp(max=4.0)[:, None] + (torch.rand(N, S, generator=g) - 0.5) * 6 * trunc
    sdf = torch.randn(N, S, generator=g) * 0.6
    sw = torch.rand(N, S, generator=g)
    rays_d = torch.randn(N, 3, generator=g)
    fs, sl, front, smask = nh.get_sdf_loss(z_vals, target_d.reshape(-1, 1).expand(-1, S), sdf, trunc, cfg,
                                           return_mask=True, sample_weights=sw, rays_d=rays_d)
    sv('ref_py_sdf_loss.npz', z_vals=z_vals, target_d=target_d, sdf=sdf, sample_weights=sw, trunc=trunc,
       fs_loss=fs, sdf_loss=sl, front_mask=front, sdf_mask=smask, sc_factor=sc)

    # ---- 6. raw2outputs (nerf_runner.py:1132-1169) + the loss assembly of train_loop (:679-758), executed verbatim
    #         on a stand-in `self` whose render() returns raw2outputs of a leaf `raw` tensor, so that the reference's
    #         own train_loop produces loss and dloss/draw.
    for tag, ffw, use_pose in [('a', 10, True), ('b', 1, False)]:
        N, S = 64, 48
        cfg2 = dict(cfg)
        cfg2['first_frame_weight'] = ffw
        cfg2['i_print'] = 999999
        batch = torch.zeros(N, 12)
        batch[:, 0:2] = (torch.rand(N, 2, generator=g) - 0.5)
        batch[:, 2] = -1
        batch[:, 3:6] = torch.rand(N, 3, generator=g)
        depth = torch.rand(N, generator=g) * 2.0 + 1.5
        depth[5::9] = 99 * sc
        batch[:, 6] = depth
        batch[:, 7] = 1
        batch[:, 8] = torch.randint(0, 4, (N,), generator=g).float()
        batch[:, 9] = 0
        batch[3::11, 9] = 1
        batch[:, 10] = 0.5
        batch[:, 11] = 5.0
        z_vals = depth.clamp(max=4.0)[:, None] + (torch.rand(N, S, generator=g) - 0.5) * 5 * trunc
        valid_samples = torch.rand(N, S, generator=g) > 0.1
        valid_samples[7] = False
        raw = (torch.randn(N, S, 4, generator=g) * 0.7).requires_grad_(True)

        fake = types.SimpleNamespace()
        fake.cfg = cfg2
        fake.global_step = 1
        fake.ray_dir_slice = [0, 1, 2]; fake.ray_rgb_slice = [3, 4, 5]; fake.ray_depth_slice = 6
        fake.ray_mask_slice = 7; fake.ray_frame_id_slice = 8; fake.ray_type_slice = 9
        fake.ray_near_slice = 10; fake.ray_far_slice = 11
        fake.data_loader = types.SimpleNamespace(batch_ray_ids=torch.arange(N))
        fake.get_truncation = types.MethodType(nr.NerfRunner.get_truncation, fake)
        fake.raw2outputs = types.MethodType(nr.NerfRunner.raw2outputs, fake)
        pose = None
        if use_pose:
            pose = nh.PoseArray(4, max_trans=cfg['max_trans'] * sc, max_rot=cfg['max_rot'])
        fake.models = {'feature_array': None, 'pose_array': pose}
        fake.optimizer = mock.MagicMock()
        captured = {}
        class Scaler:
            def scale(self, loss):
                captured['loss'] = loss.detach().clone()
                return loss
            def step(self, opt): pass
            def update(self): pass
        fake.amp_scaler = Scaler()
        fake._run = None
        def render(rays, ray_ids=None, frame_ids=None, depth=None, **kw):
            rgb_map, weights = fake.raw2outputs(raw, z_vals, rays[:, :3], valid_samples=valid_samples, depth=depth)
            captured['rgb_map'] = rgb_map.detach().clone(); captured['weights'] = weights.detach().clone()
            return rgb_map, {'raw': raw, 'z_vals': z_vals, 'valid_samples': valid_samples, 'weights': weights}
        fake.render = render
        nr.NerfRunner.train_loop(fake, batch)
        sv(f'ref_py_train_loop_{tag}.npz', batch=batch, z_vals=z_vals, valid_samples=valid_samples, raw=raw,
           loss=captured['loss'], rgb_map=captured['rgb_map'], weights=captured['weights'], draw=raw.grad,
           trunc=trunc, sc_factor=sc, first_frame_weight=ffw)

    # ---- 7. camera rays + transform_pts + DataLoader order (layout contracts)
    K = np.array([[600., 0, 320], [0, 600., 240], [0, 0, 1]])
    dirs = nh.get_camera_rays_np(8, 10, K)
    tfm = torch.randn(5, 4, 4, generator=g)
    pts = torch.randn(5, 3, generator=g)
    tp = U.transform_pts(pts, tfm)
    U.set_seed(0)
    dl = nr.DataLoader(rays=torch.arange(23).float().reshape(-1, 1), batch_size=5)
    torch.Tensor.cuda = lambda self, *a, **k: self            # DataLoader.__next__ calls .cuda()
    order = [next(dl).reshape(-1).numpy().copy() for _ in range(9)]
    sv('ref_py_misc.npz', dirs=dirs, K=K, tf=tfm, pts=pts, tp=tp, dl_order=np.stack(order))

    # ---- 8. tool.py: scene normalisation (tool.py:18-39) and depth back-projection (Utils.py:219-231) of the reference itself
    import importlib
    tool = importlib.import_module('tool')
    rng = np.random.default_rng(9)
    cloud = np.concatenate([rng.normal([0.02, -0.01, 0.55], [0.04, 0.05, 0.09], size=(3000, 3)),      # the object
                            rng.normal([0.40, 0.30, 0.90], 0.01, size=(60, 3))])                       # a detached blob DBSCAN must drop
    t0, s0, k0 = tool.compute_translation_scales(cloud.copy(), cluster=True, eps=0.06, min_samples=1)
    t1, s1, k1 = tool.compute_translation_scales(cloud.copy(), cluster=False)
    depth = (rng.random((12, 16)) * 1.5).astype(np.float32)
    depth[depth < 0.3] = 0.05
    xyz = U.depth2xyzmap(depth, K)
    sv('ref_py_tool.npz', cloud=cloud, translation_cluster=t0, sc_cluster=s0, keep_cluster=k0, translation_all=t1, sc_all=s1, keep_all=k1,
       depth=depth, K=K, xyz=xyz, glcam_in_cvcam=U.glcam_in_cvcam)
    print('golden CPU fixtures written to', HERE)


if __name__ == '__main__':
    main()
