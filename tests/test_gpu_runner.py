"""GPU: the NerfRunner drop-in (reference API surface) end to end on a small synthetic sequence."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import nof_oracle as O

pytestmark = pytest.mark.gpu


def _runner(amp, n_frames=5, N=256, ff=0, **over):
    from bundlesdf_b200 import synthetic as syn
    from bundlesdf_b200.nerf_runner import NerfRunner
    seq = syn.make_sequence(n_frames, H=120, W=160, device='cuda', seed=3, pose_noise=True)
    cfg = syn.default_cfg(N_rand=N, N_samples=64, N_samples_around_depth=64, num_levels=16, finest_res=256, log2_hashmap_size=14, amp=amp,
                          sc_factor=seq['sc_factor'], translation=seq['translation'].tolist(), n_step=60, frame_features=ff, **over)
    r = NerfRunner(cfg, seq['images'], seq['depths'], seq['masks'], None, seq['poses'], seq['K'], build_octree_pcd=syn.PointCloud(seq['pcd_normalized']))
    return r, seq


@pytest.mark.parametrize('amp', [True, False])
def test_single_step_gradients_match_oracle_through_runner(amp):
    r, seq = _runner(amp, ff=2)
    with torch.no_grad():
        r.models['embed_fn'].embeddings.uniform_(-0.3, 0.3)
        if r.table_f16 is not None:
            r.table_f16.copy_(r.table)
        r.models['pose_array'].data.normal_(0, 0.1)
    batch = next(r.data_loader)
    N, S = batch.shape[0], 128
    t_rand = torch.rand(N, S, device='cuda')
    b = r._forward_backward(batch, t_rand=t_rand)
    torch.cuda.synchronize()
    scale = r.amp_scaler.get_scale()
    # oracle on the runner's own parameters
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in r.models['model'].state_dict().items()}
    P = dict(sd)
    P['embeddings'] = r.models['embed_fn'].embeddings.detach().cpu().clone().requires_grad_(True)
    P['offsets'] = r.models['embed_fn'].offsets.cpu().numpy()
    P['S'] = float(np.log2(r.models['embed_fn'].per_level_scale)); P['H'] = 16
    P['pose_data'] = r.models['pose_array'].data.detach().cpu().clone().requires_grad_(True)
    P['feature_data'] = r.models['feature_array'].data.detach().cpu().clone().requires_grad_(True)
    ref = O.forward_step(P, batch.cpu(), r.c2w_array.cpu(), r.octree_m.occ.cpu().numpy(), r.cfg, half=amp, z_vals=b['z_vals'].cpu())
    ref['loss'].backward()
    gtol = 3e-2 if amp else 2e-3
    rel = lambda a, w: np.abs(a - w).max() / max(np.abs(w).max(), 1e-30)
    assert rel(r.models['embed_fn'].embeddings.grad.cpu().numpy() / scale, P['embeddings'].grad.numpy()) < gtol
    for k, p in r.models['model'].named_parameters():
        assert rel(p.grad.cpu().numpy() / scale, sd[k].grad.numpy()) < gtol, k
    assert rel(r.models['pose_array'].data.grad.cpu().numpy() / scale, P['pose_data'].grad.numpy()) < 2 * gtol   # scaled like every segment
    assert rel(r.models['feature_array'].data.grad.cpu().numpy() / scale, P['feature_data'].grad.numpy()) < gtol
    m = r.get_metrics()
    assert abs(m['loss'] - float(ref['loss'].detach())) <= (5e-3 if amp else 2e-4) * abs(float(ref['loss'].detach()))


@pytest.mark.parametrize('amp', [True, False])
def test_training_reduces_loss_and_recovers_geometry(amp):
    r, seq = _runner(amp, n_frames=6, N=512)
    r.cfg['i_print'] = 999999
    losses = []
    for it in range(120):
        r.train_loop(next(r.data_loader))
        r.global_step += 1
        if it in (0, 119):
            losses.append(r.get_metrics())
    assert np.isfinite(losses[1]['loss'])
    geo0 = losses[0]['sdf_loss'] + losses[0]['fs_loss']
    geo1 = losses[1]['sdf_loss'] + losses[1]['fs_loss']
    assert geo1 < 0.2 * geo0, (geo0, geo1)
    assert losses[1]['rgb_loss'] < losses[0]['rgb_loss'] and losses[1]['loss'] < 0.6 * losses[0]['loss'], \
        {k: (losses[0][k], losses[1][k]) for k in ('loss', 'rgb_loss', 'fs_loss', 'sdf_loss')}
    # SDF sign: a point at the object's centre must be inside (negative), a far corner of an occupied cell positive-ish
    sdf, _ = r.run_network_density(torch.tensor([[0.0, 0.0, 0.0]], device='cuda'))
    assert sdf.item() < 0.0
    assert r.adam_step_count.item() == 120
    assert r.amp_scaler.found_inf.item() == 0
    # extract_mesh (nerf_runner.py:1351-1409) on the trained field: a closed surface inside the bounding box, roughly jug-sized
    mesh, sigma, pts = r.extract_mesh(voxel_size=0.01, isolevel=0.0, return_sigma=True)
    assert mesh is not None and len(mesh.faces) > 200
    v = np.asarray(mesh.vertices)
    assert np.abs(v).max() <= 1.0
    from bundlesdf_b200.mesh import TriMesh
    assert TriMesh(v, np.asarray(mesh.faces)).is_watertight
    extent = (v.max(0) - v.min(0)) / r.cfg['sc_factor']                  # metres: the synthetic jug is 0.10 x 0.10 x ~0.23
    assert 0.05 < extent.min() and extent.max() < 0.45, extent


def test_train_api_and_checkpoint_roundtrip(tmp_path):
    r, seq = _runner(True, n_frames=4, N=256)
    r.cfg['n_step'] = 12
    r.N_iters = 13
    r.cfg['save_dir'] = str(tmp_path)
    r.train()                                    # reference API: n_step+1 iterations (nerf_runner.py:855-863)
    assert r.global_step == 13
    ck = os.path.join(tmp_path, 'model_latest.pth')
    r.save_weights(ck, r.models)
    before = {k: v.clone() for k, v in r.models['model'].state_dict().items()}
    emb = r.models['embed_fn'].embeddings.detach().clone()
    mats = r.models['pose_array'].get_matrices(np.arange(4))
    assert mats.shape == (4, 4, 4) and torch.allclose(mats[0], torch.eye(4, device='cuda'))
    r2, _ = _runner(True, n_frames=4, N=256)
    r2.load_weights(ck)
    for k, v in r2.models['model'].state_dict().items():
        assert torch.equal(v, before[k])
    assert torch.equal(r2.models['embed_fn'].embeddings.detach(), emb)
    assert torch.equal(r2.table_f16, emb.half())
    assert r2.adam_step_count.item() == r.adam_step_count.item()
    # the two runners now take the same next step
    batch = next(r.data_loader)
    t_rand = torch.rand(batch.shape[0], 128, device='cuda')
    r.train_loop(batch, t_rand=t_rand); r2.global_step = r.global_step; r2.train_loop(batch, t_rand=t_rand)
    torch.cuda.synchronize()
    np.testing.assert_allclose(r2.mlp_flat.cpu().numpy(), r.mlp_flat.cpu().numpy(), rtol=1e-4, atol=1e-6)
    ckpt = torch.load(ck, weights_only=False)
    assert set(['model', 'embed_fn', 'pose_array', 'optimizer', 'global_step', 'octree']) <= set(ckpt.keys())
    assert set(ckpt['embed_fn'].keys()) == {'embeddings', 'offsets'} and set(ckpt['pose_array'].keys()) == {'data'}


def test_add_new_frames_and_extract_sigma():
    from bundlesdf_b200 import synthetic as syn
    r, seq = _runner(True, n_frames=3, N=256)
    n0 = r.rays.shape[0]
    seq2 = syn.make_sequence(5, H=120, W=160, device='cuda', seed=3, pose_noise=True, sc_factor=seq['sc_factor'])
    r.add_new_frames(seq2['images'][3:], seq2['depths'][3:], seq2['masks'][3:], None, seq2['poses'], new_pcd=syn.PointCloud(seq2['pcd_normalized']))
    assert r.rays.shape[0] > n0 and r.rays.is_cuda and len(r.images) == 5
    assert r.models['pose_array'].data.shape[0] == 5
    for _ in range(5):
        r.train_loop(next(r.data_loader)); r.global_step += 1
    mesh, sigma, pts = r.extract_mesh(voxel_size=0.02, return_sigma=True)
    assert sigma.ndim == 3 and np.isfinite(sigma).all()


def test_render_contract():
    r, _ = _runner(False, n_frames=3, N=128)
    batch = next(r.data_loader)
    rgb, extras = r.render(batch, depth=batch[:, 6], perturb=False)
    assert rgb.shape == (128, 3) and extras['raw'].shape == (128, 128, 4) and extras['z_vals'].shape == (128, 128)
    assert extras['valid_samples'].dtype == torch.bool
    assert float(r.adam_segs['table']['grad'].abs().max()) == 0.0


@pytest.mark.parametrize('amp,defer', [(False, False), (True, False), (True, True)])
def test_training_trajectory_matches_oracle(amp, defer):
    """same init, same batches, same jitter -> the loss trajectory and the optimised parameters of the CUDA path
    follow the CPU oracle (torch autograd + the oracle's Adam) over 25 optimizer steps. `defer`: the table's Adam pass of step k
    runs at the start of step k+1 on a side stream (cfg defer_table_update) — same trajectory."""
    r, seq = _runner(amp, n_frames=4, N=192, defer_table_update=defer)
    steps = 25
    enc = r.models['embed_fn']
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in r.models['model'].state_dict().items()}
    P['embeddings'] = enc.embeddings.detach().cpu().clone().requires_grad_(True)
    P['offsets'] = enc.offsets.cpu().numpy(); P['S'] = float(np.log2(enc.per_level_scale)); P['H'] = 16
    P['pose_data'] = r.models['pose_array'].data.detach().cpu().clone().requires_grad_(True)
    P['feature_data'] = None
    leaves = [P['embeddings']] + [P[k] for k in P if 'net' in k] + [P['pose_data']]
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in leaves]
    c2w, occ = r.c2w_array.cpu(), r.octree_m.occ.cpu().numpy()
    gen = torch.Generator().manual_seed(5)
    got, want = [], []
    lr = 0.01
    for it in range(steps):
        batch = next(r.data_loader)
        t_rand = torch.rand(batch.shape[0], 128, generator=gen)
        b = r.train_loop(batch, t_rand=t_rand.cuda())
        got.append(r.get_metrics())
        zv = b['z_vals'].cpu()
        out = O.forward_step(P, batch.cpu(), c2w, occ, r.cfg, half=amp, z_vals=zv)
        for p in leaves:
            p.grad = None
        out['loss'].backward()
        with torch.no_grad():
            for p, (m, v) in zip(leaves, state):
                O.adam_update(p, p.grad, m, v, it + 1, lr)
        if it % 10 == 0 and it > 0:                        # schedule_lr (nerf_runner.py:762-763,579-583), applies from the next step
            lr = O.lr_at(r.cfg, 0.01, it)
        want.append({k: float(out[k].detach()) for k in ('loss', 'rgb_loss', 'fs_loss', 'sdf_loss')})
        r.global_step += 1
    tol = 0.08 if amp else 0.02
    for it in range(steps):
        for k in ('loss', 'rgb_loss', 'sdf_loss'):
            assert abs(got[it][k] - want[it][k]) <= tol * max(abs(want[it][k]), 1e-3), (it, k, got[it][k], want[it][k])
    assert want[-1]['sdf_loss'] < 0.8 * want[0]['sdf_loss']
    # Adam (eps=1e-15) turns a noise-sized gradient into a full +-lr step, so single entries may differ by several lr after
    # 25 steps; compare in the L2 sense.
    rel = lambda a, w: np.linalg.norm(a - w) / max(np.linalg.norm(w), 1e-30)
    assert rel(r.models['model'].state_dict()['color_net.4.weight'].cpu().numpy(), P['color_net.4.weight'].detach().numpy()) < (0.1 if amp else 0.03)
    assert rel(r.models['pose_array'].data.detach().cpu().numpy(), P['pose_data'].detach().numpy()) < (0.2 if amp else 0.1)
    r.synchronize_parameters()
    assert r.adam_step_count.item() == steps
    tab = rel(enc.embeddings.detach().cpu().numpy(), P['embeddings'].detach().numpy())
    assert tab < (0.35 if amp else 0.15), tab


def test_deferred_table_update_graph_path_matches_plain():
    """cfg defer_table_update with CUDA-graph replay (two graph variants: with / without a pending update) against the plain step
    order: same batches -> same losses within the noise of the atomics, same Adam step count after synchronize_parameters()."""
    ra, _ = _runner(True, n_frames=4, N=256)
    rb, _ = _runner(True, n_frames=4, N=256, defer_table_update=True)
    la, lb = [], []
    for it in range(40):
        batch = next(ra.data_loader)
        next(rb.data_loader)
        ra.train_loop(batch)
        rb.train_loop(batch.clone())
        ra.global_step += 1
        rb.global_step += 1
        if it == 17:                                        # an external reader in the middle: flushes, next step has nothing pending
            sa, _ = ra.run_network_density(torch.zeros(1, 3, device='cuda'))
            sb_, _ = rb.run_network_density(torch.zeros(1, 3, device='cuda'))
            assert abs(sa.item() - sb_.item()) < 5e-2 * max(abs(sa.item()), 0.05)
        if it % 5 == 4:
            la.append(ra.get_metrics()['loss'])
            lb.append(rb.get_metrics()['loss'])
    rb.synchronize_parameters()
    assert rb.adam_step_count.item() == ra.adam_step_count.item() == 40
    assert rb.tick.item() == ra.tick.item() == 40 and rb.march_tick.item() == ra.march_tick.item() == 40
    np.testing.assert_allclose(lb, la, rtol=0.05)
    rel = lambda a, w: float((a - w).norm() / w.norm())
    assert rel(rb.table, ra.table) < 0.2 and rel(rb.mlp_flat, ra.mlp_flat) < 0.1


def test_train_steps_blocks_match_per_step_loop():
    """NerfRunner.train_steps: batches gathered on the device at the loader's cursor and runs of 10 steps replayed as ONE CUDA graph
    (bench.py's timed region, and what train() uses) against the per-step public loop train_loop(next(data_loader)): same
    permutation -> same batches, same sampler ticks -> the same trajectory within the noise of the atomics."""
    ra, _ = _runner(True, n_frames=4, N=256)
    rb, _ = _runner(True, n_frames=4, N=256)
    from bundlesdf_b200.nerf_runner import set_seed
    n = 47                                                  # 2 eager + singles up to step 10 + 10-step blocks + a tail; crosses an epoch
    set_seed(0)                                             # the mid-run reshuffle draws from the global CPU RNG: same stream for both
    for it in range(n):
        ra.train_loop(next(ra.data_loader))
        ra.global_step += 1
    set_seed(0)
    rb.train_steps(n)
    assert rb.global_step == ra.global_step == n
    assert rb.adam_step_count.item() == ra.adam_step_count.item() == n
    assert rb.march_tick.item() == ra.march_tick.item() == n
    assert rb.data_loader.pos == ra.data_loader.pos
    assert torch.equal(rb.data_loader.batch_ray_ids, ra.data_loader.batch_ray_ids)
    assert any(k[0] == 'blk' and k[2] == 10 for k in rb._graph), list(rb._graph)
    assert float(rb.lr_dev[0]) == pytest.approx(float(ra.lr_dev[0]), rel=1e-6) and float(rb.lr_dev[0]) < 0.01
    la, lb = ra.get_metrics(), rb.get_metrics()
    for k in ('loss', 'rgb_loss', 'sdf_loss', 'fs_loss'):
        assert lb[k] == pytest.approx(la[k], rel=0.05), k
    rel = lambda a, w: float((a - w).norm() / w.norm())
    assert rel(rb.table, ra.table) < 0.2 and rel(rb.mlp_flat, ra.mlp_flat) < 0.1


@pytest.mark.parametrize('defer', [False, True])
def test_train_steps_from_a_host_pool_match_the_device_pool(defer):
    """train_steps(n, host_pool=pinned copy of the ray pool): batches gathered on the host in the loader's order and uploaded by one H2D node
    per step inside the block graph, losses returned by one D2H node per step — same batches, same ticks, same trajectory as the
    device-resident pool (within the noise of the atomics), and one loss row per step that ran through a graph."""
    from bundlesdf_b200.nerf_runner import set_seed
    ra, _ = _runner(True, n_frames=4, N=256, defer_table_update=defer)
    rb, _ = _runner(True, n_frames=4, N=256, defer_table_update=defer)
    n = 53                                                  # eager steps, singles, 10-step blocks, a tail; crosses an epoch boundary
    set_seed(0)
    ra.train_steps(n)
    set_seed(0)
    pool = rb.rays.cpu().pin_memory()
    rb.train_steps(31, host_pool=pool)
    first = rb.collect_host_losses()
    rb.train_steps(n - 31, host_pool=pool)
    rest = rb.collect_host_losses()
    ra.synchronize_parameters(); rb.synchronize_parameters()
    assert rb.global_step == ra.global_step == n
    assert rb.adam_step_count.item() == ra.adam_step_count.item()
    assert rb.march_tick.item() == ra.march_tick.item() == n
    assert rb.data_loader.pos == ra.data_loader.pos
    assert any(k[0] == 'hblk' and k[2] == 10 for k in rb._graph), list(rb._graph)
    rows = np.concatenate([first, rest], 0)
    assert len(rows) == n - 2 and np.isfinite(rows).all()   # the first two steps run eagerly (buffer allocation), the rest through graphs
    la, lb = ra.get_metrics(), rb.get_metrics()
    assert lb['loss'] == pytest.approx(la['loss'], rel=0.05)
    assert rows[-1, 0] == pytest.approx(float(rb._step_buf['losses'][0]), rel=1e-6)      # the last row IS the last step's loss
    rel = lambda a, w: float((a - w).norm() / w.norm())
    assert rel(rb.table, ra.table) < 0.2 and rel(rb.mlp_flat, ra.mlp_flat) < 0.1


@pytest.mark.parametrize('decay', ['linear', 'exp'])
def test_annealed_truncation_runs_inside_the_graph_blocks(decay):
    """trunc_decay_type linear | exp (nerf_runner.py:663-676): the per-step truncation lives in a device table indexed by a device step
    counter (nof_step_prologue), so annealing no longer forces one launch sequence per step. Checked: the scalar the kernels read equals
    get_truncation(step) after every stretch; one fused step in the middle of the schedule matches the oracle evaluated at that truncation; the blocks and the
    per-step loop walk the same trajectory."""
    from bundlesdf_b200.nerf_runner import set_seed
    over = dict(trunc_decay_type=decay, trunc_start=0.03, trunc=0.01)
    ra, _ = _runner(True, n_frames=4, N=256, **over)
    rb, _ = _runner(True, n_frames=4, N=256, **over)
    assert ra.get_truncation(0) > 2.5 * ra.get_truncation(60)
    n = 41
    set_seed(0)
    for _ in range(n):
        ra.train_loop(next(ra.data_loader))
        ra.global_step += 1
    set_seed(0)
    rb.train_steps(n)
    assert any(k[0] == 'blk' and k[2] == 10 for k in rb._graph), list(rb._graph)
    for r in (ra, rb):
        ts = r._trunc_schedule()
        assert int(ts['gstep'].item()) == n
        assert float(ts['out'].item()) == np.float32(r.get_truncation(n - 1))        # what the last step's kernels read
    la, lb = ra.get_metrics(), rb.get_metrics()
    assert lb['loss'] == pytest.approx(la['loss'], rel=0.05)
    # one step at a given point of the schedule against the oracle at that truncation
    r = rb
    g0 = 37 if decay == 'linear' else 5                      # a step where the schedule is still well above its final value
    r.global_step = g0
    r._sync_trunc_step()
    batch = next(r.data_loader)
    t_rand = torch.rand(batch.shape[0], 128, device='cuda')
    r.synchronize_parameters()
    b = r._forward_backward(batch, t_rand=t_rand)
    torch.cuda.synchronize()
    P = {k: v.detach().cpu().clone() for k, v in r.models['model'].state_dict().items()}
    P['embeddings'] = r.models['embed_fn'].embeddings.detach().cpu().clone()
    P['offsets'] = r.models['embed_fn'].offsets.cpu().numpy()
    P['S'] = float(np.log2(r.models['embed_fn'].per_level_scale)); P['H'] = 16
    P['pose_data'] = r.models['pose_array'].data.detach().cpu().clone()
    P['feature_data'] = None
    ref = O.forward_step(P, batch.cpu(), r.c2w_array.cpu(), r.octree_m.occ.cpu().numpy(), r.cfg, global_step=g0, half=True, z_vals=b['z_vals'].cpu())
    assert float(b['losses'][0]) == pytest.approx(float(ref['loss'].detach()), rel=5e-3)
    # the ray march read the same scalar: the around-depth samples span depth -/+ the truncation OF THAT STEP, wider than the final one
    tr37, tr_end = r.get_truncation(g0), r.get_truncation(r.cfg['n_step'])
    ok = (batch[:, 6] >= r.cfg['near'] * r.cfg['sc_factor']) & (batch[:, 6] <= r.cfg['far'] * r.cfg['sc_factor'])
    dz = (b['z_vals'][ok][:, 64:] - batch[ok, 6:7]).abs()
    assert float(dz.max()) <= tr37 * (1 + 1e-5) and float(dz.max()) > 1.5 * tr_end, (float(dz.max()), tr37, tr_end)

def test_pose_regulariser_follows_the_loss_scale():
    """pose_reg_weight > 0 under AMP: the regulariser's gradient is added to the (loss-scaled) pose gradient buffer multiplied by the
    loss scale, so that the single unscale inside nof_adam_step recovers d/dpose [pose_reg_weight * ||data[1:]||] (nerf_runner.py:748-758)."""
    r, _ = _runner(True, n_frames=4, N=128, pose_reg_weight=0.5)
    with torch.no_grad():
        r.models['pose_array'].data.normal_(0, 0.1)
    batch = next(r.data_loader)
    t_rand = torch.rand(batch.shape[0], 128, device='cuda')
    r.cfg['pose_reg_weight'] = 0.0
    r._forward_backward(batch, t_rand=t_rand)
    g0 = r.adam_segs['pose']['grad'].clone().view(-1, 6)
    for s_ in r.adam_segs.values():
        s_['grad'].zero_()
    r.cfg['pose_reg_weight'] = 0.5
    r._forward_backward(batch, t_rand=t_rand)
    g1 = r.adam_segs['pose']['grad'].clone().view(-1, 6)
    d = r.models['pose_array'].data.detach()[1:]
    want = 0.5 * d / d.norm()
    scale = r.amp_scaler.get_scale()
    got = (g1 - g0)[1:] / scale
    assert torch.allclose(got, want, rtol=2e-2, atol=2e-3 * float(want.abs().max())), (got, want)
    assert float((g1 - g0)[0].abs().max()) == 0.0


@pytest.mark.parametrize('amp', [True, False])
def test_run_network_density_normals_match_oracle(amp):
    """run_network_density(get_normals=True) (nerf_runner.py:1342-1345): sdf | d sdf / d x through the op-level grid encoder's dy_dx and
    autograd, against oracle.sdf_normals on the same parameters. fp32: 1e-4 of max|n|; AMP (fp16 table / dy_dx / activations): 3e-2."""
    r, seq = _runner(amp, n_frames=4, N=256)
    with torch.no_grad():
        r.models['embed_fn'].embeddings.uniform_(-0.5, 0.5)
        if r.table_f16 is not None:
            r.table_f16.copy_(r.table)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(3000, 3, generator=g) * 2 - 1) * 0.98
    out, ok = r.run_network_density(x.cuda(), get_normals=True)
    assert out.shape == (3000, 4) and bool(ok.all())
    plain, _ = r.run_network_density(x.cuda())
    P = {k: v.detach().cpu() for k, v in r.models['model'].state_dict().items()}
    P['embeddings'] = r.models['embed_fn'].embeddings.detach().cpu()
    P['offsets'] = r.models['embed_fn'].offsets.cpu().numpy()
    P['S'] = float(np.log2(r.models['embed_fn'].per_level_scale)); P['H'] = 16
    sdf, n = O.sdf_normals(P, x, torch.ones(3000, dtype=torch.bool), half=amp)
    rel = lambda a, w: np.abs(a - w).max() / max(np.abs(w).max(), 1e-30)
    e = (rel(out[:, 0].cpu().numpy(), sdf.detach().numpy()), rel(plain[:, 0].cpu().numpy(), sdf.detach().numpy()))
    assert e[0] < (5e-3 if amp else 1e-4) and e[1] < (5e-3 if amp else 1e-4), e
    # The SDF is continuous but its gradient jumps across cell faces: a probe within fp32 rounding of a face of one of the 16 levels
    # (3000 x 16 x 3 chances, ~1.5e-5 each at the finest level) may sit in the neighbouring cell in the other implementation. So the
    # normals are compared point-wise: at least 99.5 % of the probes within tol of max|n|, and the typical probe far inside it.
    nw = n.detach().numpy()
    err = np.abs(out[:, 1:].cpu().numpy() - nw).max(axis=1) / np.abs(nw).max()
    tol = 3e-2 if amp else 1e-4
    print('run_network_density vs oracle: sdf %.3g / %.3g, normals median %.3g, within tol %.4f, max %.3g' % (e + (np.median(err), (err < tol).mean(), err.max())))
    assert (err < tol).mean() > 0.995 and np.median(err) < tol / 5, (np.median(err), (err < tol).mean())
