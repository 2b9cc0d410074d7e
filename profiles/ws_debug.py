"""Debug aid: run one fused step with two AMP implementations on the same inputs and print the relative difference of every output."""
import sys, os
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import helpers
from bundlesdf_b200 import _lib

def run(impl, scene, t_rand):
    lib = _lib.load()
    old = lib.nof_set_amp_impl(impl)
    res = helpers.run_fused_step(scene, amp=True, t_rand=t_rand, loss_scale=1024.0)
    lib.nof_set_amp_impl(old)
    return res

def rel(a, b):
    a = a.float().cpu().numpy().astype(np.float64); b = b.float().cpu().numpy().astype(np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)

cases = [(4, 128, 14, 32, 32, 64, 0), (16, 256, 12, 64, 64, 48, 0), (16, 256, 12, 128, 64, 18, 2), (4, 128, 14, 16, 16, 30, 0)]
if len(sys.argv) > 1:
    cases = [cases[int(sys.argv[1])]]
for (L, finest, log2T, S_occ, S_d, N, ff) in cases:
    cfg = helpers.make_cfg(L, finest, log2T, S_occ, S_d, ff=ff)
    scene = helpers.make_scene(n_frames=4, N=N, cfg=cfg)
    t_rand = np.random.default_rng(11).random((N, S_occ + S_d), dtype=np.float32)
    a = run(2, scene, t_rand)
    b = run(1, scene, t_rand)
    print('case', (L, S_occ + S_d, N, ff), 'found_inf', a['found_inf'].item(), b['found_inf'].item())
    for k in ('raw', 'rgb_map', 'weights', 'losses', 'grad_table', 'grad_mlp', 'grad_tf', 'grad_pose'):
        print(f'  {k:12s} rel diff {rel(a[k], b[k]):.3e}   |ws| {float(a[k].abs().max()):.4e} |ref| {float(b[k].abs().max()):.4e}')
    for k in a['grad_mlp_named']:
        print(f'    {k:22s} {rel(a["grad_mlp_named"][k], b["grad_mlp_named"][k]):.3e}')
    if ff:
        print('  grad_feat', rel(a['grad_feat'], b['grad_feat']))
    gt_a, gt_b = a['grad_table'].cpu().numpy(), b['grad_table'].cpu().numpy()
    offs = scene['params']['offsets']
    for l in range(L):
        sa, sb = gt_a[offs[l]:offs[l + 1]], gt_b[offs[l]:offs[l + 1]]
        print(f'    level {l}: max|ws| {np.abs(sa).max():.3e} max|ref| {np.abs(sb).max():.3e} diff {np.abs(sa - sb).max():.3e} nnz {int((sa != 0).sum())} {int((sb != 0).sum())}')
