"""Per-kernel timing of one train step (CUDA events, warm caches, launches back to back) for a bench config.
Usage: [NOF_LIB=/path/to/variant.so] python profiles/kernel_time.py [--config C2] [--frames 24]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bundlesdf_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--frames', type=int, default=24)
ap.add_argument('--reps', type=int, default=50)
ap.add_argument('--fp32', action='store_true', help='cfg amp: false (fp32 policy kernel)')
args = ap.parse_args()
if args.fp32:
    _mk = bench.make_cfg
    bench.make_cfg = lambda c: dict(_mk(c), amp=False)
c = dict(bench.CONFIGS[args.config]); c['frames'] = min(c['frames'], args.frames)
torch.cuda.set_device(0)
runner, seq = bench.build_runner(c, 0, torch.device('cuda', 0), eager=True)
cfg = runner.cfg; sc = cfg['sc_factor']


def timeit(fn, n=args.reps):
    for _ in range(3):
        fn()
    return bench.time_steps(fn, n) / n * 1e6


batch = next(runner.data_loader)
runner._forward_backward(batch)
b = runner._step_buf; sb = b['sb']; pa = runner.models['pose_array']
res = {'lib': os.environ.get('NOF_LIB', 'default'), 'config': args.config, 'N': c['N'], 'S': c['S_occ'] + c['S_d']}
res['step_fused_us'] = timeit(sb.launch)
res['ray_march_us'] = timeit(lambda: ops.ray_march(batch, b['tf'], runner.octree_m.occ_bits, runner.octree_m.level, cfg['N_samples'],
                                                   cfg['N_samples_around_depth'], runner.get_truncation(), cfg['near'] * sc, cfg['far'] * sc,
                                                   cfg['neg_trunc_ratio'], perturb=True, z_vals=b['z_vals']))
res['pose_fwd_us'] = timeit(lambda: ops.pose_forward(pa.data.data, runner.c2w_array, cfg['max_trans'] * sc, cfg['max_rot'], out=b['tf']))
res['pose_bwd_us'] = timeit(lambda: ops.pose_backward(pa.data.data, runner.c2w_array, b['grad_tf'], runner.adam_segs['pose']['grad'].view(-1, 6),
                                                      cfg['max_trans'] * sc, cfg['max_rot'], runner.amp_scaler.state))
res['gather_us'] = timeit(lambda: next(runner.data_loader))
for s in runner.adam_segs.values():
    s['grad'].zero_()
runner.amp_scaler.found_inf.zero_()
res['adam_us'] = timeit(runner._optimizer_step)


def step():
    runner.train_loop(next(runner.data_loader)); runner.global_step += 1


res['step_eager_us'] = timeit(step, 100)
runner.cfg['use_cuda_graph'] = True
for _ in range(5):
    step()
res['step_graph_us'] = timeit(step, 100)
res['rays_per_s_graph'] = c['N'] / res['step_graph_us'] * 1e6
print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}))
