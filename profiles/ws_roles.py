"""Role profile of the warp-specialised step kernel (CTA 0): per warp, cycles in its role loop and cycles of those spent waiting on
mbarriers. Needs the library built with -DNOF_WS_PROF:
    NOF_BUILD_DIR=bundlesdf_b200/lib_prof NOF_EXTRA_FLAGS=-DNOF_WS_PROF python -m bundlesdf_b200.build
    NOF_LIB=bundlesdf_b200/lib_prof/libnof_sm100.so NOF_AMP_IMPL=ws python profiles/ws_roles.py [C2|C3|C5]"""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench

name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
c = bench.CONFIGS[name]
runner, seq = bench.build_runner(c, 0, torch.device('cuda', 0), eager=True)
for _ in range(5):
    runner.train_loop(next(runner.data_loader)); runner.global_step += 1
runner.synchronize_parameters()
batch = next(runner.data_loader)
runner._forward_backward(batch)
sb = runner._step_buf['sb']
for _ in range(3):
    sb.launch()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    sb.launch()
ev1.record(); torch.cuda.synchronize()
print(f'{name}: fused step launch (pack + step) {ev0.elapsed_time(ev1) / 20 * 1e3:.1f} us')
ws = sb.keep['workspace']
pr = ws[24 * 1024:24 * 1024 + 24 * 16].view(torch.int64).cpu().numpy().reshape(24, 2)
roles = ['EPI'] * 8 + ['WGRAD'] * 4 + ['GATHER'] * 4 + ['SCATTER'] * 4 + ['MMA'] + ['idle'] * 3
for w in range(24):
    tot, wait = pr[w]
    if tot:
        print(f'warp {w:2d} {roles[w]:8s} loop {tot:9d} cyc   waiting {wait:9d} cyc ({100.0 * wait / tot:5.1f} %)   busy {tot - wait:9d}')
