"""Fused-step launch time at C2 of a library variant, plus a checksum of its results (variants of one kernel must agree).
    python -m bundlesdf_b200.build                                                      # default build -> bundlesdf_b200/lib
    NOF_BUILD_DIR=bundlesdf_b200/lib_x NOF_EXTRA_FLAGS=-DNOF_EXP_... python -m bundlesdf_b200.build
    python profiles/variant_time.py ; NOF_LIB=bundlesdf_b200/lib_x/libnof_sm100.so python profiles/variant_time.py
Used for: -DNOF_EXP_STAGE_L0 (level 0 of the table staged in shared memory by TMA) and the component ablations -DNOF_EXP_NO_GATHER /
NO_SCATTER / NO_RED / NO_WGRAD / NO_TC (one part of tc::step_tc_kernel compiled out: what the rest costs). Results in profiles/README.md."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
c = bench.CONFIGS['C2']
runner, seq = bench.build_runner(c, 0, torch.device('cuda', 0), eager=True)
for _ in range(5):
    runner.train_loop(next(runner.data_loader)); runner.global_step += 1
runner.synchronize_parameters()
torch.manual_seed(0)
batch = runner.rays[torch.arange(0, c['N'] * 37, 37, device='cuda') % len(runner.rays)].contiguous()
for s_ in runner.adam_segs.values():
    s_['grad'].zero_()
runner._forward_backward(batch, t_rand=torch.rand(c['N'], 128, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1)))
sb = runner._step_buf['sb']
torch.cuda.synchronize()
print('grad_table checksum', float(runner.adam_segs['table']['grad'].double().abs().sum()), 'loss', float(runner._step_buf['losses'][0]))
for _ in range(3):
    sb.launch()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(50):
    sb.launch()
ev1.record(); torch.cuda.synchronize()
print(f'{os.environ.get("NOF_LIB", "default build")}: fused step launch {ev0.elapsed_time(ev1) / 50 * 1e3:.1f} us')
