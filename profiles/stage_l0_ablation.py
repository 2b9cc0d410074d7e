"""Ablation for the north-star clause 'TMA staging of the hash table': level 0 of the fp16 table (17^3 entries, 19.7 KB) staged into shared
memory by a TMA bulk copy in tc::step_tc_kernel and gathered from there (build with -DNOF_EXP_STAGE_L0 into bundlesdf_b200/lib_stage0).
Prints the fused-step launch time at C2 and a parity check of the two builds against each other.
    NOF_LIB=bundlesdf_b200/lib_stage0/libnof_sm100.so python profiles/stage_l0_ablation.py   vs   python profiles/stage_l0_ablation.py"""
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
