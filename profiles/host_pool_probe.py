"""train_steps(host_pool=...) against the device-resident pool: 20-step blocks at C2 (sync on both sides of every block, like bench.py), plus the
host-side staging time alone. Measured on the B200: device pool 5.42 ms per 20 steps; host pool 6.45 ms with the prologue reading the pinned
staging block directly (shipped), 6.52 ms with one H2D + one D2H copy node per step on separate streams, 6.74 ms with both on one stream;
staging 10 batches on the host 0.27 ms (overlapped). The per-step public loop (train_loop + H2D/D2H, bench.py's e2e) is at 6.3 ms."""
import os, sys, time
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
c = bench.CONFIGS['C2']
dev = torch.device('cuda', 0)

def run(debug, host):
    runner, seq = bench.build_runner(c, 0, dev, eager=False)
    pool = runner.rays.cpu().pin_memory() if host else None
    runner.train_steps(5)
    while runner.global_step % 10 != 1:
        runner.train_steps(1)
    def block():
        runner.train_steps(20, host_pool=pool)
        if host:
            runner.collect_host_losses()
    for _ in range(4):
        block()
    ts = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(15):
        torch.cuda.synchronize(); ev0.record(); block(); ev1.record(); torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    print(f'host={host} debug={sorted(debug)}: 20-step block median {np.median(ts):.3f} ms  min {min(ts):.3f}', flush=True)
    if host and not debug:
        h = runner._host
        t0 = time.perf_counter()
        for _ in range(10):
            runner._stage_host(h, 0, 0, 10, pool)
        print(f'  staging 10 batches on the host: {(time.perf_counter() - t0) * 100:.3f} ms', flush=True)

run((), False)
run((), True)



