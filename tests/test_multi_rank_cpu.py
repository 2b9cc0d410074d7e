"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path (independent sequences, weak scaling: per-rank seeds,
barrier, MAX-over-ranks time, SUM of units). No data-path collective exists in this design."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import bench
    from bundlesdf_b200 import synthetic as syn
    seed = bench.rank_seed(7, rank)
    seq = syn.make_sequence(2, H=48, W=64, seed=seed, depth_noise_m=0.001)
    dist.barrier()
    # rank r "processed" 1000 units in (1+r) seconds: whole-job value must be 2000 / 2 s
    value, t = bench.aggregate_throughput(1000, 1.0 + rank, world)
    # bench.py's timed region: three K-step blocks; per block the MAX over ranks, the job's block time is the median over blocks
    blocks = bench.max_over_ranks([1.0 + rank, 3.0 - rank, 1.5], world)
    assert blocks == [2.0, 3.0, 1.5], blocks
    assert bench.whole_job_value(1000, blocks, world) == pytest.approx(2 * 1000 / 2.0)
    q.put((rank, seed, float(seq['depths'].sum()), value, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_weak_scaling_aggregation():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, d0, v0, t0), (r1, s1, d1, v1, t1) = res
    assert (s0, s1) == (7, 8) and d0 != d1            # independent sequences (different depth noise) per rank
    assert v0 == v1 == pytest.approx(1000.0) and t0 == t1 == pytest.approx(2.0)
