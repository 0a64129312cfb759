"""CPU-only, world_size=2 (gloo): the N>1 path of bench.py shards stereo pairs by rank with NO data-path
collective; the only communication is the barrier + max-over-ranks of the timing.  This test exercises that
host logic (rank -> pair assignment, barrier, MAX all-reduce, rank-0 aggregation) without a GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard_pairs(n_pairs, rank, world):
    """Contiguous shard of the pair indices owned by `rank` (SceneFlow config: batch 64 over 1/2/4/8 GPUs)."""
    per = (n_pairs + world - 1) // world
    return list(range(rank * per, min(n_pairs, (rank + 1) * per)))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_pairs(64, rank, world)
    dist.barrier()
    # each rank "times" its shard; whole-job time = max over ranks, throughput = all pairs / that time
    t = torch.tensor([10.0 + rank, float(len(mine))], dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((mine, float(tmax[0]), float(tsum[1])))
    else:
        q.put((mine, None, None))
    dist.destroy_process_group()


def test_pairs_shard_without_overlap_and_timing_is_max_over_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(i for r in res for i in r[0])
    assert owned == list(range(64))                      # every pair exactly once
    tmax = [r[1] for r in res if r[1] is not None][0]
    total = [r[2] for r in res if r[2] is not None][0]
    assert tmax == 11.0 and total == 64.0


def test_shard_sizes_for_the_bench_world_sizes():
    for world in (1, 2, 4, 8):
        sizes = [len(shard_pairs(64, r, world)) for r in range(world)]
        assert sum(sizes) == 64 and max(sizes) - min(sizes) == 0
