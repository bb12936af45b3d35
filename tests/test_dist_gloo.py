"""N>1 host logic on CPU: env sharding, max-over-ranks timing and the metrics all-gather,
world_size 2 over gloo (the GPU path uses the same functions over NCCL)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vlfm_b200.utils.dist import aggregate_throughput, env_shard, gather_metrics, max_over_ranks, owner_of


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    envs = env_shard(rank, world, 3)
    ms = 10.0 + 5.0 * rank  # rank 1 is the slow one
    mx = max_over_ranks(ms, dev)
    g = gather_metrics([rank, ms, float(sum(envs))], dev)
    dist.barrier()
    q.put((rank, envs, mx, g))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4, 5]            # disjoint contiguous shards
    assert all(owner_of(e, 3) == r for r, envs, _, _ in res for e in envs)
    assert res[0][2] == res[1][2] == 15.0                                 # max over ranks
    assert res[0][3] == res[1][3] == [[0.0, 10.0, 3.0], [1.0, 15.0, 12.0]]
    assert aggregate_throughput(2, 3, 100, 15.0) == 2 * 3 * 100 / 0.015


def test_single_rank_paths():
    assert max_over_ranks(3.5, torch.device("cpu")) == 3.5
    assert gather_metrics([1, 2], torch.device("cpu")) == [[1.0, 2.0]]
