"""N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1), batch sharding + timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scanobjectnn_b200.shard import max_over_ranks, rank_seed, shard_range, sum_over_ranks


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard_range(65, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    mx = max_over_ranks(10.0 + rank)
    total = sum_over_ranks(len(mine))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, gathered, mx, total, rank_seed(1001, rank)))


def test_world_size_2_sharding_and_reductions():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seeds = set()
    for rank, gathered, mx, total, seed in res:
        flat = [i for part in gathered for i in part]
        assert flat == list(range(65))                      # disjoint, complete, ordered
        assert abs(len(gathered[0]) - len(gathered[1])) <= 1
        assert mx == 11.0 and total == 65.0
        seeds.add(seed)
    assert len(seeds) == world


def test_shard_range_edges():
    assert list(shard_range(3, 0, 4)) == [0] and list(shard_range(3, 3, 4)) == []
    assert sum(len(shard_range(256, r, 8)) for r in range(8)) == 256
    assert max_over_ranks(3.5) == 3.5                       # no process group: identity
