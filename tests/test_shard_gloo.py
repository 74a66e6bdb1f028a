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


def _bucket_worker(rank, world, port, q):
    """Data-parallel training's host logic on CPU tensors: every trainable variable is a view of ONE flat buffer with the same layout
    on every rank, and ONE all_reduce of the flat gradient bucket sums every variable's gradient (training.py: FlatParams,
    PointNet2ClsTrainer.allreduce_grads)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scanobjectnn_b200 import pointnet2_cls_ssg
    from scanobjectnn_b200.training import FlatParams

    params = pointnet2_cls_ssg.init_params(seed=1, device="cpu")
    fp = FlatParams(params)
    layout = [(k, (fp.views[k].data_ptr() - fp.flat.data_ptr()) // 4, fp.views[k].numel()) for k in fp.names]
    # the store aliases the flat buffer: an optimiser step on `flat` is visible through the TF variable names
    fp.flat.add_(1.0)
    aliased = all(bool((params[k] == fp.views[k]).all()) and params[k].data_ptr() == fp.views[k].data_ptr() for k in fp.names)
    # rank-dependent local gradients, written through the per-variable views like the backward kernels do
    for i, k in enumerate(fp.names):
        fp.grad_of(k).fill_(float((rank + 1) * (i + 1)))
    local = fp.grad.clone()
    dist.all_reduce(fp.grad)                                   # the ONE collective of a training step
    expect = local * (sum(r + 1 for r in range(world)) / (rank + 1))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, fp.total, len(fp.names), layout, aliased, bool(torch.equal(fp.grad, expect)), float(fp.grad.sum())))


def test_world_size_2_flat_gradient_bucket_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, total0, n0, layout0, al0, ok0, s0), (r1, total1, n1, layout1, al1, ok1, s1) = res
    # 11 conv / FC layers with batch norm x (weights, biases, gamma, beta) + fc3 (weights, biases) = 46 variables (the 11 biases under
    # batch norm and fc3's bias included; 34 of them carry a non-zero gradient), ~1.47 M floats = 5.9 MB
    assert n0 == n1 == 46 and total0 == total1 and total0 >= 1_460_000
    assert layout0 == layout1                                                   # same bucket layout on every rank
    assert all(off % 64 == 0 for _, off, _ in layout0)                          # 256-byte aligned segments
    assert al0 and al1 and ok0 and ok1 and s0 == s1
