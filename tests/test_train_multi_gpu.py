"""Data-parallel training over NCCL: two ranks (one process per GPU, 127.0.0.1 rendezvous), each with its own half of the
batch and local batch statistics; ONE all-reduce of the flat gradient bucket.  Checks that (a) the reduced gradient equals the
sum of the two local gradients computed by a single process, bit for bit (NCCL sums two fp32 buffers: one add per element),
(b) both ranks hold identical parameters after Adam, (c) the step matches a single-process emulation of the same update.
Skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B_LOCAL, N, STEPS = 4, 512, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(rank):
    from scanobjectnn_b200.synthetic import make_clouds
    xyz = make_clouds("ball", B_LOCAL, N, seed=400 + rank)
    labels = ((np.arange(B_LOCAL) * 3 + rank) % 15).astype(np.int32)
    return xyz, labels


def _trainer(dev, pg=None):
    from scanobjectnn_b200 import pointnet2_cls_ssg
    from scanobjectnn_b200.training import PointNet2ClsTrainer
    params = pointnet2_cls_ssg.init_params(seed=11, device=dev)
    return PointNet2ClsTrainer(params, B_LOCAL, N, device=dev, process_group=pg), params


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    tr, _ = _trainer(dev)
    assert tr.world == world
    xyz, labels = _data(rank)
    x, y = torch.from_numpy(xyz).to(dev), torch.from_numpy(labels).to(dev)
    out = []
    for _ in range(STEPS):
        logits = tr.forward(x, 0.5)
        loss, dl = tr.loss_and_grad(logits, y)
        tr.backward(dl)
        local = tr.fp.grad.clone()
        tr.allreduce_grads()
        reduced = tr.fp.grad.clone()
        tr.adam(1e-3)
        out.append((local.cpu().numpy(), reduced.cpu().numpy(), tr.fp.flat.detach().cpu().numpy().copy(), float(loss.item())))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradient_allreduce_and_identical_parameters():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for s in range(STEPS):
        l0, r0, p0, loss0 = res[0][s]
        l1, r1, p1, loss1 = res[1][s]
        assert np.array_equal(r0, r1), "ranks disagree on the reduced gradient"
        assert np.array_equal(r0, l0 + l1), "all-reduce is not the plain sum of the two local buckets"
        assert np.array_equal(p0, p1), "parameters diverged between the ranks"
        assert np.isfinite(loss0) and np.isfinite(loss1)
    # single-process emulation of step 1: the same two local gradients, summed, averaged inside Adam (scale 1/world)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    grads = []
    for rank in range(world):
        tr, _ = _trainer(dev)
        xyz, labels = _data(rank)
        logits = tr.forward(torch.from_numpy(xyz).to(dev), 0.5)
        _, dl = tr.loss_and_grad(logits, torch.from_numpy(labels).to(dev))
        tr.backward(dl)
        grads.append(tr.fp.grad.clone())
        assert np.array_equal(grads[-1].cpu().numpy(), res[rank][0][0]), "the local gradient is not reproducible across processes"
    tr, _ = _trainer(dev)
    tr.world = world
    tr.fp.grad.copy_(grads[0] + grads[1])
    # running statistics are local to a rank (the reference's single-GPU trainer has no cross-replica batch norm); compare the
    # trainable parameters only: everything Adam touched
    before = tr.fp.flat.detach().clone()
    tr.adam(1e-3)
    after = tr.fp.flat.detach().cpu().numpy()
    moved = (before.cpu().numpy() != after)
    assert moved.any()
    assert np.array_equal(after[moved], res[0][0][2][moved])
