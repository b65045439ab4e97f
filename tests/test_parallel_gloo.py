"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (no GPU, no CUDA kernels).
Checks the shard arithmetic, that the flat gradient bucket all-reduce reproduces the single-process
gradient of the global batch (mean losses => average of shard gradients), and shard gathering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arbitrarystyletransfer_b200 import parallel as P


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 32, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [P.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Conv2d(4, 6, 3, padding=1), torch.nn.ReLU(),
                               torch.nn.Conv2d(6, 3, 3, padding=1))


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        model = _model()
        if rank == 1:   # a diverged replica: broadcast must repair it
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        P.broadcast_parameters(list(model.parameters()))
        bucket = P.GradBucket(model.parameters())
        g = torch.Generator().manual_seed(11)
        x = torch.randn(8, 4, 10, 10, generator=g)
        y = torch.randn(8, 3, 10, 10, generator=g)
        xs, ys = P.shard_batch(x, rank, world), P.shard_batch(y, rank, world)
        bucket.zero()
        loss = torch.nn.functional.huber_loss(model(xs), ys)   # 'mean' loss on the shard
        loss.backward()
        bucket.all_reduce_mean()
        outs = P.gather_shards(model(xs).detach(), 8)
        if rank == 0:
            out_q.put((bucket.flat.numpy().copy(), outs.numpy().copy()))   # by value: the producer may exit first
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_bucket_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, outs = (torch.from_numpy(a) for a in q.get())
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    # single-process reference on the global batch
    model = _model()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 4, 10, 10, generator=g)
    y = torch.randn(8, 3, 10, 10, generator=g)
    out = model(x)
    torch.nn.functional.huber_loss(out, y).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in reversed(list(model.parameters()))])
    torch.testing.assert_close(flat, ref, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(outs, out.detach(), rtol=1e-6, atol=1e-6)


def _worker_ref_step(rank, world, port, out_q):
    """The reference's own step glue (train.py:287-300): ``optim.zero_grad()`` with torch's default
    set_to_none=True drops the bucket views; 7 samples over 2 ranks are unequal shards (4 + 3); a BatchNorm buffer
    diverges on rank 1 before the broadcast."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        torch.manual_seed(3)
        model = torch.nn.Sequential(torch.nn.Conv2d(4, 6, 3, padding=1), torch.nn.BatchNorm2d(6), torch.nn.ReLU(),
                                    torch.nn.Conv2d(6, 3, 3, padding=1)).eval()
        if rank == 1:
            with torch.no_grad():
                model[1].running_mean.add_(5.0)
                model[0].weight.add_(1.0)
        P.broadcast_module(model)
        bucket = P.GradBucket(model.parameters())
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(7, 4, 10, 10, generator=g)
        y = torch.randn(7, 3, 10, 10, generator=g)
        xs, ys = P.shard_batch(x, rank, world), P.shard_batch(y, rank, world)
        opt.zero_grad()                                   # set_to_none=True: the views are gone
        assert all(p.grad is None for p in model.parameters())
        loss = torch.nn.functional.huber_loss(model(xs), ys)
        loss.backward()
        assert not bucket.aliased()
        bucket.all_reduce_mean(local_count=xs.shape[0])   # adopts the fresh .grad tensors, weights by shard size
        assert bucket.aliased()
        if rank == 0:
            out_q.put((bucket.flat.numpy().copy(), model[1].running_mean.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_reference_zero_grad_and_unequal_shards():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_ref_step, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, rm = (torch.from_numpy(a) for a in q.get())
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Conv2d(4, 6, 3, padding=1), torch.nn.BatchNorm2d(6), torch.nn.ReLU(),
                                torch.nn.Conv2d(6, 3, 3, padding=1)).eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(7, 4, 10, 10, generator=g)
    y = torch.randn(7, 3, 10, 10, generator=g)
    torch.nn.functional.huber_loss(model(x), y).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in reversed(list(model.parameters()))])
    torch.testing.assert_close(flat, ref, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(rm, model[1].running_mean)   # rank 0's buffers won the broadcast


def test_bucket_adopts_foreign_gradients_single_process():
    model = _model()
    bucket = P.GradBucket(model.parameters())
    assert bucket.aliased()
    for p in model.parameters():
        p.grad = None
    assert not bucket.aliased()
    x = torch.randn(2, 4, 5, 5)
    model(x).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for p in reversed(list(model.parameters()))])
    assert bucket.adopt() == len(list(model.parameters()))
    assert bucket.aliased() and torch.equal(bucket.flat, want)
    bucket.zero()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in model.parameters())
