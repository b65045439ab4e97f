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
            out_q.put((bucket.flat.clone(), outs))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_bucket_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, outs = q.get()
    for p in procs:
        p.join(60)
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
