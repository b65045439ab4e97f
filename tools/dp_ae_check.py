"""2+ GPU data-parallel check of the AutoEncoder training step (BASELINE config 3): every rank takes its
contiguous shard of the global batch (per-shard BatchNorm statistics = DDP semantics), gradients go through ONE
NCCL all-reduce of the flat bucket (parallel.GradBucket), and rank 0 compares the result with the average of
the per-shard gradients computed one after the other on a single GPU.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dp_ae_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from arbitrarystyletransfer_b200 import models as M, mobilenet as MB, losses as Ls, parallel as P

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S = 4 * world, 64


def build():
    torch.manual_seed(0); enc = M.PretrainedEncoder().to(dev).eval(); M.calibrate_encoder_bias(enc, n_convs=16, size=64)
    for p in enc.parameters():
        p.requires_grad_(False)
    torch.manual_seed(2); ae = MB.AutoEncoder().to(dev).train()
    return enc, ae


def loss_of(enc, ae, x):
    recon = ae(x)
    loss = 100.0 * Ls.compute_content_loss(recon, x)
    with torch.no_grad():
        cm = enc(x)
    for a, b in zip(enc(recon), cm):
        loss = loss + 0.01 * Ls.compute_content_loss(a, b)
    return loss


x = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(301)).to(dev)
enc, ae = build()
P.broadcast_parameters(list(ae.parameters()))
bucket = P.GradBucket(ae.parameters())
bucket.zero()
loss_of(enc, ae, P.shard_batch(x, rank, world)).backward()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); bucket.all_reduce_mean(); ev1.record(); torch.cuda.synchronize()
if rank == 0:
    ref = torch.zeros_like(bucket.flat)
    for r in range(world):
        enc1, ae1 = build()
        loss_of(enc1, ae1, P.shard_batch(x, r, world)).backward()
        ref += torch.cat([p.grad.reshape(-1) for p in reversed(list(ae1.parameters()))]) / world
    rel = ((bucket.flat - ref).norm() / ref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(bucket.flat, ref, dim=0).item()
    print(f"AE DP world={world}: bucket {bucket.numel} floats ({bucket.numel * 4 / 1e6:.1f} MB), all-reduce "
          f"{ev0.elapsed_time(ev1) * 1e3:.0f} us, all-reduced vs mean of per-shard gradients rel {rel:.3e} cos {cos:.6f}",
          flush=True)
    assert bucket.numel == 2925931 and cos > 0.9999 and rel < 1e-3, (rel, cos)
dist.barrier(); dist.destroy_process_group()
