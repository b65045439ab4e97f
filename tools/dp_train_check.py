"""2+ GPU data-parallel check of the decoder training step: every rank takes its contiguous shard of
the global batch, gradients go through ONE NCCL all-reduce of the flat bucket (parallel.GradBucket),
and rank 0 compares the result with the same step on the whole batch on one GPU.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_train_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from arbitrarystyletransfer_b200 import models as M, losses as Ls, parallel as P

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
taps = ['relu_1', 'relu_3', 'relu_5', 'relu_9']
B, S = 8, 128

def build():
    torch.manual_seed(0); enc = M.PretrainedEncoder(taps).to(dev); M.calibrate_encoder_bias(enc, size=64)
    torch.manual_seed(1); dec = M.ClassicDecoder().to(dev)
    return enc, dec

def loss_of(enc, dec, c, s):
    ada = M.AdaIN()
    with torch.no_grad():
        fc = enc(c)[-1]; st = enc(s); t = ada(fc, st[-1])
    gt = enc(dec(t))
    loss = Ls.compute_content_loss(gt[-1], t)
    for a, b in zip(gt, st):
        loss = loss + Ls.compute_style_loss(a, b)
    return loss

g = torch.Generator().manual_seed(201)
c = torch.rand(B, 3, S, S, generator=g).to(dev); s = torch.rand(B, 3, S, S, generator=g).to(dev)
enc, dec = build()
P.broadcast_parameters(list(dec.parameters()))
bucket = P.GradBucket(dec.parameters())
bucket.zero()
cs, ss = P.shard_batch(c, rank, world), P.shard_batch(s, rank, world)
loss = loss_of(enc, dec, cs, ss)
loss.backward()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); bucket.all_reduce_mean(); ev1.record(); torch.cuda.synchronize()
if rank == 0:
    enc1, dec1 = build()
    l1 = loss_of(enc1, dec1, c, s)
    l1.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in reversed(list(dec1.parameters()))])
    rel = ((bucket.flat - ref).norm() / ref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(bucket.flat, ref, dim=0).item()
    print(f"DP world={world}: bucket {bucket.numel} floats ({bucket.numel * 4 / 1e6:.1f} MB), all-reduce {ev0.elapsed_time(ev1) * 1e3:.0f} us, "
          f"sharded-vs-full-batch gradient rel {rel:.3e} cos {cos:.6f}", flush=True)
    # the style-loss statistics (mean/std/Gram per sample) are per-sample, so DP == full batch up to
    # fp32/bf16 reduction order; Huber 'mean' over a shard averages to the global mean when shards are equal
    assert cos > 0.999, (rel, cos)
dist.barrier(); dist.destroy_process_group()
