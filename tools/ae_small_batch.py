"""Config-3 AutoEncoder training step on ONE GPU at the per-GPU batch sizes of 2 / 4 / 8-way sharding (16, 8, 4): where the
step's floor is when the kernels no longer fill the GPU (what SCALE's train_ae at N GPUs can reach at best)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for gb in (32, 16, 8, 4):
    r = bench.time_train_ae(dev, 0, 1, steps=6, warmup=3, global_batch=gb, size=256, cpu=False)
    print(f"batch {gb:2d}: {r['ms_per_step']:7.2f} ms/step graph ({r['value']:6.1f} steps/s), eager {1e3 / r['eager_steps_per_s']:7.2f} ms; mode: {r['mode'][:40]}", flush=True)
