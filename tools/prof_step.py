"""Fixed workload for ncu: ONE stylise pass at the bench shape (batch 32, 512x512) = 18 encoder launches, the native
AdaIN (3 launches) and 9 decoder launches; everything before cudaProfilerStart is warm-up."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = bench.build_engine(dev)
if len(sys.argv) > 2:
    eng.fold = sys.argv[2] != "nofold"
ci = torch.rand(N, 3, 512, 512, device=dev)
si = torch.rand(N, 3, 512, 512, device=dev)
for _ in range(2):
    img = eng.stylize(ci, si)
torch.cuda.synchronize()
torch.cuda.profiler.start()
img = eng.stylize(ci, si)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(img.mean()))
