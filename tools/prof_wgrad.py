"""Fixed workload for ncu: K2wn (csrc/wgrad_mn.cu) once per classic-decoder layer at the config-2 shapes (batch 8,
256x256 image); everything before cudaProfilerStart is warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import engine as E, train_ops as T

dev = torch.device("cuda")
N, size = 8, 256
h = size // 8
cases = []
for i in range(9):
    cin, cout, relu, up = E.DECODER_SPEC[i] if i < 8 else (64, 3, False, False)
    cz = 64 if cout == 3 else cout
    x = (torch.randn(N, h + 2, h + 2, cin, device=dev) * 0.5).to(torch.bfloat16)
    dz = torch.zeros(N, h + 4, h + 4, cz, device=dev, dtype=torch.bfloat16)
    dz[:, 2:-2, 2:-2, :cout] = (torch.randn(N, h, h, cout, device=dev) * 0.5).to(torch.bfloat16)
    cases.append((dz, x, h, cin, cout, torch.empty(cout, cin, 3, 3, device=dev), torch.empty(cout, device=dev)))
    if up:
        h *= 2


def run():
    for dz, x, hh, cin, cout, w, b in cases:
        T.conv_wgrad_native(dz, 2, x, N, hh, hh, cin, cout, w, b)


run(); run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
