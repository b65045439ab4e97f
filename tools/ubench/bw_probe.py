"""Pure-write vs pure-read vs copy bandwidth of the box (torch fill_ / sum / copy_ on 1 GiB), to put write-dominated
kernels (conv1_1: 91 % of its traffic is the 64-channel output) on the right roof."""
import torch
dev = torch.device("cuda")
n = 1 << 28                      # 2^28 fp32 = 1 GiB
x = torch.empty(n, device=dev)
y = torch.empty(n, device=dev)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


gb = n * 4 / 1e9
ms = t(lambda: x.fill_(1.0)); print(f"write-only (fill_ 1 GiB):   {ms*1e3:7.1f} us  {gb/ms*1e3:7.0f} GB/s")
ms = t(lambda: x.sum());      print(f"read-only  (sum 1 GiB):     {ms*1e3:7.1f} us  {gb/ms*1e3:7.0f} GB/s")
ms = t(lambda: y.copy_(x));   print(f"copy (1 GiB -> 1 GiB):      {ms*1e3:7.1f} us  {2*gb/ms*1e3:7.0f} GB/s (read + write)")
