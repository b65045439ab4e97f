"""K1 (fused AdaIN) alone at the BASELINE shapes: GB/s against the measured HBM copy peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import functional as Fn
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (N, C, H, W, K, dt) in [(32, 512, 64, 64, 1, torch.float32), (1, 512, 256, 256, 4, torch.float32), (1, 512, 32, 32, 1, torch.float32),
                            (32, 512, 64, 64, 1, torch.bfloat16), (8, 512, 128, 128, 1, torch.float32)]:
    c = torch.relu(torch.randn(N, C, H, W, device=dev) * 3 + 1).to(dt)
    ss = [(torch.randn(N, C, H, W, device=dev) * 2 + 3).to(dt) for _ in range(K)]
    w = [1.0 / K] * K
    Fn.adain_forward(c, ss, w); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); Fn.adain_forward(c, ss, w); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = (2 + K) * c.numel() * c.element_size()
    print(f"({N},{C},{H},{W}) K={K} {str(dt).split('.')[-1]:8s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  {nbytes/ms/1e6/6452.5*100:5.1f} % of measured HBM copy peak", flush=True)
