"""Micro-benchmark of the fused conv1_1 + conv1_2 + pool kernel (ast_conv12_fused) at the bench shape, next to the two
separate launches.  AST_CONV_DBGFLAGS selects bottleneck-elimination variants (see conv12_fused.cuh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import engine as E, _lib as L

N, S = int(os.environ.get("N", 32)), int(os.environ.get("S", 512))
dev = torch.device("cuda")
g = torch.Generator().manual_seed(1)
w1 = (torch.randn(64, 3, 3, 3, generator=g) * 0.4).to(dev)
b1 = (torch.randn(64, generator=g) * 0.2).to(dev)
w2 = (torch.randn(64, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5).to(dev)
b2 = (torch.randn(64, generator=g) * 0.1).to(dev)
img = torch.rand(N, 3, S, S, generator=g).to(dev)
wpk2 = E.pack_conv_weight(w2)
out = E.native_empty(N, S // 2, S // 2, 64, dev, True)
x = E.native_empty(N, S, S, 64, dev, True)


def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]

def sustained(fn, seconds=1.0):
    """back-to-back launches for ~`seconds`: the clock the power cap allows, as inside the bench loop"""
    ms1 = timed(fn, 5)
    reps = max(10, int(seconds * 1e3 / ms1))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps // 2): fn()
    a.record()
    for _ in range(reps): fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

flags = os.environ.get("AST_CONV_DBGFLAGS", "0")
ms = timed(lambda: E.conv12_fused(img, w1, b1, wpk2, b2, out))
fl = 2.0 * 64 * (27 + 576) * S * S * N
line = f"flags {flags:>4} V={os.environ.get('AST_CONV12_V', '3')}: fused {ms * 1e3:7.1f} us ({fl / ms / 1e9:6.0f} TFLOP/s)"
if os.environ.get("SUSTAINED"):
    line += f", sustained {sustained(lambda: E.conv12_fused(img, w1, b1, wpk2, b2, out)) * 1e3:7.1f} us"
if flags == "0":
    m1 = timed(lambda: E.conv3x3_first(img, w1, b1, x))
    m2 = timed(lambda: E.conv3x3(x, wpk2, b2, out, N=N, H=S, W=S, cin=64, cout=64, relu=True, epilogue=L.EPI_POOL2, halo=L.HALO_KEEP))
    line += f"; separate: conv1_1 {m1 * 1e3:.1f} + conv1_2 {m2 * 1e3:.1f} = {(m1 + m2) * 1e3:.1f} us"
    if os.environ.get("SUSTAINED"):
        def both():
            E.conv3x3_first(img, w1, b1, x)
            E.conv3x3(x, wpk2, b2, out, N=N, H=S, W=S, cin=64, cout=64, relu=True, epilogue=L.EPI_POOL2, halo=L.HALO_KEEP)
        line += f", sustained {sustained(both) * 1e3:.1f} us"
print(line)
