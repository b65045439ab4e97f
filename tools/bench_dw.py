"""Micro-benchmark of the depthwise kernels (forward / data gradient / weight gradient) at decoder shapes.
Prints per-kernel time, algorithmic GB/s (one read + one write of the tensor; wgrad: two reads) and TFMA/s."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import _lib as L, mobilenet as MB

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", default="")
args = ap.parse_args()
lib = L.load()
dev = torch.device("cuda")
st = L.stream_ptr(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for C, k, S in [(240, 5, 256), (160, 5, 256), (144, 3, 256), (96, 3, 256), (320, 3, 128), (384, 3, 32)]:
    if args.only and args.only != f"{C}x{k}":
        continue
    N = args.n
    x = torch.randn(N, S, S, C, device=dev).to(torch.bfloat16)
    dy = torch.randn(N, S, S, C, device=dev).to(torch.bfloat16)
    w = torch.randn(k * k, C, device=dev)
    out = torch.empty_like(x)
    pool = torch.empty(N, C, device=dev)
    dwg = torch.zeros(C, 1, k, k, device=dev)
    nbytes = x.numel() * 2
    fma = x.numel() * k * k
    f = lambda: L.check(lib.ast_dw_conv(x.data_ptr(), w.data_ptr(), None, out.data_ptr(), pool.data_ptr(), N, C, S, S, k, 1, 0, 2, st))
    d = lambda: L.check(lib.ast_dw_conv_dgrad(dy.data_ptr(), w.data_ptr(), x.data_ptr(), None, None, out.data_ptr(), N, C, S, S, k, 1, 0, st))
    g = lambda: L.check(lib.ast_dw_conv_wgrad(dy.data_ptr(), x.data_ptr(), dwg.data_ptr(), N, C, S, S, k, 1, 0, st))
    for name, fn, nb in (("fwd", f, 2 * nbytes), ("dgrad", d, 3 * nbytes), ("wgrad", g, 2 * nbytes)):
        ms = timed(fn, args.reps)
        print(f"C={C:4d} k={k} {S}x{S} N={N} {name:6s} {ms*1e3:9.1f} us  {nb/ms/1e6:8.1f} GB/s  {fma/ms/1e9:7.2f} TFMA/s", flush=True)
