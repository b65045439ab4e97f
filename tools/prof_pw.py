"""One pointwise-conv shape through ast_pw_conv of the in-tree library or of an alternate build (A/B under ncu)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import mobilenet as MB, _lib as L

alt = os.environ.get("PW_ALT")
lib = L.load()
if alt:
    a = C.CDLL(os.path.abspath(alt))
    fn = a.ast_pw_conv
    fn.restype, fn.argtypes = lib.ast_pw_conv.restype, lib.ast_pw_conv.argtypes
    lib.ast_pw_conv = fn
cin, cout, raw, f16 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
px_side = int(sys.argv[5]) if len(sys.argv) > 5 else 256
dev = torch.device("cuda")
dt = torch.float16 if f16 else torch.bfloat16
x = torch.randn(32, px_side, px_side, cin, device=dev).to(dt)
w = (torch.randn(cout, cin, device=dev) * 0.1).to(dt)
b = torch.randn(cout, device=dev)
f = lambda: MB.pw_conv(x, w, b, 1 if raw else 0, cout, want_raw=bool(raw), f16=bool(f16))
for _ in range(3): f()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    f(); ev[i + 1].record()
torch.cuda.synchronize()
print(f"{'alt ' + alt if alt else 'in-tree'}: {cin}->{cout} raw={raw} f16={f16}: {sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5] * 1e3:.1f} us")
