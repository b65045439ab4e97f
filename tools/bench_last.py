"""The last decoder convolution (64 -> 3, reflect pad; models.py:626-627) alone at the bench shape: the taps-in-N kernel
(conv_last_tn.cu, default) against the taps-in-K implicit GEMM (tap-box implementation, and the kw-box one when run
with AST_LAST_TAPS_IN_K=1), as algorithmic GB/s against the measured HBM copy peak."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import _lib as L, engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda")
N, S = args.n, args.size
torch.manual_seed(0)
x = torch.randn(N, 64, S, S, device=dev) * 0.5
xn = E.nchw_to_native(x, reflect=True)
del x
w = torch.randn(3, 64, 3, 3, device=dev) * 0.05
b = torch.randn(3, device=dev)
wpk = E.pack_conv_weight(w, cout_pad=16)
out = torch.empty(N, 3, S, S, device=dev)
ref = torch.empty_like(out)
alg = N * (S + 2) * (S + 2) * 64 * 2 + N * 3 * S * S * 4
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.isfile(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def run(impl, o):
    for _ in range(3):
        E.conv3x3_last(xn, w, wpk, b, o, False, impl=impl)
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        E.conv3x3_last(xn, w, wpk, b, o, False, impl=impl)
    c.record()
    torch.cuda.synchronize()
    return a.elapsed_time(c) / args.reps


for name, impl, o in (("default (AUTO)", L.CONV_AUTO, out), ("tap-box implicit GEMM, taps in K", L.CONV_TC_TAPBOX, ref)):
    ms = run(impl, o)
    print(f"{name:36s} {ms * 1e3:8.1f} us  {alg / ms / 1e6:7.0f} GB/s algorithmic = {alg / ms / 1e6 / peak * 100:5.1f} % of the HBM copy peak "
          f"({alg / 1e6:.0f} MB)")
print("max |diff| between the two:", float((out - ref).abs().max()), " taps-in-K env:", os.environ.get("AST_LAST_TAPS_IN_K"))
