"""The first VGG layer (Normalization + conv_1 3 -> 64 + ReLU, models.py:129-131, 198-224) alone at the bench shape,
as algorithmic GB/s (fp32 NCHW image read + bf16 64-channel native output written) against the measured HBM copy peak."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda")
N, S = args.n, args.size
torch.manual_seed(0)
img = torch.rand(N, 3, S, S, device=dev)
w = torch.randn(64, 3, 3, 3, device=dev) * 0.2
b = torch.randn(64, device=dev) * 0.1
out = E.native_empty(N, S, S, 64, dev, True)
alg = N * 3 * S * S * 4 + N * S * S * 64 * 2
pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.isfile(pk) else 6650.0
for _ in range(3):
    E.conv3x3_first(img, w, b, out)
torch.cuda.synchronize()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.reps):
    E.conv3x3_first(img, w, b, out)
c.record()
torch.cuda.synchronize()
ms = a.elapsed_time(c) / args.reps
print(f"conv3x3_first ({N},3,{S},{S}) -> 64 ch: {ms * 1e3:.1f} us  {alg / ms / 1e6:.0f} GB/s algorithmic = "
      f"{alg / ms / 1e6 / peak * 100:.1f} % of the HBM copy peak ({alg / 1e6:.0f} MB)")
