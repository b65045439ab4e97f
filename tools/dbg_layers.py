"""Per-role wait-cycle breakdown of the conv kernel on the bench layers (AST_CONV_DEBUG=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import _lib as L, engine as E
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
def layer(cin, cout, hw, epi, halo):
    x = torch.randn(N, hw + 2, hw + 2, cin, device=dev).bfloat16()
    w = torch.randn(9, cout, cin, device=dev).bfloat16()
    b = torch.zeros(cout, device=dev)
    ho = hw // 2 if epi == 1 else (2 * hw if epi == 2 else hw)
    y = torch.empty(N, ho + 2, ho + 2, cout, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        E.conv3x3(x, w, b, y, N=N, H=hw, W=hw, cin=cin, cout=cout, relu=True, epilogue=epi, halo=halo)
    torch.cuda.synchronize()
for cfg in [(64, 64, 512, 1, 0), (64, 64, 512, 0, 1), (64, 128, 256, 0, 0), (128, 128, 256, 1, 0), (128, 64, 256, 2, 1),
            (256, 128, 128, 2, 1), (256, 256, 128, 0, 0)]:
    layer(*cfg)
