"""K1 on the long-row path only (BASELINE config 5: (1,512,256,256), K = 4 styles) for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import functional as Fn
dev = torch.device("cuda")
c = torch.relu(torch.randn(1, 512, 256, 256, device=dev) * 3 + 1)
ss = [torch.randn(1, 512, 256, 256, device=dev) * 2 + 3 for _ in range(4)]
out = torch.empty_like(c)
for _ in range(3):
    Fn.adain_forward(c, ss, weights=[0.4, 0.3, 0.2, 0.1], alpha=0.6, out=out)
torch.cuda.synchronize()
print("ok", float(out.mean()))
