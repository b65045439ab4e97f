"""One eager config-2 training step under the profiler range (for the ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import models as M, losses as Ls
dev = torch.device("cuda")
taps = ['relu_1', 'relu_3', 'relu_5', 'relu_9']
torch.manual_seed(0); enc = M.PretrainedEncoder(taps).to(dev); M.calibrate_encoder_bias(enc)
torch.manual_seed(1); dec = M.ClassicDecoder().to(dev)
opt = torch.optim.Adam(dec.parameters(), lr=2e-4, eps=1e-5, capturable=True)
ada = M.AdaIN()
c = torch.rand(8, 3, 256, 256, device=dev); s = torch.rand(8, 3, 256, 256, device=dev)
def step():
    with torch.no_grad():
        fc = enc(c)[-1]; st = enc(s); t = ada(fc, st[-1])
    opt.zero_grad(set_to_none=True)
    gt = enc(dec(t))
    loss = Ls.compute_content_loss(gt[-1], t)
    for a, b in zip(gt, st):
        loss = loss + Ls.compute_style_loss(a, b)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(dec.parameters(), 2.0)
    opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
