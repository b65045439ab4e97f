import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import models as M, losses as Ls
B, S = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
taps = ['relu_1', 'relu_3', 'relu_5', 'relu_9']
torch.manual_seed(0); enc = M.PretrainedEncoder(taps).to(dev); M.calibrate_encoder_bias(enc, size=64)
torch.manual_seed(1); dec = M.ClassicDecoder().to(dev)
opt = torch.optim.Adam(dec.parameters(), lr=2e-4, eps=1e-5, capturable=True)
ada = M.AdaIN()
c = torch.rand(B, 3, S, S, device=dev); s = torch.rand(B, 3, S, S, device=dev)
state = {}
def p_encode():
    with torch.no_grad():
        state['fc'] = enc(c)[-1]; state['st'] = enc(s); state['t'] = ada(state['fc'], state['st'][-1])
def p_dec():
    state['g'] = dec(state['t'])
def p_enc_grad():
    state['gt'] = enc(state['g'])
def p_loss():
    l = Ls.compute_content_loss(state['gt'][-1], state['t'])
    for a, b in zip(state['gt'], state['st']):
        l = l + Ls.compute_style_loss(a, b)
    state['loss'] = l
def p_bwd():
    opt.zero_grad(set_to_none=True); state['loss'].backward()
def p_clip():
    torch.nn.utils.clip_grad_norm_(dec.parameters(), 2.0)
def p_opt():
    opt.step()
pieces = [p_encode, p_dec, p_enc_grad, p_loss, p_bwd, p_clip, p_opt]
for _ in range(3):
    for f in pieces: f()
torch.cuda.synchronize()
# capture cumulative prefixes to find the first piece that breaks capture
for k in range(1, len(pieces) + 1):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for f in pieces[:k]: f()
        g.replay(); torch.cuda.synchronize()
        print("capture ok up to", pieces[k - 1].__name__, flush=True)
    except Exception as e:
        print("capture FAILED at", pieces[k - 1].__name__, repr(e)[:300], flush=True)
        traceback.print_exc(limit=6)
        break
