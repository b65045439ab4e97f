"""Stage-by-stage comparison of the GPU AutoEncoder (train mode) with the CPU oracle on the non-degenerate state:
where does the relative error grow?  Also reports what bf16 storage alone does to the oracle (each block's output
rounded to bf16) so that accumulation of rounding can be told from a defect."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from arbitrarystyletransfer_b200 import mobilenet as MB
from oracle import restate as R, restate_ae as A

def rel(a, b): return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()
def nchw(t): return t.float().permute(0, 3, 1, 2).contiguous().cpu()
bf = lambda t: t.to(torch.bfloat16).float()

torch.manual_seed(2)
ae = MB.AutoEncoder().cuda().train()
sd = A.activate_gates(A.make_ae_state(2))
ae.load_state_dict(sd, strict=True)
x = R.rand_image(2, 32, 301)
with torch.no_grad():
    # ---- oracle, fp32 and with bf16-rounded block outputs ----
    def oracle(rounder):
        P = A.clone_state(sd)
        outs = {}
        h = F.hardswish(F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), P["encoder.mob_net.0.0.weight"]))
        h = rounder(h); outs["stem"] = h
        taps = {}
        for i, (inp, oup, s, t, k) in enumerate(A.encoder_block_specs(), start=1):
            h = rounder(A.depthwise_block(P, f"encoder.mob_net.{i}", h, inp, oup, s, t, k, norm=True, training=True))
            outs[f"enc{i}"] = h
            if i in (12, 14): taps[i] = h
        z = rounder(A.depthwise_block(P, "ada_out", torch.cat((taps[12], taps[14]), 1), 256, 128, 1, 3, 3, norm=False, use_identity=False))
        outs["code"] = z
        h = z
        for i, (inp, oup, s, t, k, up) in enumerate(A.decoder_block_specs()):
            b = f"decoder._decoder_blocks.{i}"
            h = rounder(A.depthwise_block(P, b + "._conv", h, inp, oup, s, t, k, norm=False))
            if up:
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = rounder(A.depthwise_block(P, b + "._upsample_2", h, oup, oup, 1, 1, 3, norm=False))
            outs[f"dec{i}"] = h
        return outs
    o32 = oracle(lambda t: t)
    o16 = oracle(bf)
    # ---- GPU ----
    g = {}
    y = MB._StemFn.apply(x.cuda(), ae.encoder.mob_net[0][0].weight) if False else None
    lib_x = x.cuda().float().contiguous()
    yy = ae.encoder.forward_nhwc(lib_x, (0,), False)[0]; g["stem"] = nchw(yy)
    h = yy
    taps = {}
    for i, layer in enumerate(ae.encoder.mob_net):
        if i == 0: continue
        h = layer.forward_nhwc(h); g[f"enc{i}"] = nchw(h)
        if i in (12, 14): taps[i] = h
    z = ae.ada_out.forward_nhwc(torch.cat((taps[12], taps[14]), dim=3)); g["code"] = nchw(z)
    h = z
    for i, blk in enumerate(ae.decoder._decoder_blocks):
        h = blk.forward_nhwc(h); g[f"dec{i}"] = nchw(h)
print(f"{'stage':8s} {'gpu vs fp32':>12s} {'bf16-oracle vs fp32':>20s} {'gpu vs bf16-oracle':>20s}")
for k in o32:
    print(f"{k:8s} {rel(g[k], o32[k]):12.3e} {rel(o16[k], o32[k]):20.3e} {rel(g[k], o16[k]):20.3e}")
