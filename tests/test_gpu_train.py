"""Training path (BASELINE config 2) parity: the backward building blocks against torch CPU autograd
of the same reference modules, then one full decoder training step (content + style losses through
the frozen VGG taps) against the CPU oracle.  Gradients are carried in bf16 between layers, so the
bars are statistical: relative L2 <= 3e-2 and cosine >= 0.999 per gradient tensor for single ops,
relative L2 <= 1e-1 / cosine >= 0.99 for the 18-layer end-to-end step; losses within 2 %."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate as R
from tests.gpu_util import bf16r

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def cos(a, b):
    return F.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0).item()


@pytest.mark.parametrize("shape", [(2, 12, 20, 64, 128), (1, 16, 16, 128, 64), (3, 8, 24, 256, 256),
                                   (2, 16, 16, 64, 3)])
def test_wgrad_gemm(shape):
    """dW, db of a reflect-padded 3x3 conv from planar operands vs torch autograd."""
    from arbitrarystyletransfer_b200 import engine as E, train_ops as T
    N, H, W, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = bf16r(torch.randn(N, cin, H, W, generator=g))
    dz = bf16r(torch.randn(N, cout, H, W, generator=g))
    w = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    (F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b) * dz).sum().backward()
    xin = E.nchw_to_native(x.cuda(), reflect=True)
    cz = max(cout, 8) if cout % 8 else cout
    dzn = torch.zeros(N, H + 4, W + 4, 64 if cout == 3 else cout, device="cuda", dtype=torch.bfloat16)
    from arbitrarystyletransfer_b200 import _lib as L
    lib = L.load()
    dzd = dz.cuda().contiguous()
    L.check(lib.ast_nchw_to_native_ex(dzd.data_ptr(), dzn.data_ptr(), N, cout, H, W, dzn.shape[3], 2, L.stream_ptr()))
    dzT = T.to_planar(dzn, N, dzn.shape[3], H, W, 2, False)
    xT = T.to_planar(xin, N, cin, H, W, 1, True, nshift=3)
    # planar layout: q = (n*(H+2)+ph)*wp + pw, zero beyond pw = W+1, copy s holds x[q+s-1]
    wp = T._wp(W)
    refx = F.pad(F.pad(x, (1, 1, 1, 1), mode="reflect"), (0, wp - W - 2)).permute(1, 0, 2, 3).reshape(cin, -1)
    assert torch.equal(xT[1].float().cpu(), refx)
    assert torch.equal(xT[0].float().cpu()[:, 1:], refx[:, :-1]) and torch.equal(xT[2].float().cpu()[:, :-1], refx[:, 1:])
    gw, gb = T.conv_wgrad(dzT, xT, N, H, W, cin, cout, w.detach().cuda(), b.detach().cuda())
    assert rel(gw.cpu(), w.grad) < 2e-3 and rel(gb.cpu(), b.grad) < 2e-3


def test_decoder_backward_small():
    """ClassicDecoder autograd (all 9 convs, reflection pad, 3 upsamples) vs torch CPU autograd of the
    reference's commented nn.Sequential arithmetic (oracle.decoder_forward)."""
    from arbitrarystyletransfer_b200 import models as M
    dw, db = R.make_decoder_weights(1)
    g = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn(2, 512, 4, 6, generator=g) + 0.5)
    gimg = torch.randn(2, 3, 32, 48, generator=g)
    wr = [w.clone().requires_grad_(True) for w in dw]
    br = [b.clone().requires_grad_(True) for b in db]
    ref = R.decoder_forward(x, wr, br)
    (ref * gimg).sum().backward()
    dec = M.ClassicDecoder().cuda()
    with torch.no_grad():
        for c, w, b in zip(dec._convs(), dw, db):
            c.weight.copy_(w); c.bias.copy_(b)
    out = dec(x.cuda())
    assert out.requires_grad
    assert R.psnr(out.detach().cpu(), ref.detach()) >= 40.0
    (out * gimg.cuda()).sum().backward()
    for i, c in enumerate(dec._convs()):
        assert c.weight.grad.shape == wr[i].grad.shape
        r_w, c_w = rel(c.weight.grad.cpu(), wr[i].grad), cos(c.weight.grad.cpu(), wr[i].grad)
        r_b, c_b = rel(c.bias.grad.cpu(), br[i].grad), cos(c.bias.grad.cpu(), br[i].grad)
        assert c_w >= 0.995 and r_w <= 1e-1, f"conv {i} weight grad rel {r_w} cos {c_w}"
        assert c_b >= 0.995 and r_b <= 1e-1, f"conv {i} bias grad rel {r_b} cos {c_b}"


@pytest.mark.parametrize("taps", [['relu_1', 'relu_3', 'relu_5', 'relu_9'], ['conv_1', 'conv_3', 'conv_5']])
def test_encoder_input_gradient(taps):
    """PretrainedEncoder with an input that requires grad: taps and d(sum taps*g)/d(img) vs the oracle."""
    from arbitrarystyletransfer_b200 import models as M
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    g = torch.Generator().manual_seed(3)
    img = torch.rand(2, 3, 32, 48, generator=g)
    imr = img.clone().requires_grad_(True)
    ref = R.vgg_forward(imr, vw, vb, taps)
    gts = [torch.randn(t.shape, generator=g) * (1.0 / t.numel() ** 0.5) for t in ref]
    sum((t * gt).sum() for t, gt in zip(ref, gts)).backward()
    enc = M.PretrainedEncoder(taps).cuda()
    with torch.no_grad():
        for c, w, b in zip(enc._convs(), vw, vb):
            c.weight.copy_(w); c.bias.copy_(b)
    imd = img.cuda().requires_grad_(True)
    outs = enc(imd)
    assert len(outs) == len(ref)
    for o, t in zip(outs, ref):
        assert rel(o.detach().cpu(), t.detach()) < 5e-2
    sum((o * gt.cuda()).sum() for o, gt in zip(outs, gts)).backward()
    r, c = rel(imd.grad.cpu(), imr.grad), cos(imd.grad.cpu(), imr.grad)
    assert c >= 0.99 and r <= 1.5e-1, f"image grad rel {r} cos {c}"
    assert all(p.grad is None for p in enc.parameters())   # frozen loss network


def test_full_training_step_vs_oracle():
    """Config-2 shaped step at reduced size: batch 2 at 64x64, classic AdaIN objective
    (content loss at relu4_1 + style losses at relu1_1..relu4_1), Adam step on the decoder."""
    from arbitrarystyletransfer_b200 import models as M, losses as Ls
    taps = ['relu_1', 'relu_3', 'relu_5', 'relu_9']
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)
    c, s = R.rand_image(2, 64, 201), R.rand_image(2, 64, 202)
    # ---- oracle (CPU fp32 autograd over the restated reference functions)
    wr = [w.clone().requires_grad_(True) for w in dw]
    br = [b.clone().requires_grad_(True) for b in db]
    with torch.no_grad():
        fc = R.vgg_relu4_1(c, vw, vb)
        st = R.vgg_forward(s, vw, vb, taps)
        t = R.adain(fc, st[-1])
    gimg = R.decoder_forward(t, wr, br)
    gt = R.vgg_forward(gimg, vw, vb, taps)
    loss_c = R.compute_content_loss(gt[-1], t)
    loss_s = sum(R.compute_style_loss(a, b) for a, b in zip(gt, st))
    loss_ref = loss_c + loss_s
    loss_ref.backward()
    # ---- this package on the GPU
    enc = M.PretrainedEncoder(taps).cuda()
    dec = M.ClassicDecoder().cuda()
    with torch.no_grad():
        for cv, w, b in zip(enc._convs(), vw, vb):
            cv.weight.copy_(w); cv.bias.copy_(b)
        for cv, w, b in zip(dec._convs(), dw, db):
            cv.weight.copy_(w); cv.bias.copy_(b)
    opt = torch.optim.Adam(dec.parameters(), lr=2e-4, betas=(0.9, 0.999), eps=1e-5)   # train.py:61
    cd, sd = c.cuda(), s.cuda()
    with torch.no_grad():
        fcd = enc(cd)[-1]
        std_ = enc(sd)
        td = M.AdaIN()(fcd, std_[-1])
    opt.zero_grad()
    gd = dec(td)
    gtd = enc(gd)
    ld = Ls.compute_content_loss(gtd[-1], td) + sum(Ls.compute_style_loss(a, b) for a, b in zip(gtd, std_))
    ld.backward()
    assert ld.item() == pytest.approx(loss_ref.item(), rel=2e-2)
    worst = 1.0
    for i, cv in enumerate(dec._convs()):
        cw = cos(cv.weight.grad.cpu(), wr[i].grad)
        worst = min(worst, cw)
        assert cw >= 0.98, f"decoder conv {i}: weight-grad cosine {cw}, rel {rel(cv.weight.grad.cpu(), wr[i].grad)}"
    total = torch.cat([cv.weight.grad.flatten() for cv in dec._convs()]).cpu()
    total_ref = torch.cat([w.grad.flatten() for w in wr])
    assert cos(total, total_ref) >= 0.99 and rel(total, total_ref) <= 1.5e-1
    torch.nn.utils.clip_grad_norm_(dec.parameters(), 2.0, error_if_nonfinite=True)       # train.py:292
    before = dec._convs()[0].weight.detach().clone()
    opt.step()
    assert not torch.equal(before, dec._convs()[0].weight.detach())
    # the packed-weight cache must notice the optimiser step
    with torch.no_grad():
        out2 = dec(td)
    assert not torch.equal(out2, gd.detach())
