"""MobileNet-style Encoder / Decoder / AutoEncoder (BASELINE config 3):
every K4 kernel in isolation against the same op in torch fp32 on bf16-rounded operands, then whole
blocks and the whole autoencoder (eval forward, train forward, one train_autoencoder.py step's
gradients) against the CPU oracle and the golden vectors made from the genuine reference.

Storage formats (include/ast_b200.h, K4): forward ACTIVATIONS are fp16 (11-bit significand), GRADIENTS bf16.  Tensors
handed to the kernels are built with ``nhwc`` (activation: fp16 bit patterns in the package's 16-bit container) or
``gnhwc`` (gradient: bf16) and read back with ``nchw`` / ``gnchw``; the torch references run in fp32 on operands
rounded the same way (``f16r`` / ``bf16r``).

Tolerances.  A single kernel agrees with fp32 arithmetic on the same rounded inputs to one rounding of its output:
<= 1e-3 relative L2 for fp16 outputs, <= 5e-3 for bf16 outputs; whole blocks and the whole autoencoder are held to the
bars stated with each test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate as R
from oracle import restate_ae as A
from tests.conftest import load_golden
from tests.gpu_util import bf16r

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()


def cos(a, b):
    return F.cosine_similarity(a.double().cpu().flatten(), b.double().cpu().flatten(), dim=0).item()


def f16r(x):
    """round-trip through fp16 (what the K4 path stores for forward activations and GEMM weights)."""
    return x.to(torch.float16).float()


def nhwc(x):
    """(N,C,H,W) fp32 -> (N,H,W,C) ACTIVATION tensor on the GPU: fp16 bit patterns in the 16-bit container dtype."""
    return x.permute(0, 2, 3, 1).contiguous().to(torch.float16).view(torch.bfloat16).cuda()


def nchw(t):
    """ACTIVATION tensor (fp16 bits) -> (N,C,H,W) fp32 on the CPU."""
    return t.contiguous().view(torch.float16).float().permute(0, 3, 1, 2).contiguous().cpu()


def gnhwc(x):
    """(N,C,H,W) fp32 -> (N,H,W,C) GRADIENT tensor (bf16) on the GPU."""
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def gnchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def wbits(w, f16):
    """2-D GEMM weights for pw_conv: fp16 bit patterns (forward) or bf16 (backward)."""
    return (w.to(torch.float16).view(torch.bfloat16) if f16 else w.to(torch.bfloat16)).cuda().contiguous()


def G(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------
# pointwise conv (tcgen05 GEMM), forward and the weight-gradient GEMM with MN-major operands
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 16, 16, 16, 96), (1, 12, 20, 96, 16), (2, 9, 7, 24, 144),
                                   (1, 16, 16, 240, 40), (2, 8, 8, 256, 768), (2, 8, 8, 768, 128),
                                   (3, 5, 5, 80, 320)])
@pytest.mark.parametrize("f16", [True, False])
def test_pw_conv_forward(shape, f16):
    """The pointwise GEMM in both of its roles: fp16 (forward activations) and bf16 (data gradients, attention)."""
    from arbitrarystyletransfer_b200 import mobilenet as MB
    N, H, W, cin, cout = shape
    g = G(sum(shape))
    rr, to_dev, back = (f16r, nhwc, nchw) if f16 else (bf16r, gnhwc, gnchw)
    tol = 1e-3 if f16 else 5e-3
    x = rr(torch.randn(N, cin, H, W, generator=g))
    w = rr(torch.randn(cout, cin, generator=g) / cin ** 0.5)
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w.view(cout, cin, 1, 1), b)
    out = MB.pw_conv(to_dev(x), wbits(w, f16), b.cuda(), 0, cout, f16=f16)
    assert rel(back(out), ref) < tol
    out = MB.pw_conv(to_dev(x), wbits(w, f16), b.cuda(), 1, cout, f16=f16)
    assert rel(back(out), F.hardswish(ref)) < tol
    raw, act = MB.pw_conv(to_dev(x), wbits(w, f16), None, 1, cout, want_raw=True, f16=f16)
    ref0 = F.conv2d(x, w.view(cout, cin, 1, 1))
    assert rel(back(raw), ref0) < tol
    assert torch.equal(back(act), rr(F.hardswish(back(raw))))


def test_pw_conv_residual_strided_input_and_per_sample_weights():
    from arbitrarystyletransfer_b200 import mobilenet as MB
    N, H, W, cin, cout = 2, 8, 12, 128, 128
    g = G(5)
    wide = f16r(torch.randn(N, 2 * cin, H, W, generator=g))
    w = f16r(torch.randn(N, cout, cin, generator=g) / cin ** 0.5)
    res = f16r(torch.randn(N, cout, H, W, generator=g))
    xw = nhwc(wide)
    x_view = xw[..., cin:]                      # channel slice: row stride 2*cin
    ref = torch.stack([F.conv2d(wide[n:n + 1, cin:], w[n].view(cout, cin, 1, 1))[0] for n in range(N)]) + res
    out = MB.pw_conv(x_view, wbits(w, True), None, 0, cout, residual=nhwc(res), per_sample=True, f16=True)
    assert rel(nchw(out), ref) < 1e-3
    # residual read through a nearest x2 upsample
    small = f16r(torch.randn(N, cout, H // 2, W // 2, generator=g))
    ref = F.conv2d(wide[:, :cin], w[0].view(cout, cin, 1, 1)) + F.interpolate(small, scale_factor=2, mode="nearest")
    out = MB.pw_conv(xw[..., :cin], wbits(w[0], True), None, 0, cout, residual=nhwc(small), res_up2=True, f16=True)
    assert rel(nchw(out), ref) < 1e-3
    # the same call on bf16 tensors (the role the backward pass uses: data gradient + residual gradient)
    wide_b, w_b, res_b = bf16r(wide), bf16r(w), bf16r(res)
    ref = torch.stack([F.conv2d(wide_b[n:n + 1, cin:], w_b[n].view(cout, cin, 1, 1))[0] for n in range(N)]) + res_b
    out = MB.pw_conv(gnhwc(wide_b)[..., cin:], wbits(w_b, False), None, 0, cout, residual=gnhwc(res_b), per_sample=True)
    assert rel(gnchw(out), ref) < 5e-3


@pytest.mark.parametrize("shape", [(2, 16, 16, 96, 16), (1, 12, 20, 16, 96), (2, 9, 7, 144, 24),
                                   (1, 16, 16, 240, 40), (2, 8, 8, 768, 256), (2, 8, 8, 128, 768),
                                   (4, 32, 32, 384, 128), (1, 3, 3, 320, 80)])
def test_pw_wgrad_mn_major_gemm(shape):
    """out[i][j] = sum_p a[p][i] b[p][j] on tcgen05 with both operands MN-major, vs fp32 matmul."""
    from arbitrarystyletransfer_b200 import mobilenet as MB
    N, H, W, ca, cb = shape
    g = G(sum(shape) + 1)
    a = bf16r(torch.randn(N, ca, H, W, generator=g))
    b = bf16r(torch.randn(N, cb, H, W, generator=g))
    ref = torch.einsum("nihw,njhw->ij", a.double(), b.double()).float()
    out = torch.zeros(ca, cb, device="cuda")
    MB._pw_wgrad(gnhwc(a), gnhwc(b), out, cb, 1)
    assert rel(out, ref) < 1e-3
    out_t = torch.zeros(cb, ca, device="cuda")
    MB._pw_wgrad(gnhwc(a), gnhwc(b), out_t, 1, ca)
    assert rel(out_t, ref.t()) < 1e-3
    # an fp16 ACTIVATION operand is converted to bf16 first (one format per tcgen05 instruction)
    out_a = torch.zeros(ca, cb, device="cuda")
    MB._pw_wgrad(nhwc(a), gnhwc(b), out_a, cb, 1, a_is_act=True)      # bf16-representable values survive both ways
    assert rel(out_a, ref) < 1e-3
    a16 = f16r(torch.randn(N, ca, H, W, generator=g))
    out_b = torch.zeros(cb, ca, device="cuda")
    MB._pw_wgrad(gnhwc(b), nhwc(a16), out_b, ca, 1, b_is_act=True)
    assert rel(out_b, torch.einsum("nihw,njhw->ji", a16.double(), b.double()).float()) < 4e-3   # one bf16 rounding of a


# ------------------------------------------------------------------------------------------------
# depthwise conv: forward (reflect, stride, virtual upsample, fused pool), data and weight gradients
# ------------------------------------------------------------------------------------------------
def _dw_ref(x, w, k, stride, up2):
    if up2:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    p = (k - 1) // 2
    return F.conv2d(F.pad(x, (p, p, p, p), mode="reflect"), w, stride=stride, groups=x.shape[1])


@pytest.mark.parametrize("cfg", [(2, 16, 10, 12, 3, 1, False), (1, 96, 16, 16, 3, 2, False), (2, 144, 9, 11, 5, 2, False),
                                 (1, 240, 8, 8, 5, 1, False), (2, 96, 6, 5, 3, 1, True), (1, 768, 4, 4, 3, 1, False),
                                 (1, 40, 3, 3, 5, 1, False), (1, 24, 2, 2, 3, 1, False),
                                 # large enough for the shared-memory-tiled stride-1 kernels (ragged tiles, borders)
                                 (2, 96, 20, 24, 5, 1, False), (1, 240, 33, 17, 5, 1, False), (2, 160, 16, 40, 3, 1, False),
                                 (1, 40, 12, 10, 3, 1, True), (1, 144, 64, 64, 3, 1, False), (1, 80, 10, 50, 5, 1, False)])
def test_dw_conv_forward_dgrad_wgrad(cfg):
    from arbitrarystyletransfer_b200 import mobilenet as MB, _lib as L
    N, C, H, W, k, stride, up2 = cfg
    g = G(sum(int(v) for v in cfg))
    x = f16r(torch.randn(N, C, H, W, generator=g)).requires_grad_(True)
    w = torch.randn(C, 1, k, k, generator=g).div_(k).requires_grad_(True)
    y = _dw_ref(x, w, k, stride, up2)
    dy = bf16r(torch.randn(y.shape, generator=g))
    y.backward(dy)
    wd = MB.prep_weight(w.detach().cuda(), C, k * k, 2)
    assert torch.equal(wd.cpu(), w.detach().view(C, k * k).t())
    out, pool = MB.dw_conv(nhwc(x.detach()), wd, None, k, stride, up2=up2, act=0, want_pool=True)
    assert rel(nchw(out), y.detach()) < 1e-3
    torch.testing.assert_close(pool.cpu(), MB.act_float(out).sum(dim=(1, 2)).cpu(), rtol=1e-4, atol=1e-3)
    out1, pool1 = MB.dw_conv(nhwc(x.detach()), wd, None, k, stride, up2=up2, act=1, want_pool=True)
    assert rel(nchw(out1), F.hardswish(y.detach())) < 1.5e-3
    out2, pool2 = MB.dw_conv(nhwc(x.detach()), wd, None, k, stride, up2=up2, act=2, want_pool=True)
    assert torch.equal(out2, out)
    torch.testing.assert_close(pool2.cpu(), F.hardswish(MB.act_float(out)).sum(dim=(1, 2)).cpu(), rtol=1e-4, atol=1e-3)
    lib = L.load()
    st = L.stream_ptr(out.device)
    dyd = gnhwc(dy)
    dx = torch.empty(N, H, W, C, device="cuda", dtype=torch.bfloat16)
    L.check(lib.ast_dw_conv_dgrad(dyd.data_ptr(), wd.data_ptr(), None, None, None, dx.data_ptr(), N, C, H, W, k,
                                  stride, int(up2), st))
    assert rel(gnchw(dx), x.grad) < 5e-3
    dw = torch.zeros(C, 1, k, k, device="cuda")
    L.check(lib.ast_dw_conv_wgrad(dyd.data_ptr(), nhwc(x.detach()).data_ptr(), dw.data_ptr(), N, C, H, W, k, stride,
                                  int(up2), st))
    assert rel(dw, w.grad) < 2e-3
    if stride == 1:   # identity branch folded into the data gradient
        dx2 = torch.empty_like(dx)
        L.check(lib.ast_dw_conv_dgrad(dyd.data_ptr(), wd.data_ptr(), None, None, dyd.data_ptr(), dx2.data_ptr(), N, C,
                                      H, W, k, stride, int(up2), st))
        extra = F.avg_pool2d(dy, 2) * 4 if up2 else dy
        assert rel(gnchw(dx2), x.grad + extra) < 5e-3


# ------------------------------------------------------------------------------------------------
# BatchNorm (batch statistics), Hardswish / SE passes and their backward kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(4, 96, 16, 16), (2, 16, 9, 7), (3, 240, 5, 5)])
def test_batchnorm_train_forward_backward(shape):
    from arbitrarystyletransfer_b200 import mobilenet as MB
    N, C, H, W = shape
    g = G(sum(shape))
    a = f16r(torch.randn(N, C, H, W, generator=g) * 2 + torch.randn(1, C, 1, 1, generator=g)).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g))
    bn_g = torch.nn.BatchNorm2d(C).cuda()
    bn_g.load_state_dict(bn.state_dict())
    y = bn(a)
    dy = bf16r(torch.randn(y.shape, generator=g))
    y.backward(dy)
    ad = nhwc(a.detach())
    stat = MB._bn_train_forward(ad, bn_g)
    yg, _ = MB.affine_act(ad, stat[2], stat[3], 0)
    assert rel(nchw(yg), y.detach()) < 1e-3
    torch.testing.assert_close(bn_g.running_mean.cpu(), bn.running_mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(bn_g.running_var.cpu(), bn.running_var, rtol=1e-4, atol=1e-5)
    assert int(bn_g.num_batches_tracked) == 1
    da, dgamma, dbeta = MB._bn_backward(gnhwc(dy), ad, stat)
    assert rel(gnchw(da), a.grad) < 6e-3
    assert rel(dgamma, bn.weight.grad) < 1e-3 and rel(dbeta, bn.bias.grad) < 1e-3


@pytest.mark.parametrize("norm", [False, True])
def test_se_hardswish_norm_backward_chain(norm):
    """u = Hardswish(BN(a)) * SE(mean Hardswish(BN(a))): forward pieces and the fused backward
    (dw_bwd_reduce -> se_bwd -> se_bn_combine -> dw_bwd_apply) vs torch autograd."""
    from arbitrarystyletransfer_b200 import mobilenet as MB, _lib as L
    N, C, H, W, S = 3, 96, 8, 8, 24
    g = G(7 + norm)
    a = f16r(torch.randn(N, C, H, W, generator=g) * 2).requires_grad_(True)
    w1 = (torch.randn(S, C, generator=g) * 0.3).requires_grad_(True)
    b1 = (torch.randn(S, generator=g) * 0.1).requires_grad_(True)
    w2 = (torch.randn(C, S, generator=g) * 0.3).requires_grad_(True)
    b2 = (torch.rand(C, generator=g)).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.5)
    z = bn(a) if norm else a
    h = F.hardswish(z)
    pool = h.mean(dim=(2, 3))
    s = F.hardtanh(F.linear(F.relu(F.linear(pool, w1, b1)), w2, b2), 0.0, 1.0)
    u = h * s.view(N, C, 1, 1)
    du = bf16r(torch.randn(u.shape, generator=g))
    pool.retain_grad()
    s.retain_grad()
    u.backward(du)

    lib = L.load()
    ad = nhwc(a.detach())
    st = L.stream_ptr(ad.device)
    stat = None
    if norm:
        bn_g = torch.nn.BatchNorm2d(C).cuda()
        bn_g.load_state_dict(bn.state_dict())
        stat = MB._bn_train_forward(ad, bn_g)
    sc, sh = (stat[2], stat[3]) if norm else (None, None)
    _, pool_g = MB.affine_act(ad, sc, sh, 1, want_out=False, want_pool=True)
    torch.testing.assert_close(pool_g.cpu() / (H * W), pool.detach(), rtol=2e-3, atol=2e-3)
    cw = [t.detach().cuda().contiguous() for t in (w1, b1, w2, b2)]
    s_g, hid, pre = MB.se_fc(pool_g, 1.0 / (H * W), *cw, save=True)
    torch.testing.assert_close(s_g.cpu(), s.detach(), rtol=2e-3, atol=2e-3)
    u_g, _ = MB.affine_act(ad, sc, sh, 1, se=s_g)
    assert rel(nchw(u_g), u.detach()) < 2e-3
    dud = gnhwc(du)
    T = torch.empty(N, 5, C, device="cuda")
    L.check(lib.ast_dw_bwd_reduce(dud.data_ptr(), ad.data_ptr(), L.ptr(stat), T.data_ptr(), N, C, H * W, st))
    dpre, dhid, gg = torch.empty(N, C, device="cuda"), torch.empty(N, S, device="cuda"), torch.empty(N, C, device="cuda")
    gw1, gb1, gw2, gb2 = (torch.empty_like(t) for t in cw)
    L.check(lib.ast_se_bwd(T.data_ptr(), 5 * C, pre.data_ptr(), hid.data_ptr(), pool_g.data_ptr(), 1.0 / (H * W),
                           cw[0].data_ptr(), cw[2].data_ptr(), dpre.data_ptr(), dhid.data_ptr(), gg.data_ptr(),
                           gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(), gb2.data_ptr(), N, C, S, st))
    assert rel(T[:, 0], s.grad) < 1e-3, rel(T[:, 0], s.grad)                 # d loss / d gate
    assert rel(gg * (H * W), pool.grad) < 1e-3, rel(gg * (H * W), pool.grad)   # d loss / d pooled mean
    for got, want in ((gw1, w1.grad), (gb1, b1.grad), (gw2, w2.grad), (gb2, b2.grad)):
        assert rel(got, want) < 1e-3, rel(got, want)
    coef = None
    if norm:
        coef = torch.empty(2, C, device="cuda")
        dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        L.check(lib.ast_se_bn_combine(T.data_ptr(), s_g.data_ptr(), gg.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                      coef.data_ptr(), N, C, float(N * H * W), st))
        assert rel(dg, bn.weight.grad) < 1e-2 and rel(db, bn.bias.grad) < 1e-2
    da = torch.empty_like(ad)
    L.check(lib.ast_dw_bwd_apply(dud.data_ptr(), ad.data_ptr(), s_g.data_ptr(), gg.data_ptr(), L.ptr(stat),
                                 L.ptr(coef), da.data_ptr(), N, C, H * W, st))
    assert rel(gnchw(da), a.grad) < 1e-2, rel(gnchw(da), a.grad)


def test_stem_and_head_forward_backward():
    from arbitrarystyletransfer_b200 import mobilenet as MB
    g = G(11)
    N, H, W = 2, 12, 10
    img = torch.rand(N, 3, H, W, generator=g)
    w = (torch.randn(16, 3, 3, 3, generator=g) * 0.4).requires_grad_(True)
    y = F.hardswish(F.conv2d(F.pad(img, (1, 1, 1, 1), mode="reflect"), w))
    dy = bf16r(torch.randn(y.shape, generator=g))
    y.backward(dy)
    wg = w.detach().cuda().requires_grad_(True)
    yg = MB._StemFn.apply(img.cuda(), wg)
    assert rel(nchw(yg), y.detach()) < 1e-3
    yg.backward(gnhwc(dy))
    assert rel(wg.grad, w.grad) < 1e-2
    # head: ReflectionPad2d(1) + Conv2d(16, 3, 3) with bias
    x = f16r(torch.randn(N, 16, H, W, generator=g)).requires_grad_(True)
    hw = (torch.randn(3, 16, 3, 3, generator=g) * 0.2).requires_grad_(True)
    hb = torch.randn(3, generator=g).requires_grad_(True)
    out = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), hw, hb)
    dY = torch.randn(out.shape, generator=g)
    out.backward(dY)
    xg = nhwc(x.detach()).requires_grad_(True)
    hwg, hbg = hw.detach().cuda().requires_grad_(True), hb.detach().cuda().requires_grad_(True)
    og = MB._HeadFn.apply(xg, hwg, hbg, False)
    assert rel(og, out.detach()) < 1e-4
    og.backward(dY.cuda())
    assert rel(hwg.grad, hw.grad) < 1e-4 and rel(hbg.grad, hb.grad) < 1e-4
    assert rel(gnchw(xg.grad), x.grad) < 5e-3
    assert rel(MB.nhwc_to_nchw(MB.nchw_to_nhwc(out.detach().cuda())), bf16r(out.detach())) == 0.0
    assert rel(MB.to_nchw(MB.to_nhwc(out.detach().cuda())), f16r(out.detach())) == 0.0
    big = torch.tensor([1e6, -1e6, 65504.0, 70000.0]).view(1, 4, 1, 1).expand(1, 4, 2, 2).contiguous()
    assert torch.equal(MB.to_nchw(MB.to_nhwc(big.cuda())).cpu(), big.clamp(-65504, 65504))   # saturates, never inf


# ------------------------------------------------------------------------------------------------
# whole blocks against the oracle (CPU fp32 autograd of oracle/restate_ae.py::depthwise_block)
# ------------------------------------------------------------------------------------------------
BLOCKS = [  # inp, oup, stride, t, k, norm, identity, up2, H, W
    (16, 16, 1, 6, 3, True, True, False, 16, 16),
    (24, 40, 2, 6, 5, True, True, False, 12, 16),
    (96, 96, 1, 3, 5, False, True, False, 8, 8),
    (40, 24, 1, 6, 5, False, True, False, 10, 6),
    (96, 96, 1, 1, 3, False, True, True, 6, 8),
    (256, 128, 1, 3, 3, False, False, False, 4, 4),
]


def _block_state(blk, prefix="b"):
    return {f"{prefix}.{k}": v.detach().cpu().clone() for k, v in blk.state_dict().items()}


@pytest.mark.parametrize("cfg", BLOCKS)
def test_block_train_forward_backward_vs_oracle(cfg):
    from arbitrarystyletransfer_b200 import mobilenet as MB
    inp, oup, stride, t, k, norm, ident, up2, H, W = cfg
    torch.manual_seed(sum(int(v) for v in cfg))
    blk = MB.DepthWiseConv(inp, oup, stride, t, kernel_size=k, use_norm=norm, use_identity=ident)
    # Make the SE gate and the norms non-trivial (the N(0, 0.01) / zero-bias init leaves the gate ~0), but keep
    # the gate's pre-activations strictly inside the ReLU / Hardtanh linear regions: a bf16-level perturbation
    # of the pooled mean must not flip a mask, which would change the gradient of a whole sample by ~10 % and
    # says nothing about kernel correctness (the masks themselves are tested on exact inputs above).
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.Linear):
                m.weight.normal_(0, 0.05)
                m.bias.uniform_(0.3, 0.7)
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.3)
    sd = _block_state(blk)
    P = A.clone_state(sd, requires_grad=True)
    N = 3
    x = f16r(torch.randn(N, inp, H, W, generator=G(3))).requires_grad_(True)
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if up2 else x
    ref = A.depthwise_block(P, "b", xin, inp, oup, stride, t, k, norm=norm, use_identity=ident, training=True)
    dy = bf16r(torch.randn(ref.shape, generator=G(4)))
    ref.backward(dy)

    blk = blk.cuda().train()
    xg = nhwc(x.detach()).requires_grad_(True)
    out = blk.forward_nhwc(xg, up2=up2)
    assert rel(nchw(out), ref.detach()) < 3e-3, rel(nchw(out), ref.detach())
    out.backward(gnhwc(dy))
    assert rel(gnchw(xg.grad), x.grad) < 2e-2, rel(gnchw(xg.grad), x.grad)
    # Hardswish' jumps by 1/2 at +-3; a pre-activation within bf16 rounding of a kink (a few elements per
    # ten thousand) takes the other branch than the fp32 oracle and moves the BatchNorm gamma / beta gradient
    # of ITS channel by ~20 % when only ~150 elements feed that channel, as in these small cases.  Those two
    # get a wider relative-L2 bar; direction (cosine) is held tight for everything.
    for name, p in blk.named_parameters():
        want = P["b." + name].grad
        r, c = rel(p.grad, want), cos(p.grad, want)
        is_norm = p.dim() == 1 and ".fc." not in name
        assert r < (1e-1 if is_norm else 4e-2) and c > (0.995 if is_norm else 0.999), (name, r, c)
    if norm:   # running statistics updated exactly like nn.BatchNorm2d
        for name, b in blk.named_buffers():
            torch.testing.assert_close(b.cpu().float(), P["b." + name].float(), rtol=2e-2, atol=2e-3)


# ------------------------------------------------------------------------------------------------
# the whole autoencoder against the goldens made from the genuine reference
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def g():
    return load_golden("autoencoder")


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture()
def ae():
    from arbitrarystyletransfer_b200 import mobilenet as MB
    torch.manual_seed(2)
    return MB.AutoEncoder().cuda()


@pytest.fixture()
def act_format():
    """set the process-wide activation format for one test and restore the default (fp16) afterwards"""
    from arbitrarystyletransfer_b200 import mobilenet as MB

    def use(fmt):
        MB.set_activation_format(fmt)
    yield use
    MB.set_activation_format("fp16")


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
def test_autoencoder_eval_forward_vs_reference_golden(g, ae, act_format, fmt):
    """The reference's FRESH initialisation in eval mode: closed SE gates and fresh running statistics shrink the
    activations by ~1e-4 per block (1e-8 after two blocks), the image is exactly the head bias.  bf16 activations
    (fp32's exponent range) follow the fp32 reference all the way down; fp16 activations bottom out at their subnormal
    quantum 2^-24 = 6e-8, i.e. they are exact to that ABSOLUTE error -- the documented price of the 11-bit
    significand (mobilenet.set_activation_format)."""
    act_format(fmt)
    x = T(g["ae_x"]).cuda()
    ae.eval()
    with torch.no_grad():
        recon = ae(x)
        taps = ae.encoder(x, out_layers=[0, 2, 12, 14])
        z = ae.encoder(x, auto_enc=True)
    for i, t in list(zip((0, 2, 12, 14), taps)) + [("z", z)]:
        ref = T(g[f"ae_eval_enc{i}"]) if i != "z" else T(g["ae_eval_autoenc"])
        if fmt == "bf16":
            assert rel(t, ref) < 3e-2, (i, rel(t, ref))
        else:
            err = (t.cpu() - ref).abs().max().item()
            assert err <= max(2e-3 * ref.abs().max().item(), 2.0 ** -24), (i, err, ref.abs().max().item())
    ref = T(g["ae_eval_recon_fresh"])
    assert rel(recon, ref) < (3e-2 if fmt == "bf16" else 1e-5), rel(recon, ref)
    assert R.psnr(recon.cpu(), ref) >= 40.0


def test_bf16_activation_format_train_step(act_format):
    """The bf16 activation format (range over precision) still runs the whole training path: one Huber step on the
    non-degenerate state against the oracle, at the looser bars that format earns (round-1 numbers)."""
    from arbitrarystyletransfer_b200 import mobilenet as MB
    act_format("bf16")
    assert MB.activation_format() == "bf16"
    sd = A.activate_gates(A.make_ae_state(2))
    torch.manual_seed(2)
    net = MB.AutoEncoder().cuda().train()
    net.load_state_dict(sd, strict=True)
    x = R.rand_image(2, 64, 301)
    loss = F.huber_loss(net(x.cuda()), x.cuda())
    loss.backward()
    P = A.clone_state(sd, requires_grad=True)
    ref_loss = F.huber_loss(A.autoencoder_forward(P, x, training=True), x)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    named = dict(net.named_parameters())
    for k in ("decoder._img_out.weight", "ada_out._layers.0.weight", "encoder.mob_net.4._layers.3.weight",
              "encoder.mob_net.0.0.weight"):
        assert cos(named[k].grad, P[k].grad) > 0.95, (k, cos(named[k].grad, P[k].grad))


def test_autoencoder_train_step_vs_reference_golden(g, ae):
    """One train_autoencoder.py:111-139 step: train-mode forward (batch statistics), loss, gradients of
    every parameter (norms) and of the stored subset (full tensors), running statistics, and the eval
    forward with the updated statistics."""
    from arbitrarystyletransfer_b200 import models as M
    from arbitrarystyletransfer_b200.losses import compute_content_loss
    x = T(g["ae_x"]).cuda()
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    enc = M.PretrainedEncoder().cuda().eval()
    with torch.no_grad():
        for c, w, b in zip(enc._convs(), vw, vb):
            c.weight.copy_(w)
            c.bias.copy_(b)
    for p in enc.parameters():
        p.requires_grad_(False)
    ae.train()
    recon = ae(x)
    assert rel(recon, T(g["ae_train_recon"])) < 4e-2, rel(recon, T(g["ae_train_recon"]))
    recon_loss = compute_content_loss(recon, x)
    with torch.no_grad():
        cm = enc(x)
    rm = enc(recon)
    perp = None
    for a, b in zip(rm, cm):
        l = compute_content_loss(a, b.detach())
        perp = l if perp is None else perp + l
    loss = 100.0 * recon_loss + 0.01 * perp
    want = g["ae_train_losses"]
    assert abs(recon_loss.item() - want[1]) / want[1] < 2e-2
    assert abs(perp.item() - want[2]) / want[2] < 5e-2
    loss.backward()
    named = dict(ae.named_parameters())
    gkeys = list(g["ae_grad_keys"])
    norms = np.array([named[k].grad.double().norm().item() for k in gkeys])
    ratio = norms / np.maximum(g["ae_grad_norm"], 1e-12)
    big = g["ae_grad_norm"] > 1e-4 * g["ae_grad_norm"].max()
    assert np.all(np.abs(ratio[big] - 1) < 0.15), (np.array(gkeys)[big][np.abs(ratio[big] - 1) >= 0.15], ratio[big])
    for k in A.GOLDEN_GRAD_KEYS:
        want = T(g["ae_grad::" + k])
        if want.norm() < 1e-4 * g["ae_grad_norm"].max():
            continue
        assert cos(named[k].grad, want) > 0.99, (k, cos(named[k].grad, want), rel(named[k].grad, want))
    sd = ae.state_dict()
    for k in A.GOLDEN_BUFFER_KEYS:
        torch.testing.assert_close(sd[k].cpu().float(), T(g["ae_buf::" + k]).float(), rtol=3e-2, atol=3e-3)
    ae.eval()
    with torch.no_grad():
        after = ae(x)
    assert rel(after, T(g["ae_eval_recon_after_step"])) < 4e-2


def _signal(img, bias):
    """image minus the head bias: what the decoder actually produced (the bias alone is ~30x larger)"""
    return img.cpu() - bias.view(1, -1, 1, 1).cpu()


def test_autoencoder_non_degenerate_state_vs_reference_golden(g, ae):
    """The whole network on the NON-DEGENERATE variant of the seeded state (oracle/restate_ae.py::activate_gates:
    the reference's fresh initialisation outputs exactly the head bias and has zero gradients almost everywhere,
    so the fresh-state tests above pin little beyond the encoder).  Fixtures made by the genuine reference:
    train-mode forward, the train_autoencoder.py loss, gradients of all 308 trainable tensors, BatchNorm running
    statistics, and the eval-mode forward with calibrated running statistics."""
    from arbitrarystyletransfer_b200 import models as M
    from arbitrarystyletransfer_b200.losses import compute_content_loss
    x = T(g["ae_x"]).cuda()
    act = A.activate_gates(A.make_ae_state(2))
    ae.load_state_dict(act, strict=True)
    bias = act["decoder._img_out.bias"]
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    enc = M.PretrainedEncoder().cuda().eval()
    with torch.no_grad():
        for c, w, b in zip(enc._convs(), vw, vb):
            c.weight.copy_(w)
            c.bias.copy_(b)
    for p in enc.parameters():
        p.requires_grad_(False)
    ae.train()
    recon = ae(x)
    want = T(g["act_train_recon"])
    # Train mode at this fixture size (2 x 32 x 32: 32 samples per channel at the deepest BatchNorms) amplifies
    # rounding.  With the round-1 bf16 activations the decoder signal sat 11 % from the reference (error growing
    # steadily through the 14 encoder blocks, profiles/r1_ae_error_growth.txt); with fp16 activations it is 1.3 %.
    sig = rel(_signal(recon.detach(), bias), _signal(want, bias))
    print(f"train-mode decoder signal rel L2 vs the reference: {sig:.3e}")
    assert sig < 2.5e-2, sig
    recon_loss = compute_content_loss(recon, x)
    with torch.no_grad():
        cm = enc(x)
    perp = None
    for a, b in zip(enc(recon), cm):
        l = compute_content_loss(a, b.detach())
        perp = l if perp is None else perp + l
    loss = 100.0 * recon_loss + 0.01 * perp
    wl = g["act_train_losses"]
    assert abs(recon_loss.item() - wl[1]) / wl[1] < 1e-3 and abs(perp.item() - wl[2]) / wl[2] < 2e-2
    loss.backward()
    named = dict(ae.named_parameters())
    gkeys = list(g["ae_grad_keys"])
    norms = np.array([named[k].grad.double().norm().item() for k in gkeys])
    # BatchNorm betas (and the biases in front of a train-mode BatchNorm) have an exactly cancelling gradient: the
    # reference's value for them is fp32 round-off (1e-9), ours bf16 round-off; only meaningful norms are compared
    big = g["act_grad_norm"] > 1e-4 * g["act_grad_norm"].max()
    assert big.sum() > 250
    ratio = (norms / np.maximum(g["act_grad_norm"], 1e-30))[big]
    gkeys = list(np.array(gkeys)[big])
    report = [(k, round(cos(named[k].grad, T(g["act_grad::" + k])), 4)) for k in A.GOLDEN_GRAD_KEYS]
    print("gradient-norm ratio min/median/max:", ratio.min(), np.median(ratio), ratio.max(), "cosines:", report)
    assert np.all(np.abs(ratio - 1) < 0.05), (np.array(gkeys)[np.abs(ratio - 1) >= 0.05], ratio[np.abs(ratio - 1) >= 0.05])
    assert abs(np.median(ratio) - 1) < 0.01
    assert min(c for _, c in report) > 0.995, report
    sd = ae.state_dict()
    for k in A.GOLDEN_BUFFER_KEYS:
        torch.testing.assert_close(sd[k].cpu().float(), T(g["act_buf::" + k]).float(), rtol=3e-2, atol=3e-3)
    # eval mode with calibrated running statistics (loaded from the oracle, which reproduces the reference bit-exactly)
    Q = A.calibrate_running_stats(A.clone_state(act), x.cpu())
    ae.load_state_dict(Q, strict=True)
    ae.eval()
    with torch.no_grad():
        rec = ae(x)
        taps = ae.encoder(x, out_layers=[0, 2, 12, 14])
        z = ae.ada_out(torch.cat((taps[2], taps[3]), dim=1))
        img = ae.decoder(T(g["act_eval_code"]).cuda())
    errs = {f"enc{i}": rel(t, T(g[f"act_eval_enc{i}"])) for i, t in zip((0, 2, 12, 14), taps)}
    errs["code"] = rel(z, T(g["act_eval_code"]))
    errs["decoder(code)"] = rel(_signal(img, bias), _signal(T(g["act_eval_dec_of_code"]), bias))
    errs["recon"] = rel(_signal(rec, bias), _signal(T(g["act_eval_recon"]), bias))
    print("eval-mode relative L2 errors:", {k: round(v, 4) for k, v in errs.items()})
    # ~150 fp16 roundings in series: shallow taps at 1e-3, the deepest 4x4 features at 1 %; the decoder alone (fed the
    # exact code) is held separately
    assert errs["enc0"] < 1e-3 and errs["enc2"] < 3e-3, errs
    assert max(errs["enc12"], errs["enc14"], errs["code"]) < 2e-2, errs
    assert errs["decoder(code)"] < 5e-3 and errs["recon"] < 2e-2, errs
    assert R.psnr(rec.cpu(), T(g["act_eval_recon"])) >= 40.0
    # The kernels against THEIR OWN arithmetic contract (oracle/restate_ae.py::autoencoder_forward_bf16: fp32 math,
    # bf16 rounding at exactly the kernels' storage points).  The contract itself sits 0.16 % / 0.87 % / 8.1 % /
    # 10.1 % from the fp32 reference at enc0 / enc2 / enc12 / enc14 on this state -- the same curve as measured
    # above -- so that distance is the price of bf16 storage, not of the kernels.
    with torch.no_grad():
        cimg, keep = A.autoencoder_forward_contract(Q, x.cpu(), want=("enc0", "enc2", "enc12", "enc14", "code"))
    cerr = {k: rel(t, keep[k]) for k, t in zip(("enc0", "enc2", "enc12", "enc14"), taps)}
    cerr["code"] = rel(z, keep["code"])
    cerr["recon"] = rel(_signal(rec, bias), _signal(cimg, bias))
    print("relative L2 vs the bf16 storage contract:", {k: round(v, 4) for k, v in cerr.items()})
    assert cerr["enc0"] < 2e-4 and cerr["enc2"] < 1e-3, cerr
    # deep features: two bf16 pipelines that differ in fp32 summation order (atomics, MMA tree) drift apart at
    # the same rate as either drifts from fp32 (measured 3-4.5 % at the 4x4 maps)
    assert max(cerr.values()) < 1.2e-2, cerr


@pytest.mark.parametrize("hw", [(40, 24), (72, 56), (24, 88)])
def test_autoencoder_ragged_sizes_vs_contract(ae, hw):
    """Sizes that are multiples of 8 but of nothing else (deepest maps 5x3, 9x7, 3x11): partial tiles, TMA zero
    fill, odd extents under the stride-2 convs, reflection on 2- and 3-pixel maps.  Eval forward and train-mode
    forward + backward against the oracle (fp32 and the bf16 storage contract)."""
    H, W = hw
    x = torch.rand(3, 3, H, W, generator=G(H * W))
    act = A.activate_gates(A.make_ae_state(2))
    Q = A.calibrate_running_stats(A.clone_state(act), x)
    ae.load_state_dict(Q, strict=True)
    bias = act["decoder._img_out.bias"]
    ae.eval()
    with torch.no_grad():
        rec = ae(x.cuda())
        cimg, _ = A.autoencoder_forward_contract(Q, x)
        ref = A.autoencoder_forward(Q, x)
    assert rec.shape == ref.shape and torch.isfinite(rec).all()
    e_contract, e_fp32 = rel(_signal(rec, bias), _signal(cimg, bias)), rel(_signal(rec, bias), _signal(ref, bias))
    print(f"ragged {H}x{W}: image signal vs contract {e_contract:.4f}, vs fp32 {e_fp32:.4f}")
    assert e_contract < 1.5e-2 and e_fp32 < 2.5e-2 and R.psnr(rec.cpu(), ref) >= 40.0, (e_contract, e_fp32)
    # train mode: loss and a few gradients against the oracle's autograd
    ae.load_state_dict(act, strict=True)
    ae.train()
    loss = F.huber_loss(ae(x.cuda()), x.cuda())
    loss.backward()
    P = A.clone_state(act, requires_grad=True)
    ref_loss = F.huber_loss(A.autoencoder_forward(P, x, training=True), x)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    named = dict(ae.named_parameters())
    for k in ("decoder._img_out.weight", "decoder._decoder_blocks.8._conv._layers.2.weight",
              "decoder._decoder_blocks.4._upsample_2._layers.1.weight", "ada_out._layers.0.weight",
              "encoder.mob_net.7._layers.3.weight", "encoder.mob_net.2._layers.0.weight", "encoder.mob_net.0.0.weight"):
        c = cos(named[k].grad, P[k].grad)
        assert c > 0.99, (k, c)


def test_autoencoder_too_small_input_raises_like_the_reference(ae):
    """16 x 48 leaves a 2-pixel map under a 5x5 reflect-padded depthwise conv: torch raises in the reference
    ("Padding size should be less than the corresponding input dimension"); this path must refuse too."""
    from arbitrarystyletransfer_b200 import _lib as L
    x = torch.rand(1, 3, 16, 48, generator=G(1))
    with pytest.raises(RuntimeError):
        A.autoencoder_forward(A.clone_state(A.make_ae_state(2)), x)
    ae.eval()
    with pytest.raises(L.AstError), torch.no_grad():
        ae(x.cuda())
        torch.cuda.synchronize()


def test_autoencoder_config3_shape_properties(ae):
    """BASELINE config 3 geometry (256x256) at a reduced batch: finite outputs, determinism, every
    parameter receives a finite gradient, BatchNorm-normalised activations have the batch statistics they
    must have, and a few Adam steps on one batch reduce the reconstruction loss."""
    x = R.rand_image(4, 256, 301).cuda()
    ae.train()
    out1 = ae(x)
    assert out1.shape == (4, 3, 256, 256) and torch.isfinite(out1).all()
    loss = F.huber_loss(out1, x)
    loss.backward()
    for n, p in ae.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    opt = torch.optim.Adam(ae.parameters(), lr=2e-4, betas=(0.9, 0.99), eps=1e-7)
    first = None
    for _ in range(6):
        opt.zero_grad()
        l = F.huber_loss(ae(x), x)
        l.backward()
        torch.nn.utils.clip_grad_norm_(ae.parameters(), 10.0)
        opt.step()
        first = first if first is not None else l.item()
    assert l.item() < first
    ae.eval()
    with torch.no_grad():
        a, b = ae(x), ae(x)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------
# BASELINE config 3 AT ITS OWN RESOLUTION (256 x 256, batch 2): fixtures made by the genuine reference AutoEncoder
# (oracle/make_golden.py::golden_autoencoder256) on the non-degenerate seeded state.
# ------------------------------------------------------------------------------------------------
def _sub(t):
    st = max(1, t.shape[2] // 32)
    return t[:, :, ::st, ::st]


# bars for the whole-network comparisons at 256 x 256 (relative L2 unless noted); see DESIGN.md section 5
AE256_TAP_BAR = 2e-2          # every encoder tap, the code and the decoder signal, eval mode
AE256_TRAIN_BAR = 2e-2        # decoder signal of the train-mode forward (batch statistics)
AE256_GRAD_COS = 0.993        # cosine of every GOLDEN_GRAD_KEYS gradient
AE256_NORM_BAND = 0.07        # |gradient-norm ratio - 1| over all non-negligible parameters


def test_autoencoder_256_eval_vs_reference_golden(golden_ae256, ae):
    g = golden_ae256
    x = R.rand_image(2, 256, 301)
    act = A.activate_gates(A.make_ae_state(2))
    Q = A.calibrate_running_stats(A.clone_state(act), x)
    ae.load_state_dict(Q, strict=True)
    ae.eval()
    bias = act["decoder._img_out.bias"]
    with torch.no_grad():
        rec = ae(x.cuda())
        taps = ae.encoder(x.cuda(), out_layers=[0, 2, 4, 7, 12, 14])
        z = ae.ada_out(torch.cat((taps[4], taps[5]), dim=1))
        dimg = ae.decoder(T(g["e256_code"]).cuda())
    errs = {f"enc{i}": rel(_sub(t), T(g[f"e256_enc{i}_sub"])) for i, t in zip((0, 2, 4, 7, 12, 14), taps)}
    errs["code"] = rel(z, T(g["e256_code"]))
    errs["decoder(code)"] = rel(_signal(dimg[:, :, ::4, ::4], bias), _signal(T(g["e256_dec_of_code_sub4"]), bias))
    errs["recon signal"] = rel(_signal(rec[:, :, ::4, ::4], bias), _signal(T(g["e256_recon_sub4"]), bias))
    errs["recon"] = rel(rec[:, :, ::4, ::4], T(g["e256_recon_sub4"]))
    print("AutoEncoder 256x256 eval, relative L2 vs the genuine reference:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < AE256_TAP_BAR, errs
    assert errs["enc0"] < 5e-3 and errs["recon"] < 1e-2, errs
    assert R.psnr(rec[:, :, 96:160, 96:160].cpu(), T(g["e256_recon_crop"])) >= 40.0


def test_autoencoder_256_train_step_vs_reference_golden(golden_ae256, ae):
    """One train_autoencoder.py:111-139 step at 256 x 256, batch 2: loss terms, every parameter's gradient norm, full
    gradients of the GOLDEN_GRAD_KEYS, BatchNorm running statistics."""
    from arbitrarystyletransfer_b200 import models as M
    from arbitrarystyletransfer_b200.losses import compute_content_loss
    g = golden_ae256
    x = R.rand_image(2, 256, 301).cuda()
    act = A.activate_gates(A.make_ae_state(2))
    ae.load_state_dict(act, strict=True)
    bias = act["decoder._img_out.bias"]
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    enc = M.PretrainedEncoder().cuda().eval()
    with torch.no_grad():
        for c, w, b in zip(enc._convs(), vw, vb):
            c.weight.copy_(w)
            c.bias.copy_(b)
    for p in enc.parameters():
        p.requires_grad_(False)
    ae.train()
    for p in ae.parameters():
        p.grad = None
    recon = ae(x)
    sig = rel(_signal(recon.detach()[:, :, ::4, ::4], bias), _signal(T(g["t256_recon_sub4"]), bias))
    recon_loss = compute_content_loss(recon, x)
    with torch.no_grad():
        cm = enc(x)
    perp = None
    for a, b in zip(enc(recon), cm):
        l = compute_content_loss(a, b.detach())
        perp = l if perp is None else perp + l
    loss = 100.0 * recon_loss + 0.01 * perp
    wl = g["t256_losses"]
    loss.backward()
    named = dict(ae.named_parameters())
    gkeys = list(g["t256_grad_keys"])
    norms = np.array([named[k].grad.double().norm().item() for k in gkeys])
    big = g["t256_grad_norm"] > 1e-4 * g["t256_grad_norm"].max()
    ratio = (norms / np.maximum(g["t256_grad_norm"], 1e-30))[big]
    report = {k.replace("encoder.mob_net.", "e").replace("decoder._decoder_blocks.", "d").replace("._layers.", ".L"):
              (round(cos(named[k].grad, T(g["t256_grad::" + k])), 4), round(rel(named[k].grad, T(g["t256_grad::" + k])), 4))
              for k in A.GOLDEN_GRAD_KEYS if T(g["t256_grad::" + k]).norm() > 1e-4 * g["t256_grad_norm"].max()}
    print(f"AutoEncoder 256x256 train step: recon signal rel {sig:.2e}; losses cuda ({recon_loss.item():.6f}, "
          f"{perp.item():.5f}) reference ({wl[1]:.6f}, {wl[2]:.5f}); gradient-norm ratio min/median/max "
          f"{ratio.min():.3f}/{np.median(ratio):.3f}/{ratio.max():.3f} over {int(big.sum())} tensors; "
          f"(cosine, rel L2) per golden gradient: {report}")
    assert sig < AE256_TRAIN_BAR
    assert abs(recon_loss.item() - wl[1]) / wl[1] < 2e-3 and abs(perp.item() - wl[2]) / wl[2] < 2e-2
    assert np.all(np.abs(ratio - 1) < AE256_NORM_BAND), (np.array(gkeys)[big][np.abs(ratio - 1) >= AE256_NORM_BAND])
    assert abs(np.median(ratio) - 1) < 0.03
    assert min(c for c, _ in report.values()) > AE256_GRAD_COS, report
    sd = ae.state_dict()
    for k in A.GOLDEN_BUFFER_KEYS:
        torch.testing.assert_close(sd[k].cpu().float(), T(g["t256_buf::" + k]).float(), rtol=2e-2, atol=2e-3)


def test_autoencoder_per_batch_resolutions_of_the_reference_loader(ae):
    """the reference's loader draws a new (h, w) from conf.img_sizes = [96, 128, 160] for every
    batch (data_loader.py:89-97, conf.py:4).  Consecutive training steps at changing, non-square resolutions -- no
    rebuild, no stale shape-keyed state -- against the CPU oracle running the same steps; then eval mode at each size."""
    sizes = [(96, 96), (128, 160), (160, 96), (96, 128), (96, 96)]
    sd = A.activate_gates(A.make_ae_state(2))
    ae.load_state_dict(sd, strict=True)
    ae.train()
    opt = torch.optim.Adam(ae.parameters(), lr=2e-4, betas=(0.9, 0.99), eps=1e-7)
    P = A.clone_state(sd, requires_grad=True)
    train = [P[k] for k in sorted(P) if P[k].requires_grad]
    opt_r = torch.optim.Adam(train, lr=2e-4, betas=(0.9, 0.99), eps=1e-7)
    for i, (h, w) in enumerate(sizes):
        x = torch.rand(2, 3, h, w, generator=G(400 + i))
        opt.zero_grad()
        loss = F.huber_loss(ae(x.cuda()), x.cuda())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ae.parameters(), 10.0)
        opt.step()
        opt_r.zero_grad()
        ref = F.huber_loss(A.autoencoder_forward(P, x, training=True), x)
        ref.backward()
        torch.nn.utils.clip_grad_norm_(train, 10.0)
        opt_r.step()
        assert loss.item() == pytest.approx(ref.item(), rel=2e-3), (i, h, w, loss.item(), ref.item())
    ae.eval()
    Q = {k: v.detach().clone() for k, v in P.items()}
    bias = Q["decoder._img_out.bias"]
    for (h, w) in sizes[:3]:
        x = torch.rand(1, 3, h, w, generator=G(h + w))
        with torch.no_grad():
            ae.load_state_dict(Q, strict=True)
            got = ae(x.cuda())
            want = A.autoencoder_forward(Q, x)
        assert got.shape == want.shape
        assert rel(_signal(got, bias), _signal(want, bias)) < 3e-2 and R.psnr(got.cpu(), want) >= 40.0


# ------------------------------------------------------------------------------------------------
# the three gaps round 1 left on rows a7 / a8: eval-mode BatchNorm WITH gradients, the gradient with respect to the
# Encoder's input image, and the backward of the exporting (Hardtanh) head
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [BLOCKS[0], BLOCKS[1]])
def test_block_eval_mode_batchnorm_with_gradients_vs_oracle(cfg):
    """nn.BatchNorm2d in eval mode inside a differentiated block: running statistics instead of batch statistics, no
    running-stat update, da = dy * scale (mobilenetv2.py:108, 128, 137, 149 under ``.eval()``)."""
    from arbitrarystyletransfer_b200 import mobilenet as MB
    inp, oup, stride, t, k, norm, ident, up2, H, W = cfg
    torch.manual_seed(sum(int(v) for v in cfg) + 1)
    blk = MB.DepthWiseConv(inp, oup, stride, t, kernel_size=k, use_norm=True, use_identity=ident)
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.Linear):
                m.weight.normal_(0, 0.05)
                m.bias.uniform_(0.3, 0.7)
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.3)
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 2.0)
    sd = _block_state(blk)
    P = A.clone_state(sd, requires_grad=True)
    x = f16r(torch.randn(3, inp, H, W, generator=G(3))).requires_grad_(True)
    ref = A.depthwise_block(P, "b", x, inp, oup, stride, t, k, norm=True, use_identity=ident, training=False)
    dy = bf16r(torch.randn(ref.shape, generator=G(4)))
    ref.backward(dy)
    blk = blk.cuda().eval()
    xg = nhwc(x.detach()).requires_grad_(True)
    out = blk.forward_nhwc(xg)
    assert rel(nchw(out), ref.detach()) < 3e-3, rel(nchw(out), ref.detach())
    out.backward(gnhwc(dy))
    assert rel(gnchw(xg.grad), x.grad) < 2e-2, rel(gnchw(xg.grad), x.grad)
    for name, p in blk.named_parameters():
        want = P["b." + name].grad
        assert cos(p.grad, want) > 0.999 and rel(p.grad, want) < 4e-2, (name, rel(p.grad, want), cos(p.grad, want))
    for name, b in blk.named_buffers():           # eval mode: running statistics untouched
        assert torch.equal(b.cpu(), sd["b." + name]), name


def test_encoder_input_image_gradient_and_exporting_head_backward():
    from arbitrarystyletransfer_b200 import mobilenet as MB
    g = G(21)
    N, H, W = 2, 10, 14
    # stem: gradient with respect to the image (reflection padding folded back)
    img = torch.rand(N, 3, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(16, 3, 3, 3, generator=g) * 0.4).requires_grad_(True)
    y = F.hardswish(F.conv2d(F.pad(img, (1, 1, 1, 1), mode="reflect"), w))
    dy = bf16r(torch.randn(y.shape, generator=g))
    y.backward(dy)
    ig, wg = img.detach().cuda().requires_grad_(True), w.detach().cuda().requires_grad_(True)
    yg = MB._StemFn.apply(ig, wg)
    yg.backward(gnhwc(dy))
    assert rel(ig.grad, img.grad) < 5e-3, rel(ig.grad, img.grad)
    assert rel(wg.grad, w.grad) < 1e-2
    # through the public module: Encoder(x) with x.requires_grad (eval mode, frozen statistics)
    xin = torch.rand(2, 3, 32, 32, generator=g)
    Q = A.calibrate_running_stats(A.clone_state(A.activate_gates(A.make_ae_state(2))), xin)   # non-degenerate state
    torch.manual_seed(2)
    enc = MB.Encoder().cuda().eval()
    enc.load_state_dict({k[len("encoder."):]: v for k, v in Q.items() if k.startswith("encoder.")}, strict=True)
    gout = torch.randn(2, 128, 4, 4, generator=g)
    xg_ = xin.clone().cuda().requires_grad_(True)
    (enc(xg_, auto_enc=True) * gout.cuda()).sum().backward()
    xr = xin.clone().requires_grad_(True)
    (A.encoder_forward(Q, xr, auto_enc=True, training=False) * gout).sum().backward()
    assert xr.grad.abs().sum() > 0
    assert cos(xg_.grad, xr.grad) > 0.995 and rel(xg_.grad, xr.grad) < 0.1, (cos(xg_.grad, xr.grad), rel(xg_.grad, xr.grad))
    # exporting head: Hardtanh(0,1) after the conv (models.py:304, 315-316), with its backward
    x = f16r(torch.randn(N, 16, H, W, generator=g)).requires_grad_(True)
    hw = (torch.randn(3, 16, 3, 3, generator=g) * 0.2).requires_grad_(True)
    hb = (torch.rand(3, generator=g) * 0.6 + 0.2).requires_grad_(True)
    out = F.hardtanh(F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), hw, hb), 0.0, 1.0)
    dY = torch.randn(out.shape, generator=g)
    out.backward(dY)
    assert 0.05 < (out.detach() == 0).float().mean() + (out.detach() == 1).float().mean() < 0.95   # both regimes present
    xg = nhwc(x.detach()).requires_grad_(True)
    hwg, hbg = hw.detach().cuda().requires_grad_(True), hb.detach().cuda().requires_grad_(True)
    og = MB._HeadFn.apply(xg, hwg, hbg, True)
    assert rel(og, out.detach()) < 1e-4
    og.backward(dY.cuda())
    assert rel(hwg.grad, hw.grad) < 1e-4 and rel(hbg.grad, hb.grad) < 1e-4
    assert rel(gnchw(xg.grad), x.grad) < 5e-3
    dec = MB.Decoder(exporting=True).cuda()
    z = torch.rand(1, 128, 4, 4, generator=g).cuda()
    dec(z).sum().backward()                                  # no longer raises
    assert dec._img_out.weight.grad is not None
