"""Drop-in for the MobileNet-style path of the reference: ``Encoder`` (models.py:140-184),
``DecoderBlock`` / ``Decoder`` (models.py:242-320), ``AutoEncoder`` (models.py:322-338) and their
building blocks ``DepthWiseConv`` / ``SELayer`` / ``conv_3x3_bn`` (mobilenetv2.py:38-43, 63-81,
95-181).  Module trees, attribute names and ModuleList indices mirror the reference so that its
state dicts (``ae.pth`` keys such as ``encoder.mob_net.1._layers.3.weight``) load unchanged.

Storage formats.  Forward activations (and the GEMM weights they meet) are IEEE **fp16**; gradients are **bf16**
(include/ast_b200.h, K4).  autograd casts a gradient to the dtype of the tensor it belongs to, so a tensor pair
(fp16 activation, bf16 gradient) cannot be declared to torch as such: every internal NHWC tensor is therefore a
``torch.bfloat16`` tensor used as an opaque 16-bit container -- forward tensors hold fp16 BIT PATTERNS (``act_bits`` /
``act_float`` convert), gradient tensors are genuine bf16 (so autograd's own accumulation of gradients is correct).
No torch arithmetic ever touches a forward tensor; the public boundary is NCHW fp32.

Two execution paths, both entirely through libast_b200 (no torch arithmetic on activations):
  * inference (eval mode, no grad): BatchNorm running statistics folded into the convolutions, SE
    excitation folded into per-sample pointwise weights, SE squeeze fused into the depthwise stencil;
  * training (train_autoencoder.py:111-148): BatchNorm with batch statistics and running-stat updates,
    every intermediate kept in NHWC bf16, and hand-written backward kernels (pointwise data / weight
    gradients on tcgen05, depthwise data / weight gradients, BatchNorm / Hardswish / SE backward), wired
    into autograd one block at a time so the reference's optimiser / clip code runs unchanged.  Eval-mode
    BatchNorm with gradients (running statistics held constant), the gradient with respect to the input image and
    the backward of the exporting (Hardtanh) head are covered as well.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L

# ---- topology: conf.py:71-113 (values are inputs to the architecture, kept verbatim) ------------
EXPAND_RATIO = 3                                   # conf.py:71
enc_conv_shapes = [                                # conf.py:75-91  (c_in, c_out, stride, kernel, t)
    (3, 16, 1, 3, 1), (16, 16, 1, 3, 6), (16, 24, 2, 3, 6), (24, 24, 1, 3, 6), (24, 40, 2, 5, 6),
    (40, 40, 1, 5, 4), (40, 40, 1, 5, 4), (40, 80, 2, 3, 4), (80, 80, 1, 3, 4), (80, 80, 1, 3, 4),
    (80, 96, 1, 5, 4), (96, 96, 1, 5, 3), (96, 128, 1, 3, 3), (128, 128, 1, 3, 3), (128, 128, 1, 3, 3)]
decoder_conv_shapes = [                            # conf.py:93-109
    (128, 128, 1, 3, 3), (128, 128, 1, 3, 3), (128, 96, 1, 3, 3), (96, 96, 1, 5, 3), (96, 80, 1, 5, 4),
    (80, 80, 1, 3, 4), (80, 80, 1, 3, 4), (80, 40, 1, 3, 4), (40, 40, 1, 5, 4), (40, 40, 1, 5, 4),
    (40, 24, 1, 5, 6), (24, 24, 1, 3, 6), (24, 16, 1, 3, 6), (16, 16, 1, 3, 6), (16, 3, 1)]
enc_out_layers = [12, 14]                          # conf.py:112
enc_out_channels = 128                             # conf.py:113


def _make_divisible(v, divisor, min_value=None):
    """mobilenetv2.py:16-35."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# ---- thin kernel wrappers (NHWC bf16 tensors, "row-strided": stride(3) == 1, rows ld apart) -------
def _st(t):
    return L.stream_ptr(t.device)


def _rows(t):
    """(tensor, ld) with t addressable as a [N*H*W][ld] row matrix (a channel slice of a wider contiguous
    NHWC buffer qualifies); anything else is made contiguous."""
    N, H, W, Cc = t.shape
    ld = t.stride(2)
    ok = (t.stride(3) == 1 and ld % 8 == 0 and ld >= Cc and t.stride(1) == W * ld
          and t.stride(0) == H * W * ld and t.data_ptr() % 16 == 0)
    if not ok:
        t = t.contiguous()
        ld = Cc
    return t, ld


BITS16 = torch.bfloat16      # torch-level dtype of every internal NHWC tensor (opaque container, see module docstring)


def _empty(N, H, W, Cc, dev):
    return torch.empty(N, H, W, Cc, device=dev, dtype=BITS16)


_ACT_F16 = [True]


def set_activation_format(fmt: str) -> None:
    """Process-wide storage format of the forward activations: ``"fp16"`` (default: 11-bit significand, values below
    6e-8 flush to zero, saturation at 65504) or ``"bf16"`` (8-bit significand, fp32's exponent range).  fp16 is what
    brings the whole network within ~1 % of the fp32 reference; bf16 is for states whose activations leave fp16's
    range -- e.g. training from the reference's fresh initialisation, whose closed SE gates put decoder activations
    around 1e-20 (oracle/restate_ae.py::activate_gates).  Switch only while no internal tensor is alive (derived weight
    caches are keyed on the format)."""
    if fmt not in ("fp16", "bf16"):
        raise L.AstError("activation format must be 'fp16' or 'bf16'")
    L.check(L.load().ast_set_act_format(L.DT_F16 if fmt == "fp16" else L.DT_BF16), "ast_set_act_format")
    _ACT_F16[0] = fmt == "fp16"


def activation_format() -> str:
    return "fp16" if _ACT_F16[0] else "bf16"


def act_is_f16() -> bool:
    return _ACT_F16[0]


def act_bits(x: torch.Tensor) -> torch.Tensor:
    """fp32 values -> the 16-bit container holding them in the activation format (what forward kernels read)."""
    return x.to(torch.float16).view(BITS16) if _ACT_F16[0] else x.to(torch.bfloat16)


def act_float(t: torch.Tensor) -> torch.Tensor:
    """A forward (activation) tensor of this module -> fp32 values."""
    return t.view(torch.float16).float() if _ACT_F16[0] else t.float()


def pw_conv(x, w_bf16, bias, act, out_channels, residual=None, per_sample=False, want_raw=False, res_up2=False,
            f16=False):
    """x: (N,H,W,Cin) row-strided -> (N,H,W,Cout).  ``f16``: x / weights / residual / outputs are fp16 (forward
    activations), else bf16 (gradients, attention rows).  ``want_raw`` (training): returns
    (raw pre-activation, Hardswish(raw)).  ``res_up2``: residual is the half-resolution tensor, read
    through a nearest x2 upsample."""
    lib = L.load()
    x, ldx = _rows(x)
    N, H, W, Cin = x.shape
    out = _empty(N, H, W, out_channels, x.device)
    out_act = _empty(N, H, W, out_channels, x.device) if want_raw else None
    ld_res = 0
    if residual is not None:
        residual, ld_res = _rows(residual)
    L.check(lib.ast_pw_conv(x.data_ptr(), ldx, w_bf16.data_ptr(), int(per_sample), L.ptr(bias), int(act),
                            L.ptr(residual), ld_res, out.data_ptr(), out_channels, N, H * W, Cin, out_channels,
                            L.ptr(out_act), out_channels, W if res_up2 else 0, L.DT_F16 if f16 else L.DT_BF16,
                            _st(x)), "ast_pw_conv")
    return (out, out_act) if want_raw else out


def dw_conv(x, w_kkc, bias, k, stride, up2=False, act=1, want_pool=True):
    """act: 0 none, 1 store Hardswish, 2 store raw + pool Hardswish (training)."""
    lib = L.load()
    N, H, W, Cc = x.shape
    assert x.is_contiguous()
    Hin, Win = (2 * H, 2 * W) if up2 else (H, W)
    pad = (k - 1) // 2
    Ho, Wo = (Hin + 2 * pad - k) // stride + 1, (Win + 2 * pad - k) // stride + 1
    out = _empty(N, Ho, Wo, Cc, x.device)
    pool = torch.empty(N, Cc, device=x.device, dtype=torch.float32) if want_pool else None
    L.check(lib.ast_dw_conv(x.data_ptr(), w_kkc.data_ptr(), L.ptr(bias), out.data_ptr(), L.ptr(pool), N, Cc,
                            H, W, k, stride, int(up2), int(act), _st(x)), "ast_dw_conv")
    return out, pool


def se_fc(pool, inv_hw, w1, b1, w2, b2, save=False):
    lib = L.load()
    N, Cc = pool.shape
    S = w1.shape[0]
    scale = torch.empty(N, Cc, device=pool.device, dtype=torch.float32)
    hid = torch.empty(N, S, device=pool.device, dtype=torch.float32) if save else None
    pre = torch.empty(N, Cc, device=pool.device, dtype=torch.float32) if save else None
    L.check(lib.ast_se_fc(pool.data_ptr(), float(inv_hw), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                          b2.data_ptr(), scale.data_ptr(), L.ptr(hid), L.ptr(pre), N, Cc, S, _st(pool)),
            "ast_se_fc")
    return scale, hid, pre


def affine_act(x, sc, sh, act, se=None, res=None, want_out=True, want_pool=False):
    lib = L.load()
    x, ldx = _rows(x)
    N, H, W, Cc = x.shape
    out = _empty(N, H, W, Cc, x.device) if want_out else None
    pool = torch.empty(N, Cc, device=x.device, dtype=torch.float32) if want_pool else None
    ld_res = 0
    if res is not None:
        res, ld_res = _rows(res)
    L.check(lib.ast_affine_act(x.data_ptr(), ldx, L.ptr(sc), L.ptr(sh), int(act), L.ptr(se), L.ptr(res), ld_res,
                               L.ptr(out), Cc, L.ptr(pool), N, Cc, H * W, _st(x)), "ast_affine_act")
    return out, pool


def prep_weight(w, rows, cols, mode):
    """fp32 parameter viewed as [rows][cols] -> 0: bf16 same layout, 1: bf16 transposed, 2: fp32 transposed,
    3: fp16 same layout (forward GEMM weights; returned in the 16-bit container dtype)."""
    lib = L.load()
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    shape = (rows, cols) if mode in (0, 3) else (cols, rows)
    out = torch.empty(shape, device=w.device, dtype=torch.float32 if mode == 2 else BITS16)
    L.check(lib.ast_prep_weight(w.data_ptr(), out.data_ptr(), rows, cols, mode, _st(w)), "ast_prep_weight")
    return out


def prep_block_weights(w1, wd, w2, hid, inp, oup, kk):
    """Every prepared form of one DepthWiseConv block's weights in ONE launch (ast_prep_block_weights): returns
    (w1_f [hid][inp], w1_t [inp][hid], wd_t fp32 [k*k][hid], w2_f [oup][hid], w2_t [hid][oup]); ``w1`` may be None
    (expand_ratio == 1).  ``*_f`` are the forward GEMM operands (activation format), ``*_t`` the bf16 transposed
    operands of the data-gradient GEMMs."""
    lib = L.load()

    def f32(w):
        w = w.detach()
        return w if (w.dtype == torch.float32 and w.is_contiguous()) else w.float().contiguous()
    wd, w2 = f32(wd), f32(w2)
    dev = w2.device
    w1_f = w1_t = None
    if w1 is not None:
        w1 = f32(w1)
        w1_f = torch.empty((hid, inp), device=dev, dtype=BITS16)
        w1_t = torch.empty((inp, hid), device=dev, dtype=BITS16)
    wd_t = torch.empty((kk, hid), device=dev, dtype=torch.float32)
    w2_f = torch.empty((oup, hid), device=dev, dtype=BITS16)
    w2_t = torch.empty((hid, oup), device=dev, dtype=BITS16)
    L.check(lib.ast_prep_block_weights(L.ptr(w1), hid, inp, L.ptr(w1_f), L.ptr(w1_t), wd.data_ptr(), hid, kk,
                                       wd_t.data_ptr(), w2.data_ptr(), oup, hid, w2_f.data_ptr(), w2_t.data_ptr(),
                                       _st(w2)), "ast_prep_block_weights")
    return w1_f, w1_t, wd_t, w2_f, w2_t


def nchw_to_nhwc(x, f16=False):
    """NCHW fp32 -> NHWC 16-bit: fp16 bit patterns (``f16``, forward activations) or bf16 (gradients, attention)."""
    lib = L.load()
    x = x.float().contiguous()
    N, Cc, H, W = x.shape
    out = _empty(N, H, W, Cc, x.device)
    L.check(lib.ast_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), Cc, N, Cc, H * W, L.DT_F16 if f16 else L.DT_BF16,
                                 _st(x)), "ast_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, f16=False):
    lib = L.load()
    x, ld = _rows(x)
    N, H, W, Cc = x.shape
    out = torch.empty(N, Cc, H, W, device=x.device, dtype=torch.float32)
    L.check(lib.ast_nhwc_to_nchw(x.data_ptr(), ld, out.data_ptr(), N, Cc, H * W, L.DT_F16 if f16 else L.DT_BF16,
                                 _st(x)), "ast_nhwc_to_nchw")
    return out


def act_to_grad_format(x):
    """fp16 activation rows -> a bf16 copy: the weight-gradient GEMM multiplies it with a bf16 gradient and the
    tensor cores take one 16-bit format per instruction (ast_cvt_f16_to_bf16)."""
    lib = L.load()
    x, ld = _rows(x)
    N, H, W, Cc = x.shape
    out = _empty(N, H, W, Cc, x.device)
    L.check(lib.ast_cvt_f16_to_bf16(x.data_ptr(), ld, out.data_ptr(), Cc, N * H * W, Cc, _st(x)),
            "ast_cvt_f16_to_bf16")
    return out


class _ToNHWC(torch.autograd.Function):
    """NCHW fp32 -> NHWC fp16 activation; the (bf16) gradient goes back through the inverse conversion."""

    @staticmethod
    def forward(ctx, x):
        return nchw_to_nhwc(x, f16=act_is_f16())

    @staticmethod
    def backward(ctx, g):
        return nhwc_to_nchw(g, f16=False)


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return nhwc_to_nchw(x, f16=act_is_f16())

    @staticmethod
    def backward(ctx, g):
        return nchw_to_nhwc(g, f16=False)


def to_nhwc(x):
    """NCHW fp32 -> the module's forward (activation) layout."""
    return _ToNHWC.apply(x) if (torch.is_grad_enabled() and x.requires_grad) else nchw_to_nhwc(x, f16=act_is_f16())


def to_nchw(x):
    """A forward (activation) tensor -> NCHW fp32."""
    return _ToNCHW.apply(x) if (torch.is_grad_enabled() and x.requires_grad) else nhwc_to_nchw(x, f16=act_is_f16())


def _fold_bn(conv_w, bn):
    """eval-mode BatchNorm2d folded into the preceding bias-free conv: (w', b')."""
    w = conv_w.detach().float()
    if bn is None:
        return w, None
    inv = (bn.running_var.detach().float() + bn.eps).rsqrt() * bn.weight.detach().float()
    return w * inv.view(-1, 1, 1, 1), bn.bias.detach().float() - bn.running_mean.detach().float() * inv


# ---- training: BatchNorm with batch statistics, block forward / backward ----------------------------
def _bn_train_forward(a, bn):
    """Batch statistics of the NHWC tensor ``a`` for nn.BatchNorm2d ``bn`` (mobilenetv2.py:108 etc.): returns
    stat = float[4][C] (mean, invstd, scale, shift) and updates the running statistics in place."""
    lib = L.load()
    a, ld = _rows(a)
    N, H, W, Cc = a.shape
    if bn.momentum is None:
        raise L.AstError("BatchNorm2d(momentum=None) is not supported")
    sums = torch.empty(2, Cc, device=a.device, dtype=torch.float64)
    stat = torch.empty(4, Cc, device=a.device, dtype=torch.float32)
    L.check(lib.ast_bn_stats(a.data_ptr(), ld, sums.data_ptr(), N, Cc, H * W, _st(a)), "ast_bn_stats")
    track = bn.training and bn.track_running_stats
    L.check(lib.ast_bn_finalize(sums.data_ptr(), float(N * H * W), bn.weight.data_ptr(), bn.bias.data_ptr(),
                                bn.running_mean.data_ptr() if track else None,
                                bn.running_var.data_ptr() if track else None, float(bn.momentum), float(bn.eps),
                                stat.data_ptr(), Cc,
                                bn.num_batches_tracked.data_ptr() if track and bn.num_batches_tracked is not None else None,
                                _st(a)), "ast_bn_finalize")   # the counter's += 1 rides in the same launch
    return stat


def _bn_eval_stat(bn, dev):
    """stat = float[4][C] (mean, invstd, scale, shift) of nn.BatchNorm2d in EVAL mode: the running statistics take the
    place of the batch statistics and nothing is updated (ast_bn_finalize on sums that encode them with count 1)."""
    lib = L.load()
    rm, rv = bn.running_mean.detach().double(), bn.running_var.detach().double()
    sums = torch.stack((rm, rv + rm * rm)).contiguous()
    Cc = rm.numel()
    stat = torch.empty(4, Cc, device=dev, dtype=torch.float32)
    L.check(lib.ast_bn_finalize(sums.data_ptr(), 1.0, bn.weight.data_ptr(), bn.bias.data_ptr(), None, None, 0.0,
                                float(bn.eps), stat.data_ptr(), Cc, None, L.stream_ptr(dev)), "ast_bn_finalize")
    return stat


def _bn_forward(a, bn, frozen):
    return _bn_eval_stat(bn, a.device) if frozen else _bn_train_forward(a, bn)


def _bn_backward(dy, a, stat, frozen=False):
    """Generic BatchNorm backward on NHWC tensors: returns (da, dgamma, dbeta).  ``frozen`` (eval mode): the statistics
    are constants, so da = dy * scale -- the two batch-coupling coefficients are zeroed -- while dgamma / dbeta are the
    same sums."""
    lib = L.load()
    dy, ld_dy = _rows(dy)
    a, ld_a = _rows(a)
    N, H, W, Cc = a.shape
    dev = a.device
    sums = torch.empty(2, Cc, device=dev, dtype=torch.float64)
    coef = torch.empty(2, Cc, device=dev, dtype=torch.float32)
    dgamma = torch.empty(Cc, device=dev, dtype=torch.float32)
    dbeta = torch.empty(Cc, device=dev, dtype=torch.float32)
    da = _empty(N, H, W, Cc, dev)
    st = _st(a)
    L.check(lib.ast_bn_bwd_reduce(dy.data_ptr(), ld_dy, a.data_ptr(), ld_a, stat.data_ptr(), sums.data_ptr(),
                                  N, Cc, H * W, st), "ast_bn_bwd_reduce")
    L.check(lib.ast_bn_bwd_finalize(sums.data_ptr(), float(N * H * W), dgamma.data_ptr(), dbeta.data_ptr(),
                                    coef.data_ptr(), Cc, st), "ast_bn_bwd_finalize")
    if frozen:
        coef.zero_()
    L.check(lib.ast_bn_bwd_apply(dy.data_ptr(), ld_dy, a.data_ptr(), ld_a, stat.data_ptr(), coef.data_ptr(),
                                 da.data_ptr(), N, Cc, H * W, st), "ast_bn_bwd_apply")
    return da, dgamma, dbeta


def _pw_wgrad(a, b, out, si, sj, a_is_act=False, b_is_act=False):
    """out[i*si + j*sj] += sum_p a[p][i] * b[p][j]  (a, b NHWC row-strided; out fp32, zero-filled).  The GEMM runs
    bf16 x bf16: an operand that is a forward activation (fp16) is converted first."""
    lib = L.load()
    if a_is_act and act_is_f16():
        a = act_to_grad_format(a)
    if b_is_act and act_is_f16():
        b = act_to_grad_format(b)
    a, lda = _rows(a)
    b, ldb = _rows(b)
    P = a.shape[0] * a.shape[1] * a.shape[2]
    L.check(lib.ast_pw_wgrad(a.data_ptr(), lda, a.shape[3], b.data_ptr(), ldb, b.shape[3], P, out.data_ptr(),
                             si, sj, _st(a)), "ast_pw_wgrad")


class _BlockFn(torch.autograd.Function):
    """One DepthWiseConv block (mobilenetv2.py:95-165) in training mode on NHWC bf16 tensors.
    ``params`` follow DepthWiseConv._param_list()."""

    @staticmethod
    def forward(ctx, x, mod, up2, *params):
        lib = L.load()
        P = mod._unpack(params)
        norm, expand = mod.use_norm, mod.expand
        bns = mod._bns()
        x = _rows(x)[0] if expand else x.contiguous()
        N, H, W, _ = x.shape
        dev = x.device
        hid, k, stride = mod.hidden, mod.k, mod.stride
        frozen = norm and not mod.training       # eval-mode BatchNorm with gradients: running statistics, no update
        ctx.frozen = frozen
        a1 = stat1 = stat2 = stat3 = a3 = None
        # all five prepared weight forms of the block (forward + data-gradient operands) in one launch
        w1b, w1t, wd, w2b, w2t = prep_block_weights(P["w1"] if expand else None, P["wd"], P["w2"], hid, mod.inp, mod.oup,
                                                    k * k)
        if expand:
            if up2:
                raise L.AstError("the upsampled input is only supported for expand_ratio == 1 blocks")
            if norm:
                a1 = pw_conv(x, w1b, None, 0, hid, f16=act_is_f16())
                stat1 = _bn_forward(a1, bns[0], frozen)
                dw_in, _ = affine_act(a1, stat1[2], stat1[3], 1)
            else:
                a1, dw_in = pw_conv(x, w1b, None, 1, hid, want_raw=True, f16=act_is_f16())
        else:
            dw_in = x
        a2, pool = dw_conv(dw_in, wd, None, k, stride, up2=up2, act=0 if norm else 2, want_pool=not norm)
        Ho, Wo = a2.shape[1], a2.shape[2]
        if norm:
            stat2 = _bn_forward(a2, bns[1], frozen)
            _, pool = affine_act(a2, stat2[2], stat2[3], 1, want_out=False, want_pool=True)
        inv_hw = 1.0 / (Ho * Wo)
        s, sehid, sepre = se_fc(pool, inv_hw, P["se_w1"], P["se_b1"], P["se_w2"], P["se_b2"], save=True)
        u, _ = affine_act(a2, stat2[2] if norm else None, stat2[3] if norm else None, 1, se=s)
        res = x if mod.identity else None
        if norm:
            a3 = pw_conv(u, w2b, None, 0, mod.oup, f16=act_is_f16())
            stat3 = _bn_forward(a3, bns[2], frozen)
            out, _ = affine_act(a3, stat3[2], stat3[3], 0, res=res)
        else:
            out = pw_conv(u, w2b, None, 0, mod.oup, residual=res, res_up2=bool(up2 and mod.identity), f16=act_is_f16())
        ctx.mod, ctx.up2, ctx.geom = mod, up2, (N, H, W, Ho, Wo)
        ctx.n_params = len(params)
        ctx.save_for_backward(x, a1, dw_in if expand else None, a2, u, a3, s, sehid, sepre, pool, stat1, stat2,
                              stat3, wd, w1t, w2t, *params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = L.load()
        mod, up2 = ctx.mod, ctx.up2
        N, H, W, Ho, Wo = ctx.geom
        saved = ctx.saved_tensors
        x, a1, h1, a2, u, a3, s, sehid, sepre, pool, stat1, stat2, stat3, wd, w1t, w2t = saved[:16]
        P = mod._unpack(saved[16:])
        norm, expand = mod.use_norm, mod.expand
        hid, k, stride, inp, oup = mod.hidden, mod.k, mod.stride, mod.inp, mod.oup
        dev = x.device
        st = _st(x)
        grads = {}
        d_out, _ = _rows(d_out.to(torch.bfloat16))
        # ---- pw-linear (+ its norm) ----
        if norm:
            d_a3, grads["g3"], grads["b3"] = _bn_backward(d_out, a3, stat3, ctx.frozen)
        else:
            d_a3 = d_out
        d_u = pw_conv(d_a3, w2t, None, 0, hid)
        # one zero fill for the block's three atomically accumulated weight gradients (segments 256-byte aligned)
        names = ("w2", "wd") + (("w1",) if expand else ())
        offs, tot = [], 0
        for nm in names:
            offs.append(tot)
            tot += (P[nm].numel() + 63) // 64 * 64
        flat = torch.zeros(tot, device=dev, dtype=torch.float32)
        zg = {nm: flat[o:o + P[nm].numel()].view(P[nm].shape) for nm, o in zip(names, offs)}
        dW2 = zg["w2"]
        _pw_wgrad(u, d_a3, dW2, 1, hid, a_is_act=True)                   # dW2[j][i] += sum u[p][i] d_a3[p][j]
        grads["w2"] = dW2
        # ---- SE, Hardswish and the depthwise norm ----
        HWo = Ho * Wo
        T = torch.empty(N, 5, hid, device=dev, dtype=torch.float32)
        L.check(lib.ast_dw_bwd_reduce(d_u.data_ptr(), a2.data_ptr(), L.ptr(stat2), T.data_ptr(), N, hid, HWo, st),
                "ast_dw_bwd_reduce")
        S = P["se_w1"].shape[0]
        dpre = torch.empty(N, hid, device=dev, dtype=torch.float32)
        dhid = torch.empty(N, S, device=dev, dtype=torch.float32)
        g = torch.empty(N, hid, device=dev, dtype=torch.float32)
        for nm in ("se_w1", "se_b1", "se_w2", "se_b2"):
            grads[nm] = torch.empty_like(P[nm], dtype=torch.float32)
        L.check(lib.ast_se_bwd(T.data_ptr(), 5 * hid, sepre.data_ptr(), sehid.data_ptr(), pool.data_ptr(),
                               1.0 / HWo, P["se_w1"].data_ptr(), P["se_w2"].data_ptr(), dpre.data_ptr(),
                               dhid.data_ptr(), g.data_ptr(), grads["se_w1"].data_ptr(), grads["se_b1"].data_ptr(),
                               grads["se_w2"].data_ptr(), grads["se_b2"].data_ptr(), N, hid, S, st), "ast_se_bwd")
        coef2 = None
        if norm:
            coef2 = torch.empty(2, hid, device=dev, dtype=torch.float32)
            grads["g2"] = torch.empty(hid, device=dev, dtype=torch.float32)
            grads["b2"] = torch.empty(hid, device=dev, dtype=torch.float32)
            L.check(lib.ast_se_bn_combine(T.data_ptr(), s.data_ptr(), g.data_ptr(), grads["g2"].data_ptr(),
                                          grads["b2"].data_ptr(), coef2.data_ptr(), N, hid, float(N * HWo), st),
                    "ast_se_bn_combine")
            if ctx.frozen:
                coef2.zero_()
        d_a2 = _empty(N, Ho, Wo, hid, dev)
        L.check(lib.ast_dw_bwd_apply(d_u.data_ptr(), a2.data_ptr(), s.data_ptr(), g.data_ptr(), L.ptr(stat2),
                                     L.ptr(coef2), d_a2.data_ptr(), N, hid, HWo, st), "ast_dw_bwd_apply")
        # ---- depthwise conv ----
        dw_in = h1 if expand else x
        dWd = zg["wd"]
        L.check(lib.ast_dw_conv_wgrad(d_a2.data_ptr(), dw_in.data_ptr(), dWd.data_ptr(), N, hid, H, W, k, stride,
                                      int(up2), st), "ast_dw_conv_wgrad")
        grads["wd"] = dWd
        d_in = _empty(N, H, W, hid, dev)
        dres = d_out if (not expand and mod.identity) else None
        if dres is not None and not dres.is_contiguous():
            dres = dres.contiguous()
        L.check(lib.ast_dw_conv_dgrad(d_a2.data_ptr(), wd.data_ptr(), L.ptr(a1) if expand else None,
                                      L.ptr(stat1) if expand else None, L.ptr(dres), d_in.data_ptr(), N, hid, H, W,
                                      k, stride, int(up2), st), "ast_dw_conv_dgrad")
        # ---- pw expand (+ its norm) ----
        if expand:
            if norm:
                d_a1, grads["g1"], grads["b1"] = _bn_backward(d_in, a1, stat1, ctx.frozen)
            else:
                d_a1 = d_in
            dW1 = zg["w1"]
            _pw_wgrad(d_a1, x, dW1, inp, 1, b_is_act=True)               # dW1[i][j] += sum d_a1[p][i] x[p][j]
            grads["w1"] = dW1
            d_x = None
            if ctx.needs_input_grad[0]:
                d_x = pw_conv(d_a1, w1t, None, 0, inp,
                              residual=d_out if mod.identity else None)
        else:
            d_x = d_in if ctx.needs_input_grad[0] else None
        return (d_x, None, None) + tuple(grads[nm] for nm in mod._param_names())


class _StemFn(torch.autograd.Function):
    """conv_3x3_bn (mobilenetv2.py:38-43) from the NCHW fp32 image to NHWC bf16, with the weight gradient."""

    @staticmethod
    def forward(ctx, img, w):
        lib = L.load()
        img = img.float().contiguous()
        N, _, H, W = img.shape
        cout = w.shape[0]
        y = _empty(N, H, W, cout, img.device)
        z = _empty(N, H, W, cout, img.device)
        wf = w.detach().float().contiguous()
        L.check(lib.ast_stem_conv(img.data_ptr(), wf.data_ptr(), y.data_ptr(), z.data_ptr(), N, H, W, cout,
                                  _st(img)), "ast_stem_conv")
        ctx.save_for_backward(img, z, wf)
        ctx.wshape = tuple(w.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        img, z, wf = ctx.saved_tensors
        N, _, H, W = img.shape
        dy = dy.to(torch.bfloat16).contiguous()
        dw = None
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(ctx.wshape, device=img.device, dtype=torch.float32)
            L.check(lib.ast_stem_wgrad(dy.data_ptr(), z.data_ptr(), img.data_ptr(), dw.data_ptr(), N, H, W,
                                       ctx.wshape[0], _st(img)), "ast_stem_wgrad")
        dimg = None
        if ctx.needs_input_grad[0]:
            dimg = torch.empty_like(img)
            L.check(lib.ast_stem_dgrad(dy.data_ptr(), z.data_ptr(), wf.data_ptr(), dimg.data_ptr(), N, H, W,
                                       ctx.wshape[0], _st(img)), "ast_stem_dgrad")
        return dimg, dw


class _HeadFn(torch.autograd.Function):
    """Decoder._ref_out + _img_out (models.py:300-314): NHWC bf16 -> NCHW fp32 image, with all gradients."""

    @staticmethod
    def forward(ctx, x, w, b, clamp):
        lib = L.load()
        x = x if x.is_contiguous() else x.contiguous()
        N, H, W, Cc = x.shape
        co = w.shape[0]
        out = torch.empty(N, co, H, W, device=x.device, dtype=torch.float32)
        wf, bf = w.detach().float().contiguous(), b.detach().float().contiguous()
        L.check(lib.ast_head_conv(x.data_ptr(), wf.data_ptr(), bf.data_ptr(), out.data_ptr(), N, H, W, Cc, co,
                                  int(clamp), _st(x)), "ast_head_conv")
        ctx.save_for_backward(x, wf, out if clamp else None)
        return out

    @staticmethod
    def backward(ctx, dY):
        lib = L.load()
        x, wf, y_clamped = ctx.saved_tensors
        N, H, W, Cc = x.shape
        co = wf.shape[0]
        dY = dY.float().contiguous()
        st = _st(x)
        if y_clamped is not None:     # Hardtanh(0,1) of the exporting head (models.py:304, 315-316)
            masked = torch.empty_like(dY)
            L.check(lib.ast_hardtanh01_bwd(dY.data_ptr(), y_clamped.data_ptr(), masked.data_ptr(), dY.numel(), st),
                    "ast_hardtanh01_bwd")
            dY = masked
        dw = torch.zeros_like(wf)
        db = torch.zeros(co, device=x.device, dtype=torch.float32)
        L.check(lib.ast_head_wgrad(dY.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), N, H, W, Cc, co, st),
                "ast_head_wgrad")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _empty(N, H, W, Cc, x.device)
            L.check(lib.ast_head_dgrad(dY.data_ptr(), wf.data_ptr(), dx.data_ptr(), N, H, W, Cc, co, st),
                    "ast_head_dgrad")
        return dx, dw, db, None


def _wants_grad(module, *tensors):
    return torch.is_grad_enabled() and (any(t.requires_grad for t in tensors)
                                        or any(p.requires_grad for p in module.parameters()))


# ---- modules ---------------------------------------------------------------------------------------
class SELayer(nn.Module):
    """mobilenetv2.py:63-81 (parameter container; the math runs inside DepthWiseConv.forward_nhwc)."""

    def __init__(self, channel, reduction=4):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        sq = _make_divisible(channel // reduction, 8)
        self.fc = nn.Sequential(nn.Linear(channel, sq), nn.ReLU(inplace=True), nn.Linear(sq, channel),
                                nn.Hardtanh(0.0, 1.0))


def conv_3x3_bn(inp, oup, stride):
    """mobilenetv2.py:38-43: reflect-padded 3x3 conv without bias + Hardswish (no BN despite the name)."""
    return nn.Sequential(nn.Conv2d(inp, oup, 3, stride, 1, bias=False, padding_mode="reflect"),
                         nn.Hardswish(True))


class DepthWiseConv(nn.Module):
    """mobilenetv2.py:95-181: [pw expand -> (BN) -> Hardswish ->] dw k x k reflect -> (BN) -> Hardswish ->
    SE -> pw linear -> (BN) [+ identity]."""

    def __init__(self, inp, oup, stride, expand_ratio, kernel_size=3, use_norm=False, padding=0,
                 use_identity=True, use_relu=False):
        super().__init__()
        hidden_dim = round(inp * expand_ratio)
        self.identity = stride == 1 and inp == oup and use_identity
        self.inp, self.oup, self.hidden, self.stride, self.k = inp, oup, hidden_dim, stride, kernel_size
        self.expand = expand_ratio != 1
        self.use_norm = use_norm
        layers = []

        def bn(c):
            if use_norm:
                layers.append(nn.BatchNorm2d(c, affine=True, track_running_stats=True))

        if not self.expand:
            layers.append(nn.ReflectionPad2d((1, 1, 1, 1)))
            layers.append(nn.Conv2d(hidden_dim, hidden_dim, kernel_size, stride, 0, groups=hidden_dim, bias=False))
            bn(hidden_dim)
            layers.append(nn.Hardswish(True))
            layers.append(SELayer(hidden_dim))
            layers.append(nn.Conv2d(hidden_dim, oup, 1, 1, 0, bias=False))
            bn(oup)
        else:
            layers.append(nn.Conv2d(inp, hidden_dim, 1, 1, 0, bias=False))
            bn(hidden_dim)
            layers.append(nn.Hardswish(True))
            layers.append(nn.Conv2d(hidden_dim, hidden_dim, kernel_size, stride, (kernel_size - 1) // 2,
                                    groups=hidden_dim, bias=False, padding_mode="reflect"))
            bn(hidden_dim)
            layers.append(nn.Hardswish(True))
            layers.append(SELayer(hidden_dim))
            layers.append(nn.Conv2d(hidden_dim, oup, 1, 1, 0, bias=False))
            bn(oup)
        self._layers = nn.ModuleList(layers)
        self._initialize_weights()
        self._prep = None

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    # -- derived kernel parameters, rebuilt when any parameter / buffer changes ---------------------
    def _prepared(self):
        ver = (act_is_f16(),) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._prep is not None and self._prep[0] == ver:
            return self._prep[1]
        mods = list(self._layers)
        convs = [m for m in mods if isinstance(m, nn.Conv2d)]
        se = next(m for m in mods if isinstance(m, SELayer))

        def bn_after(conv):
            i = mods.index(conv)
            return mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None

        d = {}
        if self.expand:
            pw1, dw, pw2 = convs
            w, b = _fold_bn(pw1.weight, bn_after(pw1))
            d["w1"] = act_bits(w.view(self.hidden, self.inp)).contiguous()               # fp16 bit patterns
            d["b1"] = b.contiguous() if b is not None else None
        else:
            dw, pw2 = convs
        w, b = _fold_bn(dw.weight, bn_after(dw))
        d["wd"] = w.view(self.hidden, self.k * self.k).t().contiguous()          # fp32 [k*k][C]
        d["bd"] = b.contiguous() if b is not None else None
        w, b = _fold_bn(pw2.weight, bn_after(pw2))
        d["w2"] = w.view(self.oup, self.hidden).contiguous()                      # fp32, SE-scaled per call
        d["b2"] = b.contiguous() if b is not None else None
        d["se"] = [t.detach().float().contiguous() for t in
                   (se.fc[0].weight, se.fc[0].bias, se.fc[2].weight, se.fc[2].bias)]
        self._prep = (ver, d)
        return d

    # -- parameter plumbing for the training Function ------------------------------------------------
    def _split(self):
        mods = list(self._layers)
        convs = [m for m in mods if isinstance(m, nn.Conv2d)]
        bns = [m for m in mods if isinstance(m, nn.BatchNorm2d)]
        se = next(m for m in mods if isinstance(m, SELayer))
        return convs, bns, se

    def _bns(self):
        return self._split()[1]

    def _param_names(self):
        n = self.use_norm
        names = (["w1"] + (["g1", "b1"] if n else [])) if self.expand else []
        names += ["wd"] + (["g2", "b2"] if n else [])
        names += ["se_w1", "se_b1", "se_w2", "se_b2", "w2"] + (["g3", "b3"] if n else [])
        return names

    def _param_list(self):
        convs, bns, se = self._split()
        out = []
        bi = 0
        if self.expand:
            out.append(convs[0].weight)
            if self.use_norm:
                out += [bns[bi].weight, bns[bi].bias]; bi += 1
        out.append(convs[-2].weight)
        if self.use_norm:
            out += [bns[bi].weight, bns[bi].bias]; bi += 1
        out += [se.fc[0].weight, se.fc[0].bias, se.fc[2].weight, se.fc[2].bias, convs[-1].weight]
        if self.use_norm:
            out += [bns[bi].weight, bns[bi].bias]
        return out

    def _unpack(self, params):
        return dict(zip(self._param_names(), params))

    def forward_nhwc(self, x, up2=False):
        """x: (N,H,W,inp) bf16 NHWC -> (N,Ho,Wo,oup) bf16 NHWC.  ``up2``: the block consumes the nearest
        x2 upsample of x (DecoderBlock._upsample_3 followed by _upsample_2, models.py:263-267)."""
        lib = L.load()
        grad = _wants_grad(self, x)
        if grad or (self.training and self.use_norm):
            return _BlockFn.apply(x, self, up2, *self._param_list())
        d = self._prepared()
        N = x.shape[0]
        h = pw_conv(x, d["w1"], d["b1"], act=True, out_channels=self.hidden, f16=act_is_f16()) if self.expand else x
        if not self.expand and not h.is_contiguous():
            h = h.contiguous()
        y, pool = dw_conv(h, d["wd"], d["bd"], self.k, self.stride, up2=up2, act=1, want_pool=True)
        Ho, Wo = y.shape[1], y.shape[2]
        w1, b1, w2, b2 = d["se"]
        scale, _, _ = se_fc(pool, 1.0 / (Ho * Wo), w1, b1, w2, b2)
        w2s = torch.empty(N, self.oup, self.hidden, device=x.device, dtype=BITS16)
        L.check(lib.ast_scale_weights(d["w2"].data_ptr(), scale.data_ptr(), w2s.data_ptr(), N, self.oup,
                                      self.hidden, _st(x)), "ast_scale_weights")
        res = x if self.identity else None
        return pw_conv(y, w2s, d["b2"], act=False, out_channels=self.oup, residual=res, per_sample=True,
                       res_up2=bool(up2 and self.identity), f16=act_is_f16())

    def forward(self, x):
        """NCHW fp32 in / out like the reference module."""
        L.require_cuda(x)
        return to_nchw(self.forward_nhwc(to_nhwc(x)))


class Encoder(nn.Module):
    """models.py:140-184.  ``forward(x, out_layers=[], auto_enc=False)``."""

    def __init__(self, exporting=False, use_inst_norm=False):
        super().__init__()
        blocks = [conv_3x3_bn(enc_conv_shapes[0][0], enc_conv_shapes[0][1], enc_conv_shapes[0][2])]
        for in_ch, out_ch, stride, kernel_size, expand_ratio in enc_conv_shapes[1:-1]:
            blocks.append(DepthWiseConv(in_ch, out_ch, stride, expand_ratio, use_norm=True,
                                        kernel_size=kernel_size))
        # models.py:154 re-uses the loop variables of the last iteration
        blocks.append(DepthWiseConv(in_ch, out_ch, stride, EXPAND_RATIO, use_norm=True))
        self.mob_net = nn.ModuleList(blocks)

    def forward_nhwc(self, x, out_layers=(), auto_enc=False):
        lib = L.load()
        x = x.float().contiguous()
        N, _, H, W = x.shape
        stem = self.mob_net[0][0]
        cout = stem.out_channels
        if _wants_grad(self.mob_net[0], x):
            y = _StemFn.apply(x, stem.weight)
        else:
            y = _empty(N, H, W, cout, x.device)
            L.check(lib.ast_stem_conv(x.data_ptr(), stem.weight.detach().float().contiguous().data_ptr(),
                                      y.data_ptr(), None, N, H, W, cout, _st(x)), "ast_stem_conv")
        outs = [y] if 0 in out_layers else []
        for i, layer in enumerate(self.mob_net):
            if i == 0:
                continue
            y = layer.forward_nhwc(y)
            if i in out_layers:
                outs.append(y)
        return y if auto_enc else outs

    def forward(self, x, out_layers=[], auto_enc=False):
        L.require_cuda(x)
        r = self.forward_nhwc(x, tuple(out_layers), auto_enc)
        return to_nchw(r) if auto_enc else [to_nchw(t) for t in r]


class DecoderBlock(nn.Module):
    """models.py:242-272."""

    def __init__(self, in_channels, out_channels, stride, kernel_size=3, upsample=False, expand_ratio=6):
        super().__init__()
        self._ref_pad = nn.ReflectionPad2d((1, 1, 1, 1))   # constructed but unused, as in the reference
        self._conv = DepthWiseConv(in_channels, out_channels, stride, expand_ratio, use_norm=False,
                                   kernel_size=kernel_size)
        self._should_upsample = upsample
        if self._should_upsample:
            self._ref_out = nn.ReflectionPad2d((1, 1, 1, 1))
            self._upsample_2 = DepthWiseConv(out_channels, out_channels, 1, 1, use_norm=False)
            self._upsample_3 = nn.Upsample(scale_factor=2, mode='nearest')

    def forward_nhwc(self, x):
        x = self._conv.forward_nhwc(x)
        if self._should_upsample:
            x = self._upsample_2.forward_nhwc(x, up2=True)   # nearest x2 folded into the stencil's reads
        return x

    def forward(self, x):
        L.require_cuda(x)
        return to_nchw(self.forward_nhwc(to_nhwc(x)))


class Decoder(nn.Module):
    """models.py:274-320."""

    def __init__(self, exporting=False):
        super().__init__()
        self.exporting = exporting
        blocks = []
        for i, conv_shape in enumerate(decoder_conv_shapes[:-1]):
            should_upsample = (conv_shape[0] != conv_shape[1] and i + 6 < len(decoder_conv_shapes))
            blocks.append(DecoderBlock(conv_shape[0], conv_shape[1], conv_shape[2], upsample=should_upsample,
                                       expand_ratio=conv_shape[4], kernel_size=conv_shape[3]))
        self._decoder_blocks = nn.ModuleList(blocks)
        self._ref_out = nn.ReflectionPad2d((1, 1, 1, 1))
        self._img_out = nn.Conv2d(decoder_conv_shapes[-1][0], decoder_conv_shapes[-1][1], kernel_size=(3, 3))
        self.last_act = nn.Hardtanh(0.0, 1.0)

    def forward_nhwc(self, x):
        lib = L.load()
        for block in self._decoder_blocks:
            x = block.forward_nhwc(x)
        if _wants_grad(self._img_out, x):
            return _HeadFn.apply(x, self._img_out.weight, self._img_out.bias, bool(self.exporting))
        N, H, W, Cc = x.shape
        co = self._img_out.out_channels
        x = x if x.is_contiguous() else x.contiguous()
        out = torch.empty(N, co, H, W, device=x.device, dtype=torch.float32)
        L.check(lib.ast_head_conv(x.data_ptr(), self._img_out.weight.detach().float().contiguous().data_ptr(),
                                  self._img_out.bias.detach().float().contiguous().data_ptr(), out.data_ptr(),
                                  N, H, W, Cc, co, int(self.exporting), _st(x)), "ast_head_conv")
        return out

    def forward(self, x):
        L.require_cuda(x)
        return self.forward_nhwc(to_nhwc(x))


class AutoEncoder(nn.Module):
    """models.py:322-338: encoder taps [12, 14] -> cat -> ada_out -> decoder."""

    def __init__(self):
        super().__init__()
        self.encoder = Encoder(use_inst_norm=True)
        self.ada_out = DepthWiseConv(enc_out_channels * 2, enc_out_channels, 1, EXPAND_RATIO, use_norm=False,
                                     use_identity=False)
        self.decoder = Decoder()

    def forward(self, x):
        L.require_cuda(x)
        e = self.encoder.forward_nhwc(x, tuple(enc_out_layers))
        z = self.ada_out.forward_nhwc(torch.cat((e[0], e[1]), dim=3))     # models.py:332
        return self.decoder.forward_nhwc(z)
