"""Drop-in for the MobileNet-style path of the reference: ``Encoder`` (models.py:140-184),
``DecoderBlock`` / ``Decoder`` (models.py:242-320), ``AutoEncoder`` (models.py:322-338) and their
building blocks ``DepthWiseConv`` / ``SELayer`` / ``conv_3x3_bn`` (mobilenetv2.py:38-43, 63-81,
95-181).  Module trees, attribute names and ModuleList indices mirror the reference so that its
state dicts (``ae.pth`` keys such as ``encoder.mob_net.1._layers.3.weight``) load unchanged.

Status: FORWARD, EVAL MODE (BatchNorm running statistics, folded into the convolutions).  Training
mode (batch statistics + backward) is not implemented yet and raises instead of computing silently
wrong results.  All device work goes through libast_b200 (pointwise convs on tcgen05, depthwise
stencil with the SE squeeze fused, SE excitation folded into per-sample pointwise weights).
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


# ---- thin kernel wrappers (NHWC bf16 tensors) ------------------------------------------------------
def _st(t):
    return L.stream_ptr(t.device)


def pw_conv(x, w_bf16, bias, act, out_channels, residual=None, per_sample=False):
    """x: (N,H,W,Cin) bf16 (last-dim stride 1, row stride x.stride(2)) -> (N,H,W,Cout) bf16."""
    lib = L.load()
    N, H, W, Cin = x.shape
    out = torch.empty(N, H, W, out_channels, device=x.device, dtype=torch.bfloat16)
    L.check(lib.ast_pw_conv(x.data_ptr(), x.stride(2), w_bf16.data_ptr(), int(per_sample), L.ptr(bias),
                            int(act), L.ptr(residual), residual.stride(2) if residual is not None else 0,
                            out.data_ptr(), out_channels, N, H * W, Cin, out_channels, _st(x)), "ast_pw_conv")
    return out


def dw_conv(x, w_kkc, bias, k, stride, up2=False, act=True, want_pool=True):
    lib = L.load()
    N, H, W, Cc = x.shape
    Hin, Win = (2 * H, 2 * W) if up2 else (H, W)
    pad = (k - 1) // 2
    Ho, Wo = (Hin + 2 * pad - k) // stride + 1, (Win + 2 * pad - k) // stride + 1
    out = torch.empty(N, Ho, Wo, Cc, device=x.device, dtype=torch.bfloat16)
    pool = torch.empty(N, Cc, device=x.device, dtype=torch.float32) if want_pool else None
    L.check(lib.ast_dw_conv(x.data_ptr(), w_kkc.data_ptr(), L.ptr(bias), out.data_ptr(), L.ptr(pool), N, Cc,
                            H, W, k, stride, int(up2), int(act), _st(x)), "ast_dw_conv")
    return out, pool


def nchw_to_nhwc(x):
    lib = L.load()
    x = x.float().contiguous()
    N, Cc, H, W = x.shape
    out = torch.empty(N, H, W, Cc, device=x.device, dtype=torch.bfloat16)
    L.check(lib.ast_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), N, Cc, H * W, _st(x)), "ast_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x):
    lib = L.load()
    N, H, W, Cc = x.shape
    out = torch.empty(N, Cc, H, W, device=x.device, dtype=torch.float32)
    L.check(lib.ast_nhwc_to_nchw(x.data_ptr(), x.stride(2), out.data_ptr(), N, Cc, H * W, _st(x)),
            "ast_nhwc_to_nchw")
    return out


def _fold_bn(conv_w, bn):
    """eval-mode BatchNorm2d folded into the preceding bias-free conv: (w', b')."""
    w = conv_w.detach().float()
    if bn is None:
        return w, None
    inv = (bn.running_var.detach().float() + bn.eps).rsqrt() * bn.weight.detach().float()
    return w * inv.view(-1, 1, 1, 1), bn.bias.detach().float() - bn.running_mean.detach().float() * inv


def _forward_only(module):
    if module.training and any(isinstance(m, nn.BatchNorm2d) for m in module.modules()):
        raise L.AstError("train-mode BatchNorm (batch statistics) is not implemented for the MobileNet-style "
                         "blocks yet: call .eval()")
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise L.AstError("the MobileNet-style blocks are forward-only so far (no backward kernels): "
                         "call them under torch.no_grad()")


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
        ver = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
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
            d["w1"] = w.view(self.hidden, self.inp).to(torch.bfloat16).contiguous()
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

    def forward_nhwc(self, x, up2=False):
        """x: (N,H,W,inp) bf16 NHWC -> (N,Ho,Wo,oup) bf16 NHWC.  ``up2``: the block consumes the nearest
        x2 upsample of x (DecoderBlock._upsample_3 followed by _upsample_2, models.py:263-267)."""
        lib = L.load()
        d = self._prepared()
        N = x.shape[0]
        h = pw_conv(x, d["w1"], d["b1"], act=True, out_channels=self.hidden) if self.expand else x
        y, pool = dw_conv(h, d["wd"], d["bd"], self.k, self.stride, up2=up2, act=True, want_pool=True)
        Ho, Wo = y.shape[1], y.shape[2]
        w1, b1, w2, b2 = d["se"]
        scale = torch.empty(N, self.hidden, device=x.device, dtype=torch.float32)
        L.check(lib.ast_se_fc(pool.data_ptr(), 1.0 / (Ho * Wo), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                              b2.data_ptr(), scale.data_ptr(), N, self.hidden, w1.shape[0], _st(x)), "ast_se_fc")
        w2s = torch.empty(N, self.oup, self.hidden, device=x.device, dtype=torch.bfloat16)
        L.check(lib.ast_scale_weights(d["w2"].data_ptr(), scale.data_ptr(), w2s.data_ptr(), N, self.oup,
                                      self.hidden, _st(x)), "ast_scale_weights")
        res = x if self.identity else None
        return pw_conv(y, w2s, d["b2"], act=False, out_channels=self.oup, residual=res, per_sample=True)

    def forward(self, x):
        """NCHW fp32 in / out like the reference module."""
        L.require_cuda(x)
        _forward_only(self)
        return nhwc_to_nchw(self.forward_nhwc(nchw_to_nhwc(x)))


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
        y = torch.empty(N, H, W, cout, device=x.device, dtype=torch.bfloat16)
        L.check(lib.ast_stem_conv(x.data_ptr(), stem.weight.detach().float().contiguous().data_ptr(),
                                  y.data_ptr(), N, H, W, cout, _st(x)), "ast_stem_conv")
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
        _forward_only(self)
        r = self.forward_nhwc(x, tuple(out_layers), auto_enc)
        return nhwc_to_nchw(r) if auto_enc else [nhwc_to_nchw(t) for t in r]


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
        _forward_only(self)
        return nhwc_to_nchw(self.forward_nhwc(nchw_to_nhwc(x)))


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
        N, H, W, Cc = x.shape
        co = self._img_out.out_channels
        out = torch.empty(N, co, H, W, device=x.device, dtype=torch.float32)
        L.check(lib.ast_head_conv(x.data_ptr(), self._img_out.weight.detach().float().contiguous().data_ptr(),
                                  self._img_out.bias.detach().float().contiguous().data_ptr(), out.data_ptr(),
                                  N, H, W, Cc, co, int(self.exporting), _st(x)), "ast_head_conv")
        return out

    def forward(self, x):
        L.require_cuda(x)
        _forward_only(self)
        return self.forward_nhwc(nchw_to_nhwc(x))


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
        _forward_only(self)
        e = self.encoder.forward_nhwc(x, tuple(enc_out_layers))
        z = self.ada_out.forward_nhwc(torch.cat((e[0], e[1]), dim=3))     # models.py:332
        return self.decoder.forward_nhwc(z)
