"""Drop-in for the reference's attention-normalisation layer and the network built on it:
  AdaAttN   models.py:70-115
  AST       models.py:393-582   (what train.py actually trains)
Same constructor / ``forward`` / ``encode`` signatures, return conventions and state-dict keys
(``ada_att_1.W_q.weight`` ..., ``_enc.mob_net...``, ``_dec._decoder_blocks...``, ``ada_out._layers...``).

The reference's ``AST`` does not run as shipped: models.py:459 is a syntax error and
``__init__`` leaves ``ada_att_2`` / ``ada_out`` commented out (models.py:407, 410) although ``encode`` / ``forward``
/ train.py:142-144, 295-298 use them.  This class restores those two attributes with exactly the commented
constructor calls and unpacks ``encode(..., return_maps=True)`` the way models.py:568-569 returns it.

Device work: ``libast_b200.so`` only (K6: batched tcgen05 GEMM + streaming passes; K1 for the instance norms; K4 for
the 1x1 convolutions, encoder, ``ada_out`` and decoder).  CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import functional as Fn
from . import mobilenet as MB

__all__ = ["AdaAttN", "AST", "bgemm"]

enc_out_layers = MB.enc_out_layers
enc_out_channels = MB.enc_out_channels
EXPAND_RATIO = MB.EXPAND_RATIO


def _st(t):
    return L.stream_ptr(t.device)


def _pad8(n):
    return (n + 7) // 8 * 8


def bgemm(a, a_mn, b, b_mn, M, N, K, out_dtype=torch.float32, ld_a=None, ld_b=None):
    """d[b] (M x N) = op(a[b]) op(b[b])^T on tcgen05 (ast_bgemm).  ``a``: (B, rows, ld) bf16 with rows = K if a_mn
    else M; ``b`` likewise with N.  Returns (B, M, ld_d) with ld_d = N rounded up to 8 (callers slice / pass ld)."""
    lib = L.load()
    B = a.shape[0]
    ld_a = a.stride(1) if ld_a is None else ld_a
    ld_b = b.stride(1) if ld_b is None else ld_b
    ld_d = _pad8(N)
    d = torch.empty(B, M, ld_d, device=a.device, dtype=out_dtype)
    L.check(lib.ast_bgemm(a.data_ptr(), int(a_mn), ld_a, a.stride(0), b.data_ptr(), int(b_mn), ld_b, b.stride(0),
                          d.data_ptr(), int(out_dtype == torch.bfloat16), ld_d, d.stride(0), M, N, K, B, _st(a)),
            "ast_bgemm")
    return d


def _split3_rows(x, C, pattern):
    """fp32 (..., ld) row matrix -> bf16 (rows, 3C): [hi | lo | hi] (pattern 0) or [hi | hi | lo] (pattern 1)."""
    lib = L.load()
    x2 = x.reshape(-1, x.shape[-1])
    rows = x2.shape[0]
    out = torch.empty(rows, 3 * C, device=x.device, dtype=torch.bfloat16)
    L.check(lib.ast_split3_rows(x2.data_ptr(), x2.stride(0), out.data_ptr(), rows, C, pattern, _st(x)),
            "ast_split3_rows")
    return out


def _split3_nchw(x, pattern):
    lib = L.load()
    N, C, H, W = x.shape
    out = torch.empty(N, H * W, 3 * C, device=x.device, dtype=torch.bfloat16)
    L.check(lib.ast_split3_nchw(x.data_ptr(), out.data_ptr(), N, C, H * W, pattern, _st(x)), "ast_split3_nchw")
    return out


def _hi(x3, C):
    """The leading bf16 term of a split matrix as an NHWC-style 4-D row view (rows, ld = 3C)."""
    N, HW, _ = x3.shape
    return x3[:, :, :C].unsqueeze(1)          # (N, 1, HW, C), strides (HW*3C, HW*3C, 3C, 1)


def _layer_forward(cn_f, sn_f, style, wq, wk, wv):
    """The whole layer on NCHW fp32 inputs (cn_f / sn_f already instance-normalised).  Returns the NHWC bf16 output
    and everything the backward pass needs."""
    lib = L.load()
    N, C, H, W = cn_f.shape
    Hs, Ws = style.shape[2], style.shape[3]
    HW, HWs = H * W, Hs * Ws
    dev = cn_f.device
    st = _st(cn_f)
    cn3 = _split3_nchw(cn_f, 0)                                                  # (N, HW, 3C)
    sn3 = _split3_nchw(sn_f, 0)
    xs = MB.nchw_to_nhwc(style)                                                  # (N, Hs, Ws, C) bf16
    wq3 = _split3_rows(wq.detach().float().reshape(C, C), C, 1)                  # (C, 3C)
    wk3 = _split3_rows(wk.detach().float().reshape(C, C), C, 1)
    # models.py:87-88: q = W_q(IN(content)), k = W_k(IN(style)) as split GEMMs, fp32 results
    q_f = bgemm(cn3.view(1, N * HW, 3 * C), 0, wq3.view(1, C, 3 * C), 0, N * HW, C, 3 * C)
    k_f = bgemm(sn3.view(1, N * HWs, 3 * C), 0, wk3.view(1, C, 3 * C), 0, N * HWs, C, 3 * C)
    q3 = _split3_rows(q_f[0], C, 0).view(N, HW, 3 * C)
    k3 = _split3_rows(k_f[0], C, 1).view(N, HWs, 3 * C)
    v = MB.pw_conv(xs, MB.prep_weight(wv, C, C, 0), None, False, C)              # :89 (N, Hs, Ws, C) bf16
    S = bgemm(q3, 0, k3, 0, HW, HWs, 3 * C)                                      # :97  (fp32)
    ldp = _pad8(HWs)
    P = torch.empty(N, HW, ldp, device=dev, dtype=torch.bfloat16)
    lsum = torch.empty(N * HW, device=dev, dtype=torch.float32)
    L.check(lib.ast_attn_softmax(S.data_ptr(), S.stride(1), P.data_ptr(), ldp, lsum.data_ptr(), N * HW, HWs, st),
            "ast_attn_softmax")                                                  # :99
    vv3 = torch.empty(N, HWs, 3 * C, device=dev, dtype=torch.bfloat16)
    L.check(lib.ast_attn_vv3(v.data_ptr(), C, vv3.data_ptr(), N * HWs, C, st), "ast_attn_vv3")
    mm = bgemm(P, 0, vv3, 1, HW, 3 * C, HWs)                                     # :101, :103  (fp32, ld = 3C)
    out = torch.empty(N, H, W, C, device=dev, dtype=torch.bfloat16)
    L.check(lib.ast_attn_out_fwd(mm.data_ptr(), lsum.data_ptr(), cn3.data_ptr(), 3 * C, out.data_ptr(), N * HW, C,
                                 st), "ast_attn_out_fwd")                        # :103, :115
    return out, (cn3, sn3, xs, q3, k3, v, P, lsum, mm)


class _AdaAttNFn(torch.autograd.Function):
    """models.py:86-115 after the instance norms: (IN(content), IN(style), style, W_q, W_k, W_v) -> NCHW fp32."""

    @staticmethod
    def forward(ctx, cn_f, sn_f, style, wq, wk, wv):
        cn_f, sn_f, style = cn_f.contiguous(), sn_f.contiguous(), style.contiguous()
        out, saved = _layer_forward(cn_f, sn_f, style, wq, wk, wv)
        ctx.save_for_backward(wq, wk, wv, *saved)
        ctx.dims = (cn_f.shape, style.shape)
        return MB.nhwc_to_nchw(out)

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        wq, wk, wv, cn3, sn3, xs, q3, k3, v, P, lsum, mm = ctx.saved_tensors
        (N, C, H, W), (_, _, Hs, Ws) = ctx.dims
        HW, HWs = H * W, Hs * Ws
        dev = g.device
        st = _st(g)
        dout = MB.nchw_to_nhwc(g)
        dmm = torch.empty(N, HW, 5 * C, device=dev, dtype=torch.bfloat16)
        dcn = torch.empty(N, H, W, C, device=dev, dtype=torch.bfloat16)
        L.check(lib.ast_attn_out_bwd(mm.data_ptr(), lsum.data_ptr(), cn3.data_ptr(), 3 * C, dout.data_ptr(),
                                     dmm.data_ptr(), dcn.data_ptr(), N * HW, C, st), "ast_attn_out_bwd")
        # dA = dMean v^T + dM2 (v^2)^T with dMean, dM2 as two-term splits against the exact [v | v | hi | hi | lo]
        vv5 = torch.empty(N, HWs, 5 * C, device=dev, dtype=torch.bfloat16)
        L.check(lib.ast_attn_vv5(v.data_ptr(), C, vv5.data_ptr(), N * HWs, C, st), "ast_attn_vv5")
        dA = bgemm(dmm, 0, vv5, 0, HW, HWs, 5 * C)
        # d[v | v^2] = P^T [dMean_hi | dMean_lo | dM2_hi | dM2_lo]: contraction over the query rows of both operands
        # -> MN-major operands
        dvv = bgemm(P, 1, dmm, 1, HWs, 4 * C, HW)
        ldp = P.stride(1)
        dS = torch.empty(N, HW, ldp, device=dev, dtype=torch.bfloat16)
        L.check(lib.ast_attn_softmax_bwd(P.data_ptr(), ldp, lsum.data_ptr(), dA.data_ptr(), dA.stride(1),
                                         dS.data_ptr(), ldp, N * HW, HWs, st), "ast_attn_softmax_bwd")
        # gradients of the logits' operands in plain bf16 against the leading terms of q / k (columns [0, C) of the
        # splits, row stride 3C)
        dq = bgemm(dS, 0, k3, 1, HW, C, HWs, out_dtype=torch.bfloat16)           # dS k       (N, HW, C)
        dk = bgemm(dS, 1, q3, 1, HWs, C, HW, out_dtype=torch.bfloat16)           # dS^T q     (N, HWs, C)
        dv = torch.empty(N, Hs, Ws, C, device=dev, dtype=torch.bfloat16)
        L.check(lib.ast_attn_dv(dvv.data_ptr(), v.data_ptr(), C, dv.data_ptr(), N * HWs, C, st), "ast_attn_dv")
        dq4, dk4 = dq.view(N, H, W, C), dk.view(N, Hs, Ws, C)
        need = ctx.needs_input_grad
        g_cn = g_sn = g_xs = g_wq = g_wk = g_wv = None
        if need[0]:
            g_cn = MB.nhwc_to_nchw(MB.pw_conv(dq4, MB.prep_weight(wq, C, C, 1), None, False, C, residual=dcn))
        if need[1]:
            g_sn = MB.nhwc_to_nchw(MB.pw_conv(dk4, MB.prep_weight(wk, C, C, 1), None, False, C))
        if need[2]:
            g_xs = MB.nhwc_to_nchw(MB.pw_conv(dv, MB.prep_weight(wv, C, C, 1), None, False, C))
        if need[3]:
            g_wq = torch.zeros_like(wq, dtype=torch.float32)
            MB._pw_wgrad(dq4, _hi(cn3, C), g_wq, C, 1)
        if need[4]:
            g_wk = torch.zeros_like(wk, dtype=torch.float32)
            MB._pw_wgrad(dk4, _hi(sn3, C), g_wk, C, 1)
        if need[5]:
            g_wv = torch.zeros_like(wv, dtype=torch.float32)
            MB._pw_wgrad(dv, xs, g_wv, C, 1)
        return g_cn, g_sn, g_xs, g_wq, g_wk, g_wv


class AdaAttN(nn.Module):
    """models.py:70-115.  ``forward(content_map, style_map) -> std * IN(content) + mean`` with the attention-weighted
    per-position mean / standard deviation of ``W_v(style)`` under ``softmax(W_q(IN(content)) W_k(IN(style))^T)``.
    NCHW fp32 in / out like the reference; C must be a multiple of 8."""

    def __init__(self, inp_size):
        super().__init__()
        self.W_q = nn.Conv2d(inp_size, inp_size, 1, 1, 0, bias=False)
        self.W_k = nn.Conv2d(inp_size, inp_size, 1, 1, 0, bias=False)
        self.W_v = nn.Conv2d(inp_size, inp_size, 1, 1, 0, bias=False)
        self.att_act = nn.Softmax(dim=-1)          # parameter-free members kept for attribute parity (models.py:77-81)
        self.std_act = nn.ReLU(True)
        self.inst_norm_1 = nn.InstanceNorm2d(inp_size)
        self.inst_norm_2 = nn.InstanceNorm2d(inp_size)
        self.inst_norm = nn.InstanceNorm2d(inp_size)
        self.inp_size = inp_size

    def forward(self, content_map, style_map):
        L.require_cuda(content_map, style_map)
        if content_map.dim() != 4 or style_map.dim() != 4:
            raise L.AstError("AdaAttN expects 4-D (N, C, H, W) feature maps")
        if content_map.shape[1] != self.inp_size or style_map.shape[1] != self.inp_size:
            raise L.AstError(f"AdaAttN({self.inp_size}) got {content_map.shape[1]} / {style_map.shape[1]} channels")
        if style_map.shape[0] != content_map.shape[0]:
            raise L.AstError("content and style batch sizes differ")
        if self.inp_size % 8 != 0:
            raise L.AstError("AdaAttN needs a channel count that is a multiple of 8")
        content_map, style_map = content_map.float(), style_map.float()
        eps = self.inst_norm.eps
        cn_f = Fn.instance_norm(content_map, eps)          # inst_norm_1 == inst_norm on the same input (:87, :115)
        sn_f = Fn.instance_norm(style_map, eps)            # :88
        args = (cn_f, sn_f, style_map, self.W_q.weight, self.W_k.weight, self.W_v.weight)
        if torch.is_grad_enabled() and any(a.requires_grad for a in args):
            return _AdaAttNFn.apply(*args)
        out, _ = _layer_forward(cn_f.contiguous(), sn_f.contiguous(), style_map.contiguous(), *args[3:])
        return MB.nhwc_to_nchw(out)


class _Axpby(torch.autograd.Function):
    """a*x + b*y on fp32 tensors (models.py:471)."""

    @staticmethod
    def forward(ctx, x, y, a, b):
        ctx.ab = (a, b)
        return _axpby(x, y, a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.ab
        g = g.contiguous()
        return (_axpby(g, None, a, 0.0) if ctx.needs_input_grad[0] else None,
                _axpby(g, None, b, 0.0) if ctx.needs_input_grad[1] else None, None, None)


def _axpby(x, y, a, b):
    lib = L.load()
    x = x.float().contiguous()
    y = y.float().contiguous() if y is not None else None
    out = torch.empty_like(x)
    L.check(lib.ast_axpby(x.data_ptr(), L.ptr(y), float(a), float(b), out.data_ptr(), x.numel(), _st(x)), "ast_axpby")
    return out


def blend(t, content_map, alpha):
    """models.py:471: ``alpha * t + (1 - alpha) * content_map``."""
    if alpha == 1.0:
        return t
    if torch.is_grad_enabled() and (t.requires_grad or content_map.requires_grad):
        return _Axpby.apply(t, content_map, float(alpha), float(1.0 - alpha))
    return _axpby(t, content_map, alpha, 1.0 - alpha)


class AST(nn.Module):
    """models.py:393-582.  ``forward(content_img, style_img, alpha=1.0)`` returns ``(t_cs, t_return, org_out)``, or
    ``t_cs`` when ``exporting``; ``encode``, ``save``, ``load`` as the reference."""

    def __init__(self, style_layers=[4, 7, 10, 12, 16], content_layers=[4, 7, 10, 12, 16], exporting=False):
        super().__init__()
        self._style_layers = style_layers
        self._content_layers = content_layers
        self._exporting = exporting
        self._enc = MB.Encoder(self._exporting)
        self._dec = MB.Decoder(self._exporting)
        self.upsample_att_1 = nn.Upsample(scale_factor=2, mode='nearest')
        self.ada_att_1 = AdaAttN(enc_out_channels)
        self.ada_att_2 = AdaAttN(enc_out_channels)                                    # models.py:407 (restored)
        self.ada_out = MB.DepthWiseConv(enc_out_channels * 2, enc_out_channels, 1, EXPAND_RATIO, use_norm=False,
                                        use_identity=False)                           # models.py:410 (restored)

    def forward(self, content_img, style_img, alpha=1.0):
        L.require_cuda(content_img, style_img)
        if not self._exporting:
            stylized_map_1, stylized_map_2, t = self._encode_nhwc(content_img, style_img, True)   # models.py:459
            t_return = stylized_map_1
            cm = self._enc.forward_nhwc(content_img, tuple(enc_out_layers))           # :467
            content_map = self.ada_out.forward_nhwc(torch.cat((cm[0], cm[1]), dim=3))  # :468-469
            if alpha != 1.0:
                t = MB.to_nhwc(blend(MB.to_nchw(t), MB.to_nchw(content_map), alpha))  # :471
            org_out = self._dec.forward_nhwc(content_map)                             # :476
        else:
            t = self._encode_nhwc(content_img, style_img, False)[2]                   # :479
        t_cs = self._dec.forward_nhwc(t)                                              # :506
        if self._exporting:
            return t_cs
        return t_cs, t_return, org_out

    def _encode_nhwc(self, content_img, style_img, detach):
        """models.py:535-566 -> (stylized_map_1, stylized_map_2 [NCHW fp32], ada_out code [NHWC bf16])."""
        if detach:
            self._enc.eval()                                                          # :539
            with torch.no_grad():                                                     # the taps are detached (:543-545)
                cm = self._enc(content_img, out_layers=enc_out_layers)
                sm = self._enc(style_img, out_layers=enc_out_layers)
            self._enc.train()                                                         # :547 (unconditional)
        else:
            cm = self._enc(content_img, out_layers=enc_out_layers)
            sm = self._enc(style_img, out_layers=enc_out_layers)
        s1 = self.ada_att_1(cm[0], sm[0])                                             # :554
        s2 = self.ada_att_2(cm[1], sm[1])                                             # :555
        z = self.ada_out.forward_nhwc(MB.to_nhwc(torch.cat((s1, s2), dim=1)))         # :565-566
        return s1, s2, z

    def encode(self, content_img, style_img, detach=False, return_maps=False):
        """models.py:535-572; NCHW fp32 results like the reference's."""
        L.require_cuda(content_img, style_img)
        s1, s2, z = self._encode_nhwc(content_img, style_img, detach)
        z = MB.to_nchw(z)
        return (s1, s2, z) if return_maps else z

    def save(self):
        torch.save(self._dec.state_dict(), "models/dec.pth")                          # models.py:577-578

    def load(self):
        self._dec.load_state_dict(torch.load("models/dec.pth"))                       # models.py:580-582
