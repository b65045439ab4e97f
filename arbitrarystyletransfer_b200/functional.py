"""Thin functional layer over the C ABI: one Python function (or autograd.Function) per kernel
family.  Tensors in and out are the reference's own convention -- NCHW, fp32, contiguous, CUDA.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib as L


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _flags(t: torch.Tensor, canonical=False, biased=False) -> int:
    f = 0
    if canonical:
        f |= L.F_CANONICAL
    if biased:
        f |= L.F_BIASED
    if t.dtype == torch.bfloat16:
        f |= L.F_BF16
    elif t.dtype != torch.float32:
        raise L.AstError(f"only fp32 / bf16 tensors are supported, got {t.dtype}")
    return f


# ------------------------------------------------------------------------------------------
# K1: fused AdaIN                      reference: models.py:43-51 (+ :471 alpha blend)
# ------------------------------------------------------------------------------------------
def adain_forward(content: torch.Tensor, styles: Sequence[torch.Tensor],
                  weights: Sequence[float] | None = None, alpha: float = 1.0,
                  canonical: bool = False, eps: float = 0.0, out: torch.Tensor | None = None,
                  return_stats: bool = False):
    """out = alpha * ((c - mu_c)/sigma_c * A + B) + (1 - alpha) * c with
    (A, B) = sum_k w_k (mu_s, sigma_s)  [reference, swapped as models.py:44]  or (sigma_s, mu_s)
    [canonical].  One pass over HBM; statistics in fp32 (Welford / Chan)."""
    lib = L.load()
    if isinstance(styles, torch.Tensor):
        styles = [styles]
    K = len(styles)
    if K < 1:
        raise L.AstError("adain_forward needs at least one style map")
    if K > L.MAX_STYLES:
        raise L.AstError(f"at most {L.MAX_STYLES} style maps are supported, got {K}")
    L.require_cuda(content, *styles)
    if content.dim() != 4:
        raise L.AstError("content_map must be 4-D (N, C, H, W)")
    content = _c(content)
    styles = [_c(s) for s in styles]
    N, Cc, H, W = content.shape
    for s in styles:
        if s.dim() != 4 or s.shape[0] != N or s.shape[1] != Cc or s.dtype != content.dtype:
            raise L.AstError("style maps must be (N, C, Hs, Ws) with the content's N, C and dtype")
    if weights is None:
        weights = [1.0 / K] * K
    if len(weights) != K:
        raise L.AstError("one interpolation weight per style map is required")
    if out is None:
        out = torch.empty_like(content)
    stats = (torch.empty(N * Cc, 2 + 2 * K, device=content.device, dtype=torch.float32)
             if return_stats else None)
    sp = (C.c_void_p * K)(*[s.data_ptr() for s in styles])
    shw = (C.c_int64 * K)(*[s.shape[2] * s.shape[3] for s in styles])
    rc = lib.ast_adain_fwd(content.data_ptr(), sp, shw, L.float_array(weights), K, out.data_ptr(),
                           L.ptr(stats), N, Cc, H * W, float(alpha), float(eps),
                           _flags(content, canonical), L.stream_ptr(content.device))
    L.check(rc, "ast_adain_fwd")
    return (out, stats) if return_stats else out


class _AdaIN(torch.autograd.Function):
    """AdaIN.forward (models.py:43-51) + alpha blend (:471) for feature maps that require grad: the fused forward
    kernel, and a backward made of kernels only -- ``ast_adain_bwd`` (content gradient + the per-row statistics
    gradients of every style) and ``ast_channel_stats_bwd`` per style map."""

    @staticmethod
    def forward(ctx, alpha, canonical, weights, content, *styles):
        out, stats = adain_forward(content, list(styles), list(weights), alpha, canonical, return_stats=True)
        ctx.save_for_backward(_c(content), stats, *[_c(s) for s in styles])
        ctx.alpha, ctx.canonical, ctx.weights = float(alpha), bool(canonical), [float(w) for w in weights]
        return out

    @staticmethod
    def backward(ctx, gy):
        lib = L.load()
        content, stats, *styles = ctx.saved_tensors
        K = len(styles)
        N, Cc = content.shape[:2]
        rows, HW = N * Cc, content[0, 0].numel()
        gy = _c(gy.to(content.dtype))
        gc = torch.empty_like(content)
        need_styles = any(ctx.needs_input_grad[4 + k] for k in range(K))
        aux = torch.empty(K, 4, rows, device=content.device, dtype=torch.float32) if need_styles else None
        fl = _flags(content, ctx.canonical)
        L.check(lib.ast_adain_bwd(content.data_ptr(), gy.data_ptr(), stats.data_ptr(), L.float_array(ctx.weights), K,
                                  ctx.alpha, gc.data_ptr(), L.ptr(aux), rows, HW, fl, L.stream_ptr(content.device)),
                "ast_adain_bwd")
        gs = []
        for k, s in enumerate(styles):
            if not ctx.needs_input_grad[4 + k]:
                gs.append(None)
                continue
            g = torch.empty_like(s)
            a = aux[k]
            L.check(lib.ast_channel_stats_bwd(s.data_ptr(), a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(),
                                              a[3].data_ptr(), g.data_ptr(), rows, s[0, 0].numel(),
                                              _flags(s), L.stream_ptr(s.device)), "ast_channel_stats_bwd")
            gs.append(g)
        return (None, None, None, gc if ctx.needs_input_grad[3] else None, *gs)


def adain_autograd(content, styles, weights=None, alpha: float = 1.0, canonical: bool = False):
    """Differentiable fused AdaIN (see :class:`_AdaIN`)."""
    styles = [styles] if isinstance(styles, torch.Tensor) else list(styles)
    K = len(styles)
    weights = list(weights) if weights is not None else [1.0 / K] * K
    return _AdaIN.apply(alpha, canonical, tuple(weights), content, *styles)


# ------------------------------------------------------------------------------------------
# channel_stats                          reference: model_util.py:3-8, models.py:54-62
# ------------------------------------------------------------------------------------------
class _ChannelStats(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, biased):
        lib = L.load()
        L.require_cuda(x)
        x = _c(x)
        N, Cc = x.shape[:2]
        HW = x[0, 0].numel()
        mean = torch.empty(N, Cc, device=x.device, dtype=torch.float32)
        std = torch.empty(N, Cc, device=x.device, dtype=torch.float32)
        fl = _flags(x, biased=biased)
        L.check(lib.ast_channel_stats_fwd(x.data_ptr(), mean.data_ptr(), std.data_ptr(), N * Cc, HW,
                                          float(eps), fl, L.stream_ptr(x.device)),
                "ast_channel_stats_fwd")
        ctx.save_for_backward(x, mean, std)
        ctx.fl = fl
        return mean, std

    @staticmethod
    def backward(ctx, g_mean, g_std):
        lib = L.load()
        x, mean, std = ctx.saved_tensors
        N, Cc = x.shape[:2]
        HW = x[0, 0].numel()
        gx = torch.empty_like(x)
        gm = _c(g_mean.float()) if g_mean is not None else None
        gs = _c(g_std.float()) if g_std is not None else None
        L.check(lib.ast_channel_stats_bwd(x.data_ptr(), mean.data_ptr(), std.data_ptr(), L.ptr(gm),
                                          L.ptr(gs), gx.data_ptr(), N * Cc, HW, ctx.fl,
                                          L.stream_ptr(x.device)), "ast_channel_stats_bwd")
        return gx, None, None


def channel_stats_flat(x: torch.Tensor, eps: float = 0.0, biased: bool = False):
    """(mean, std) as (N, C) fp32 tensors; std = sqrt(var + eps), unbiased unless ``biased``."""
    if x.dim() < 3:
        raise L.AstError("channel statistics need a tensor of at least 3 dimensions")
    return _ChannelStats.apply(x, eps, biased)


# ------------------------------------------------------------------------------------------
# mean_variance_norm                                          reference: models.py:64-68
# ------------------------------------------------------------------------------------------
class _MVN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, biased=False):
        lib = L.load()
        L.require_cuda(x)
        x = _c(x)
        N, Cc = x.shape[:2]
        HW = x[0, 0].numel()
        y = torch.empty_like(x)
        stats = torch.empty(N * Cc, 2, device=x.device, dtype=torch.float32)
        fl = _flags(x, biased=biased)
        L.check(lib.ast_mvn_fwd(x.data_ptr(), y.data_ptr(), stats.data_ptr(), N * Cc, HW, float(eps),
                                fl, L.stream_ptr(x.device)), "ast_mvn_fwd")
        ctx.save_for_backward(x, stats)
        ctx.fl = fl
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.load()
        x, stats = ctx.saved_tensors
        N, Cc = x.shape[:2]
        HW = x[0, 0].numel()
        gy = _c(gy.to(x.dtype))
        gx = torch.empty_like(x)
        L.check(lib.ast_mvn_bwd(x.data_ptr(), gy.data_ptr(), stats.data_ptr(), gx.data_ptr(), N * Cc,
                                HW, ctx.fl, L.stream_ptr(x.device)), "ast_mvn_bwd")
        return gx, None, None


def mean_variance_norm(feat: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return _MVN.apply(feat, eps)


def instance_norm(feat: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm2d without affine / running statistics (models.py:78-80): the same fused kernel with the
    BIASED variance, (x - mean) / sqrt(var_biased + eps) per (n, c); differentiable."""
    return _MVN.apply(feat, eps, True)


# ------------------------------------------------------------------------------------------
# K3: Huber / Gram                                           reference: losses.py:105-139
# ------------------------------------------------------------------------------------------
class _Huber(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, tgt, scale):
        lib = L.load()
        L.require_cuda(inp, tgt)
        if inp.shape != tgt.shape:
            raise L.AstError(f"huber_loss shapes differ: {tuple(inp.shape)} vs {tuple(tgt.shape)}")
        inp = _c(inp.float())
        tgt = _c(tgt.float())
        n = inp.numel()
        loss = torch.empty((), device=inp.device, dtype=torch.float32)
        wsb = lib.ast_huber_ws_bytes(n)
        ws = torch.empty(wsb, device=inp.device, dtype=torch.uint8)
        L.check(lib.ast_huber_fwd(inp.data_ptr(), tgt.data_ptr(), loss.data_ptr(), n, float(scale),
                                  ws.data_ptr(), wsb, L.stream_ptr(inp.device)), "ast_huber_fwd")
        ctx.save_for_backward(inp, tgt)
        ctx.scale = float(scale)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        inp, tgt = ctx.saved_tensors
        g = _c(g.float().reshape(1))
        gi = torch.empty_like(inp) if ctx.needs_input_grad[0] else None
        gt = None
        if gi is not None:
            L.check(lib.ast_huber_bwd(inp.data_ptr(), tgt.data_ptr(), g.data_ptr(), gi.data_ptr(),
                                      inp.numel(), ctx.scale, L.stream_ptr(inp.device)),
                    "ast_huber_bwd")
        if ctx.needs_input_grad[1]:
            gt = torch.empty_like(tgt)
            L.check(lib.ast_huber_bwd(tgt.data_ptr(), inp.data_ptr(), g.data_ptr(), gt.data_ptr(),
                                      inp.numel(), ctx.scale, L.stream_ptr(inp.device)),
                    "ast_huber_bwd")
        return gi, gt, None


def huber_loss(inp: torch.Tensor, tgt: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """scale * F.huber_loss(inp, tgt) (delta 1, mean) as a 0-dim tensor with autograd history."""
    return _Huber.apply(inp, tgt, scale)


class _TV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img):
        lib = L.load()
        L.require_cuda(img)
        if img.dim() != 4:
            raise L.AstError("tv_loss expects a 4-D (N, C, H, W) tensor")
        img = _c(img.float())
        N, Cc, H, W = img.shape
        loss = torch.empty((), device=img.device, dtype=torch.float32)
        wsb = lib.ast_huber_ws_bytes(img.numel())
        ws = torch.empty(wsb, device=img.device, dtype=torch.uint8)
        L.check(lib.ast_tv_fwd(img.data_ptr(), loss.data_ptr(), N * Cc, H, W, ws.data_ptr(), wsb,
                               L.stream_ptr(img.device)), "ast_tv_fwd")
        ctx.save_for_backward(img)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        (img,) = ctx.saved_tensors
        N, Cc, H, W = img.shape
        g = _c(g.float().reshape(1))
        gi = torch.empty_like(img)
        L.check(lib.ast_tv_bwd(img.data_ptr(), g.data_ptr(), gi.data_ptr(), N * Cc, H, W, L.stream_ptr(img.device)),
                "ast_tv_bwd")
        return gi


def tv_loss(img: torch.Tensor) -> torch.Tensor:
    """sum of squared horizontal and vertical neighbour differences (losses.py:90-103), 0-dim, differentiable."""
    return _TV.apply(img)


class _HistLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        lib = L.load()
        L.require_cuda(x, y)
        if x.dim() != 4 or y.dim() != 4 or x.shape[0] != y.shape[0]:
            raise L.AstError("compute_hist_loss expects two 4-D (B, C, H, W) tensors with the same batch size")
        x, y = _c(x.float()), _c(y.float())
        B = x.shape[0]
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        gd = torch.empty(B, 256, device=x.device, dtype=torch.float32)
        wsb = lib.ast_hist_ws_bytes(B)
        ws = torch.empty((wsb + 7) // 8, device=x.device, dtype=torch.int64)
        ctx.norms = (float(x.shape[1] * x.shape[2]), float(y.shape[1] * y.shape[2]))     # losses.py:54
        L.check(lib.ast_hist_loss_fwd(x.data_ptr(), y.data_ptr(), B, x[0].numel(), y[0].numel(), ctx.norms[0],
                                      ctx.norms[1], loss.data_ptr(), gd.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                      L.stream_ptr(x.device)), "ast_hist_loss_fwd")
        ctx.save_for_backward(x, y, gd)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        x, y, gd = ctx.saved_tensors
        g = _c(g.float().reshape(1))
        out = []
        for i, (t, sign) in enumerate(((x, 1.0), (y, -1.0))):
            if not ctx.needs_input_grad[i]:
                out.append(None)
                continue
            gt = torch.empty_like(t)
            L.check(lib.ast_hist_loss_bwd(t.data_ptr(), gd.data_ptr(), g.data_ptr(), sign, ctx.norms[i],
                                          gt.data_ptr(), t.shape[0], t[0].numel(), L.stream_ptr(t.device)),
                    "ast_hist_loss_bwd")
            out.append(gt)
        return tuple(out)


def hist_loss(t_cs: torch.Tensor, style_map: torch.Tensor) -> torch.Tensor:
    """compute_hist_loss (losses.py:82-87): mean over the batch of the squared EMD between the soft 256-bin histograms
    of the two tensors; 0-dim, differentiable in both arguments."""
    return _HistLoss.apply(t_cs, style_map)


# "tf32": Gram forward on the tensor cores in TF32 and Gram backward on the tensor cores with bf16 operands when
# the shape allows (default; the gradient is rounded to bf16 anyway when it enters the VGG backward pass);
# "fp32": CUDA-core kernels with 1e-6 agreement.  The reference computes torch.bmm in fp32 (losses.py:109); on Ampere+ GPUs
# PyTorch itself may run that bmm in TF32 when torch.backends.cuda.matmul.allow_tf32 is set.
GRAM_PRECISION = "tf32"


class _Gram(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = L.load()
        L.require_cuda(x)
        if x.dim() != 4:
            raise L.AstError("gram_matrix expects (B, C, H, W)")
        x = _c(x.float())
        B, Cc, H, W = x.shape
        g = torch.empty(B, Cc, Cc, device=x.device, dtype=torch.float32)
        if GRAM_PRECISION == "tf32" and (H * W) % 4 == 0 and Cc % 16 == 0 and H * W >= 64:
            L.check(lib.ast_gram_fwd_tf32(x.data_ptr(), g.data_ptr(), B, Cc, H * W, L.stream_ptr(x.device)),
                    "ast_gram_fwd_tf32")
        else:
            L.check(lib.ast_gram_fwd(x.data_ptr(), g.data_ptr(), B, Cc, H * W, L.stream_ptr(x.device)),
                    "ast_gram_fwd")
        ctx.save_for_backward(x)
        return g

    @staticmethod
    def backward(ctx, gg):
        lib = L.load()
        (x,) = ctx.saved_tensors
        B, Cc, H, W = x.shape
        gg = _c(gg.float())
        gx = torch.empty_like(x)
        HW = H * W
        if GRAM_PRECISION == "tf32" and Cc % 8 == 0 and HW % 8 == 0 and Cc >= 64 and HW >= 256:
            # tensor-core backward: bf16 copies of X and of (gg + gg^T)/(C*HW) as MN-major operands, fp32 accumulate
            wsb = lib.ast_gram_bwd_tc_ws_bytes(B, Cc, HW)
            ws = torch.empty(wsb, device=x.device, dtype=torch.uint8)
            L.check(lib.ast_gram_bwd_tc(x.data_ptr(), gg.data_ptr(), gx.data_ptr(), B, Cc, HW, ws.data_ptr(), wsb,
                                        L.stream_ptr(x.device)), "ast_gram_bwd_tc")
        else:
            L.check(lib.ast_gram_bwd(x.data_ptr(), gg.data_ptr(), gx.data_ptr(), B, Cc, HW,
                                     L.stream_ptr(x.device)), "ast_gram_bwd")
        return gx


def gram_matrix(x: torch.Tensor) -> torch.Tensor:
    return _Gram.apply(x)
