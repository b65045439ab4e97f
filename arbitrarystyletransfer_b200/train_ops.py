"""Autograd for the training step (BASELINE config 2): the classic decoder (forward + data and
weight gradients) and the frozen VGG-19 encoder (forward + data gradient), both on the native
bf16 layout with every convolution -- forward, dgrad and wgrad -- on the tcgen05 kernels.

Replaces what torch.autograd derives for the reference modules in a decoder training step
(train.py:287-300: zero_grad / backward / clip / Adam stay ordinary PyTorch host code and see
ordinary fp32 ``.grad`` tensors on ordinary ``nn.Parameter``s).
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib as L
from . import engine as E


def _wp(W):
    """planar row pitch: multiple of 8 pixels (16 B) >= W + 2, so row shifts stay TMA-aligned"""
    return (W + 2 + 7) // 8 * 8


def _ldq(N, H, W):
    return N * (H + 2) * _wp(W)


_RSTD = {}


def _imagenet_rstd(dev):
    """1/std of Normalization (models.py:190) as a device tensor, created once per device (a host ->
    device copy inside the backward pass would break CUDA-graph capture)."""
    t = _RSTD.get(str(dev))
    if t is None:
        t = torch.tensor([1.0 / s for s in E.IMAGENET_STD], dtype=torch.float32).to(dev)
        _RSTD[str(dev)] = t
    return t


_ZERO_HALO = {}


def _zeros_native(N, H, W, C, halo, dev, tag=None):
    """Gradient buffer whose halo must be zero.  With a ``tag`` the buffer is cached per (tag, shape):
    the kernels that fill it rewrite every interior element and never touch the halo, so the zeros
    written at allocation stay valid and no per-step memset is needed."""
    shape = (N, H + 2 * halo, W + 2 * halo, C)
    if tag is None:
        return torch.zeros(shape, device=dev, dtype=torch.bfloat16)
    key = (tag, shape, str(dev))
    t = _ZERO_HALO.get(key)
    if t is None:
        t = torch.zeros(shape, device=dev, dtype=torch.bfloat16)
        _ZERO_HALO[key] = t
    return t


def pack_ex(w, flip, rows_pad=0, cols_pad=0, row_scale=None):
    lib = L.load()
    w = w.detach().float().contiguous()
    co, ci = w.shape[:2]
    rows, cols = (ci, co) if flip else (co, ci)
    R, Cc = rows_pad or rows, cols_pad or cols
    out = torch.empty((9, R, Cc), device=w.device, dtype=torch.bfloat16)
    L.check(lib.ast_pack_conv_weight_ex(w.data_ptr(), out.data_ptr(), co, ci, int(flip), R, Cc,
                                        L.ptr(row_scale), L.stream_ptr(w.device)),
            "ast_pack_conv_weight_ex")
    return out


_PACKED = {}


def _packed_cached(p, kind, make):
    """Packed bf16 form of a conv weight, re-made only when the parameter changes (same storage and autograd version
    => same values): the VGG loss network is frozen in every training flow of the reference (train.py:55-56,
    train_autoencoder.py:24-25), so its 2 x 9..16 pack launches per step disappear; a weight an optimiser updates
    in place bumps ``_version`` and is re-packed.  Entries die with their parameter (weakref callback)."""
    key = (id(p), kind)
    ver = (p.data_ptr(), p._version, str(p.device))
    hit = _PACKED.get(key)
    if hit is not None and hit[0]() is p and hit[1] == ver:
        return hit[2]
    out = make()
    _PACKED[key] = (weakref.ref(p, lambda _r, k=key: _PACKED.pop(k, None)), ver, out)
    return out


def to_planar(native, N, C, H, W, src_halo, copy_halo, nshift=1):
    """native -> [C][ldq] (nshift 1) or the three column-shifted copies [3][C][ldq] (nshift 3)."""
    lib = L.load()
    ldq = _ldq(N, H, W)
    shape = (C, ldq) if nshift == 1 else (3, C, ldq)
    out = torch.empty(shape, device=native.device, dtype=torch.bfloat16)
    L.check(lib.ast_native_to_planar(native.data_ptr(), out.data_ptr(), N, C, H, W, src_halo,
                                     int(copy_halo), _wp(W), nshift, L.stream_ptr(native.device)),
            "ast_native_to_planar")
    return out


def conv_wgrad(dz_planar, x_planar3, N, H, W, cin, cout, w_like, b_like):
    """(dW OIHW fp32, db fp32) of one 3x3 conv from the planar operands."""
    lib = L.load()
    dev = dz_planar.device
    st = L.stream_ptr(dev)
    ldq = _ldq(N, H, W)
    dwpk = torch.empty((9, cout, cin), device=dev, dtype=torch.float32)
    L.check(lib.ast_conv3x3_wgrad(dz_planar.data_ptr(), x_planar3.data_ptr(), dwpk.data_ptr(), N, H, W,
                                  cin, cout, _wp(W), st), "ast_conv3x3_wgrad")
    gw = torch.empty_like(w_like, dtype=torch.float32, memory_format=torch.contiguous_format)
    gb = torch.empty_like(b_like, dtype=torch.float32) if b_like is not None else None
    L.check(lib.ast_unpack_wgrad(dwpk.data_ptr(), gw.data_ptr(), dz_planar.data_ptr(), L.ptr(gb), cout,
                                 cin, ldq, 0, st), "ast_unpack_wgrad")
    return gw, gb


def conv_wgrad_native(dz, dz_halo, x, N, H, W, cin, cout, w_like, b_like, need_b=True):
    """(dW OIHW fp32, db fp32) of one 3x3 conv straight from the native tensors (K2wn, csrc/wgrad_mn.cu):
    dz [N][H+2*dz_halo][W+2*dz_halo][cz] with a ZERO halo, x [N][H+2][W+2][cin] with the halo the conv saw."""
    lib = L.load()
    dev = dz.device
    st = L.stream_ptr(dev)
    dwpk = torch.empty((9, cout, cin), device=dev, dtype=torch.float32)
    gb = torch.empty_like(b_like, dtype=torch.float32) if (b_like is not None and need_b) else None
    L.check(lib.ast_conv3x3_wgrad_native(dz.data_ptr(), dz.shape[3], dz_halo, x.data_ptr(), dwpk.data_ptr(),
                                         L.ptr(gb), N, H, W, cin, cout, st), "ast_conv3x3_wgrad_native")
    gw = torch.empty_like(w_like, dtype=torch.float32, memory_format=torch.contiguous_format)
    L.check(lib.ast_unpack_wgrad(dwpk.data_ptr(), gw.data_ptr(), None, None, cout, cin, 0, 0, st),
            "ast_unpack_wgrad")
    return gw, gb


# AST_WGRAD_PLANAR=1: the round-1 weight-gradient path (channel-planar copies + one tap per work item), for A/B runs
_WGRAD_PLANAR = os.environ.get("AST_WGRAD_PLANAR", "0") not in ("", "0")


class DecoderFn(torch.autograd.Function):
    """img = decoder(x) for the classic mirrored decoder (models.py:598-628)."""

    @staticmethod
    def forward(ctx, x, exporting, *params):
        lib = L.load()
        L.require_cuda(x)
        if exporting:
            raise L.AstError("training through the Hardtanh(0,1) export epilogue is not supported")
        N, _, h, w = x.shape
        dev = x.device
        acts = [E.nchw_to_native(x, reflect=True)]          # X_0 (reflection halo)
        sizes = []
        for i in range(8):
            cin, cout, relu, up = E.DECODER_SPEC[i]
            sizes.append((h, w))
            ho, wo = (2 * h, 2 * w) if up else (h, w)
            y = E.native_empty(N, ho, wo, cout, dev, zero_halo=False)
            E.conv3x3(acts[i], E.pack_conv_weight(params[2 * i]), params[2 * i + 1], y, N=N, H=h, W=w,
                      cin=cin, cout=cout, relu=relu, epilogue=L.EPI_UP2 if up else L.EPI_PLAIN,
                      halo=L.HALO_REFLECT)
            acts.append(y)
            h, w = ho, wo
        sizes.append((h, w))
        out = torch.empty(N, 3, h, w, device=dev, dtype=torch.float32)
        wl = params[16].detach().float().contiguous()
        E.conv3x3_last(acts[8], wl, E.pack_conv_weight(wl, cout_pad=16), params[17], out, False)
        ctx.acts, ctx.sizes, ctx.N = acts, sizes, N
        ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, dimg):
        lib = L.load()
        params = ctx.saved_tensors
        acts, sizes, N = ctx.acts, ctx.sizes, ctx.N
        dev = dimg.device
        st = L.stream_ptr(dev)
        if ctx.needs_input_grad[0]:
            raise L.AstError("gradient w.r.t. the decoder's input features is not implemented "
                             "(the AdaIN target t is detached in every training flow of the reference)")
        grads = [None] * 18
        H8, W8 = sizes[8]
        dimg = dimg.float().contiguous()
        # dZ of the last conv = dimg, widened to 64 zero-padded channels, 2-pixel zero halo
        dZ = _zeros_native(N, H8, W8, 64, 2, dev)
        L.check(lib.ast_nchw_to_native_ex(dimg.data_ptr(), dZ.data_ptr(), N, 3, H8, W8, 64, 2, st),
                "ast_nchw_to_native_ex")
        for i in range(8, -1, -1):
            cin, cout, relu, up = E.DECODER_SPEC[i]
            Hi, Wi = sizes[i]
            cz = dZ.shape[3]
            need_w, need_b = ctx.needs_input_grad[2 + 2 * i], ctx.needs_input_grad[3 + 2 * i]
            if need_w or need_b:
                if _WGRAD_PLANAR:
                    dzT = to_planar(dZ, N, cz, Hi, Wi, 2, False)
                    xT = to_planar(acts[i], N, cin, Hi, Wi, 1, True, nshift=3)
                    gw, gb = conv_wgrad(dzT, xT, N, Hi, Wi, cin, cout, params[2 * i], params[2 * i + 1])
                else:
                    gw, gb = conv_wgrad_native(dZ, 2, acts[i], N, Hi, Wi, cin, cout, params[2 * i],
                                               params[2 * i + 1], need_b)
                grads[2 * i] = gw if need_w else None
                grads[2 * i + 1] = gb if need_b else None
            if i == 0:
                break
            # data gradient over the (Hi+2) x (Wi+2) padded grid, then fold pad / upsample / ReLU
            wflip = pack_ex(params[2 * i], flip=True, rows_pad=cin, cols_pad=cz)
            dXpad = torch.empty((N, Hi + 4, Wi + 4, cin), device=dev, dtype=torch.bfloat16)
            E.conv3x3(dZ, wflip, None, dXpad, N=N, H=Hi + 2, W=Wi + 2, cin=cz, cout=cin, relu=False,
                      epilogue=L.EPI_PLAIN, halo=L.HALO_KEEP)
            _, _, prelu, pup = E.DECODER_SPEC[i - 1]
            Hc, Wc = (Hi // 2, Wi // 2) if pup else (Hi, Wi)
            dZp = _zeros_native(N, Hc, Wc, cin, 2, dev, tag=("dec", i - 1))
            L.check(lib.ast_dec_bwd_fold(dXpad.data_ptr(), acts[i].data_ptr(), dZp.data_ptr(), N, cin, Hi,
                                         Wi, int(pup), int(prelu), st), "ast_dec_bwd_fold")
            dZ = dZp
        ctx.acts = None
        return (None, None) + tuple(grads)


class EncoderFn(torch.autograd.Function):
    """Feature taps of the (frozen) VGG-19 encoder with a gradient w.r.t. the input image
    (PretrainedEncoder.forward, models.py:230-240, called on a generated image)."""

    @staticmethod
    def forward(ctx, img, plan, *wb):
        # plan: list of (cin, cout, pool_after, tap) with tap in (None, 'pre', 'post'), one per conv
        lib = L.load()
        L.require_cuda(img)
        img = img.float().contiguous()
        N, _, H, W = img.shape
        dev = img.device
        _imagenet_rstd(dev)   # make sure the constant exists before any graph capture of backward
        Ys, outs = [], []
        x, h, w = None, H, W
        sizes = []
        for i, (cin, cout, pool, tap) in enumerate(plan):
            sizes.append((h, w))
            tapbuf = torch.empty(N, cout, h, w, device=dev, dtype=torch.float32) if tap else None
            y = E.native_empty(N, h, w, cout, dev, zero_halo=True)
            if i == 0:
                E.conv3x3_first(img, wb[0].detach().float().contiguous(), wb[1], y, tap=tapbuf,
                                tap_prerelu=(tap == "pre"))
            else:
                E.conv3x3(x, _packed_cached(wb[2 * i], "fwd", lambda: E.pack_conv_weight(wb[2 * i])), wb[2 * i + 1], y,
                          N=N, H=h, W=w, cin=cin,
                          cout=cout, relu=True, epilogue=L.EPI_PLAIN, halo=L.HALO_KEEP, tap=tapbuf,
                          tap_prerelu=(tap == "pre"))
            Ys.append(y)
            if tapbuf is not None:
                outs.append(tapbuf)
            x = y
            if pool and i + 1 < len(plan):
                pz = E.native_empty(N, h // 2, w // 2, cout, dev, zero_halo=True)
                L.check(lib.ast_maxpool2_native(y.data_ptr(), pz.data_ptr(), N, cout, h, w,
                                                L.stream_ptr(dev)), "ast_maxpool2_native")
                x, h, w = pz, h // 2, w // 2
        ctx.Ys, ctx.sizes, ctx.plan, ctx.N = Ys, sizes, plan, N
        ctx.wb_objs = wb          # the parameter objects themselves: keys of the packed-weight cache in backward
        ctx.save_for_backward(*wb)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gtaps):
        lib = L.load()
        wb = ctx.saved_tensors
        plan, Ys, sizes, N = ctx.plan, ctx.Ys, ctx.sizes, ctx.N
        dev = Ys[0].device
        st = L.stream_ptr(dev)
        gt = list(gtaps)
        tap_grad = {}
        k = 0
        for i, (_, _, _, tap) in enumerate(plan):
            if tap:
                tap_grad[i] = gt[k]
                k += 1
        G = None
        dimg = None
        for i in range(len(plan) - 1, -1, -1):
            cin, cout, pool, tap = plan[i]
            Hi, Wi = sizes[i]
            tpost = tpre = None
            g = tap_grad.get(i)
            if g is not None:
                tn = E.nchw_to_native(g.float().contiguous(), reflect=False)
                if tap == "pre":
                    tpre = tn
                else:
                    tpost = tn
            if G is None and tpost is None and tpre is None:
                continue  # nothing flows into this layer (deeper than the last tap with a gradient)
            pooled = pool and i + 1 < len(plan) and G is not None
            dZ = _zeros_native(N, Hi, Wi, cout, 1, dev, tag=("enc", i))
            L.check(lib.ast_vgg_bwd_prep(Ys[i].data_ptr(), L.ptr(G), L.ptr(tpost), L.ptr(tpre),
                                         dZ.data_ptr(), N, cout, Hi, Wi, int(pooled), 1, st),
                    "ast_vgg_bwd_prep")
            if i > 0:
                wflip = _packed_cached(ctx.wb_objs[2 * i], "flip", lambda: pack_ex(wb[2 * i], flip=True))
                G = torch.empty((N, Hi + 2, Wi + 2, cin), device=dev, dtype=torch.bfloat16)
                E.conv3x3(dZ, wflip, None, G, N=N, H=Hi, W=Wi, cin=cout, cout=cin, relu=False,
                          epilogue=L.EPI_PLAIN, halo=L.HALO_KEEP)
            else:
                wflip = _packed_cached(ctx.wb_objs[0], "flip0", lambda: pack_ex(wb[0], flip=True, rows_pad=16, cols_pad=64,
                                                                                row_scale=_imagenet_rstd(dev)))
                dimg = torch.empty(N, 3, Hi, Wi, device=dev, dtype=torch.float32)
                E.conv3x3_last(dZ, None, wflip, None, dimg, False, impl=L.CONV_TC)
        ctx.Ys = None
        return (dimg, None) + (None,) * len(wb)
