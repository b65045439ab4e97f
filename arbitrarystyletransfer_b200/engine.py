"""Native-layout execution of the classic VGG-19 relu4_1 -> AdaIN -> mirrored decoder path
(reference pieces: models.py:186-240 encoder, models.py:43-51 AdaIN,
models.py:471 alpha blend, models.py:598-628 decoder spec).

Activations live between layers as bf16 [N][H+2][W+2][C] whose one-pixel halo already holds the
padding of the consuming conv (zeros inside VGG, the reflection inside the decoder), so every
3x3 tap of the tcgen05 implicit-GEMM kernel is one shifted TMA box.  fp32 NCHW exists only at the
boundary: images in, image out, and optional feature taps.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import torch

from . import _lib as L

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # models.py:189
IMAGENET_STD = (0.229, 0.224, 0.225)    # models.py:190

# torchvision VGG-19 configuration 'E' (what models.py:192 instantiates), 'M' = MaxPool2d(2, 2)
VGG19_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M",
             512, 512, 512, 512, "M", 512, 512, 512, 512, "M")
# (cin, cout, relu, upsample_after) of the 9 decoder convs, models.py:598-628 / conf.py:9
DECODER_SPEC = ((512, 256, True, True), (256, 256, True, False), (256, 256, True, False),
                (256, 256, True, False), (256, 128, True, True), (128, 128, True, False),
                (128, 64, True, True), (64, 64, True, False), (64, 3, False, False))


# ---- optional per-launch timing (bench.py's roofline): CUDA events around every kernel launch of a stylise pass,
# recorded on the launching stream.  None (the default) costs one comparison per launch.
_PROFILE = None


class _Span:
    __slots__ = ("name", "a")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if _PROFILE is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _PROFILE.append((self.name, self.a, b))


def profile_launches(records):
    """``records``: a list that receives (name, start_event, end_event) per kernel launch, or None to stop."""
    global _PROFILE
    _PROFILE = records


def vgg_layer_plan(n_convs: int):
    """[(cin, cout, pool_after)] for the first ``n_convs`` VGG-19 convs."""
    plan, cin = [], 3
    cfg = list(VGG19_CFG)
    i = 0
    while i < len(cfg) and len(plan) < n_convs:
        v = cfg[i]
        if v != "M":
            pool = i + 1 < len(cfg) and cfg[i + 1] == "M"
            plan.append((cin, v, pool))
            cin = v
        i += 1
    return plan


def native_empty(N, H, W, Cc, device, zero_halo: bool):
    """bf16 [N][H+2][W+2][C]; ``zero_halo`` buffers get a zero halo ring (ast_zero_halo) and are only ever
    written in their interior, so the halo keeps VGG's zero padding.  The interior is NOT initialised."""
    shape = (N, H + 2, W + 2, Cc)
    t = torch.empty(shape, device=device, dtype=torch.bfloat16)
    if zero_halo:
        if Cc % 8:
            return t.zero_()
        # only the ring: every producer (conv epilogue, max-pool) rewrites the whole interior
        L.check(L.load().ast_zero_halo(t.data_ptr(), N, Cc, H, W, 1, L.stream_ptr(t.device)), "ast_zero_halo")
    return t


def pack_conv_weight(w: torch.Tensor, flip: bool = False, cout_pad: int = 0) -> torch.Tensor:
    """OIHW fp32 -> bf16 [9][Cout][Cin] (or the data-gradient form when ``flip``; or zero-padded
    to ``cout_pad`` output rows for the 16-wide last-layer kernel)."""
    lib = L.load()
    L.require_cuda(w)
    w = w.detach().float().contiguous()
    co, ci = w.shape[:2]
    rows = cout_pad or co
    out = torch.empty((9, ci, co) if flip else (9, rows, ci), device=w.device, dtype=torch.bfloat16)
    L.check(lib.ast_pack_conv_weight(w.data_ptr(), out.data_ptr(), co, ci, int(flip), int(cout_pad),
                                     L.stream_ptr(w.device)), "ast_pack_conv_weight")
    return out


def pack_conv_weight_fold(w: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> bf16 [16][Cout][Cin]: the pre-summed 2x2 taps of the four output parities of
    Upsample(x2) -> ReflectionPad2d(1) -> Conv2d(3x3) (models.py:602-604, 616-618, 622-624)."""
    lib = L.load()
    L.require_cuda(w)
    w = w.detach().float().contiguous()
    co, ci = w.shape[:2]
    out = torch.empty((16, co, ci), device=w.device, dtype=torch.bfloat16)
    L.check(lib.ast_pack_conv_weight_fold(w.data_ptr(), out.data_ptr(), co, ci, L.stream_ptr(w.device)),
            "ast_pack_conv_weight_fold")
    return out


# decoder convs that follow an upsample (DECODER_SPEC[i-1][3]): evaluated on the low-res map (AST_EPI_UPFOLD)
FOLD_LAYERS = tuple(i for i in range(1, 9) if DECODER_SPEC[i - 1][3] and DECODER_SPEC[i][0] % 64 == 0
                    and DECODER_SPEC[i][1] % 64 == 0)


def run_decoder(buf, t, wpk, wfold, biases, w_last, wpk_last, clamp01=False, out=None, impl=L.CONV_AUTO,
                impl_edge=L.CONV_AUTO, fold=True, key="dec"):
    """The nine decoder convs (models.py:598-628) on the native layout: t = bf16 [N][h+2][w+2][512] with its
    reflection halo -> (N,3,8h,8w) fp32.  With ``fold`` (default) a conv that follows an Upsample is evaluated as
    four 2x2 convs on the low-res map: its producer writes the low-res tensor with a clamp halo (EPI_PLAIN) instead
    of the x2-replicated tensor (EPI_UP2), and the upsampled activation never exists in HBM."""
    N, hp, wp, _ = t.shape
    h, w = hp - 2, wp - 2
    dev = t.device
    x = t
    for i in range(8):
        cin, cout, relu, up = DECODER_SPEC[i]
        folded_in = fold and i in FOLD_LAYERS            # x is the low-res map of an upsampled input
        folded_out = fold and up and (i + 1) in FOLD_LAYERS   # the next conv folds our upsample
        hi, wi = h, w                                    # this conv's input grid (low-res when folded_in)
        if folded_in:
            ho, wo = 2 * h, 2 * w
        elif up and not folded_out:
            ho, wo = 2 * h, 2 * w
        else:
            ho, wo = h, w
        y = buf.get(f"{key}{i}" + ("f" if fold else ""), N, ho, wo, cout, dev, False)
        with _Span(f"dec_conv{i + 1}"):
            if folded_in:
                conv3x3(x, wfold[i], biases[i], y, N=N, H=hi, W=wi, cin=cin, cout=cout, relu=relu,
                        epilogue=L.EPI_UPFOLD, halo=L.HALO_CLAMP if folded_out else L.HALO_REFLECT, impl=impl)
            else:
                conv3x3(x, wpk[i], biases[i], y, N=N, H=hi, W=wi, cin=cin, cout=cout, relu=relu,
                        epilogue=L.EPI_PLAIN if (folded_out or not up) else L.EPI_UP2,
                        halo=L.HALO_CLAMP if folded_out else L.HALO_REFLECT, impl=impl)
        x, h, w = y, ho, wo                              # after a folded_out layer (h, w) stay low-res
    if out is None:
        out = torch.empty(N, 3, h, w, device=dev, dtype=torch.float32)
    with _Span("dec_conv9"):
        conv3x3_last(x, w_last, wpk_last, biases[8], out, clamp01, impl=impl_edge)
    return out


def conv3x3_first(img, w, bias, out_native, tap=None, tap_prerelu=True, normalise=True,
                  impl=L.CONV_AUTO):
    """Normalization + conv_1 + relu_1 (models.py:129-131, 198-224) from NCHW fp32."""
    lib = L.load()
    N, _, H, W = img.shape
    mean = L.float_array(IMAGENET_MEAN) if normalise else None
    std = L.float_array(IMAGENET_STD) if normalise else None
    L.check(lib.ast_conv3x3_first(img.data_ptr(), w.data_ptr(), L.ptr(bias), mean, std,
                                  L.ptr(out_native), L.ptr(tap), int(tap_prerelu), N, H, W,
                                  w.shape[0], impl, L.stream_ptr(img.device)), "ast_conv3x3_first")
    return out_native


def conv12_fused(img, w1, b1, wpk2, b2, out_native, normalise=True):
    """Normalization + conv_1 + relu_1 + conv_2 + relu_2 + pool_2 (models.py:129-131, 198-224) in one kernel:
    (N,3,H,W) fp32 -> native bf16 [N][H/2+2][W/2+2][64]; the full-resolution 64-channel map never reaches HBM."""
    lib = L.load()
    N, _, H, W = img.shape
    mean = L.float_array(IMAGENET_MEAN) if normalise else None
    std = L.float_array(IMAGENET_STD) if normalise else None
    L.check(lib.ast_conv12_fused(img.data_ptr(), w1.data_ptr(), b1.data_ptr(), mean, std, wpk2.data_ptr(),
                                 b2.data_ptr(), out_native.data_ptr(), N, H, W, L.stream_ptr(img.device)),
            "ast_conv12_fused")
    return out_native


def conv3x3_last(x_native, w, wpk16, bias, out, clamp01=False, impl=L.CONV_AUTO):
    """Last decoder conv (models.py:626-627): native bf16 with reflection halo -> NCHW fp32."""
    lib = L.load()
    N, Cout, H, W = out.shape
    L.check(lib.ast_conv3x3_last(x_native.data_ptr(), L.ptr(w), L.ptr(wpk16), L.ptr(bias),
                                 out.data_ptr(), N, H, W, x_native.shape[3], Cout, int(clamp01), impl,
                                 L.stream_ptr(out.device)), "ast_conv3x3_last")
    return out


def conv3x3(x_native: torch.Tensor, wpk: torch.Tensor, bias, out_native, *, N, H, W, cin, cout,
            relu=True, epilogue=L.EPI_PLAIN, halo=L.HALO_KEEP, impl=L.CONV_AUTO, tap=None,
            tap_prerelu=True):
    lib = L.load()
    d = L.ConvDesc(N, H, W, cin, cout, int(relu), epilogue, halo, impl, int(tap_prerelu))
    L.check(lib.ast_conv3x3_fwd(C.byref(d), x_native.data_ptr(), wpk.data_ptr(), L.ptr(bias),
                                L.ptr(out_native), L.ptr(tap), L.stream_ptr(x_native.device)),
            "ast_conv3x3_fwd")
    return out_native


def nchw_to_native(x: torch.Tensor, reflect: bool, out=None) -> torch.Tensor:
    lib = L.load()
    L.require_cuda(x)
    x = x.float().contiguous()
    N, Cc, H, W = x.shape
    if out is None:
        out = native_empty(N, H, W, Cc, x.device, zero_halo=not reflect)
    L.check(lib.ast_nchw_to_native(x.data_ptr(), out.data_ptr(), N, Cc, H, W,
                                   L.HALO_REFLECT if reflect else L.HALO_KEEP,
                                   L.stream_ptr(x.device)), "ast_nchw_to_native")
    return out


def native_to_nchw(x_native: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    N, Hp, Wp, Cc = x_native.shape
    out = torch.empty(N, Cc, Hp - 2, Wp - 2, device=x_native.device, dtype=torch.float32)
    L.check(lib.ast_native_to_nchw(x_native.data_ptr(), out.data_ptr(), N, Cc, Hp - 2, Wp - 2,
                                   L.stream_ptr(x_native.device)), "ast_native_to_nchw")
    return out


def u8_to_nchw(img_u8: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 [N][H][W][3] (what PIL / the loader holds before transforms.ToTensor(), data_loader.py:114) ->
    fp32 [N][3][H][W] in [0, 1] = u8 / 255, on the device."""
    lib = L.load()
    L.require_cuda(img_u8)
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[3] != 3:
        raise L.AstError("expected a uint8 (N, H, W, 3) image batch")
    img_u8 = img_u8.contiguous()
    N, H, W, _ = img_u8.shape
    if out is None:
        out = torch.empty(N, 3, H, W, device=img_u8.device, dtype=torch.float32)
    L.check(lib.ast_u8hwc_to_nchw(img_u8.data_ptr(), out.data_ptr(), N, H, W, L.stream_ptr(img_u8.device)),
            "ast_u8hwc_to_nchw")
    return out


def nchw_to_u8(img: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 [N][3][H][W] -> uint8 [N][H][W][3] = trunc(clamp(x, 0, 1) * 255): the exporting decoder's Hardtanh(0,1)
    (models.py:315-316) followed by transforms.ToPILImage() (train.py:18)."""
    lib = L.load()
    L.require_cuda(img)
    if img.dtype != torch.float32 or img.dim() != 4 or img.shape[1] != 3:
        raise L.AstError("expected an fp32 (N, 3, H, W) image batch")
    img = img.contiguous()
    N, _, H, W = img.shape
    if out is None:
        out = torch.empty(N, H, W, 3, device=img.device, dtype=torch.uint8)
    L.check(lib.ast_nchw_to_u8hwc(img.data_ptr(), out.data_ptr(), N, H, W, L.stream_ptr(img.device)),
            "ast_nchw_to_u8hwc")
    return out


class _Buffers:
    """Shape-keyed cache of activation buffers (torch's caching allocator owns the memory)."""

    def __init__(self):
        self._b = {}

    def get(self, key, N, H, W, Cc, device, zero_halo):
        k = (key, N, H, W, Cc, str(device), zero_halo)
        t = self._b.get(k)
        if t is None:
            t = native_empty(N, H, W, Cc, device, zero_halo)
            self._b[k] = t
        return t

    def clear(self):
        self._b.clear()


class StyleTransferEngine:
    """Forward engine for configs 1 / 4 / 5: images (N,3,H,W) fp32 in [0,1] -> stylised images.

    ``vgg_w/vgg_b``: the first 9 VGG-19 conv weights / biases (OIHW fp32) -> relu4_1;
    ``dec_w/dec_b``: the 9 classic-decoder convs.  Weights are packed to bf16 [9][Cout][Cin] once.
    """

    def __init__(self, vgg_w: Sequence[torch.Tensor], vgg_b, dec_w, dec_b, device="cuda",
                 conv_impl: int = L.CONV_AUTO):
        L.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.AstError("StyleTransferEngine needs a CUDA device (no CPU fallback)")
        self.impl = conv_impl
        self.impl_edge = L.CONV_AUTO  # first / last layer: tensor cores unless set to CONV_DIRECT
        self.plan = vgg_layer_plan(9)
        dev = self.device
        self.vgg_w0 = vgg_w[0].detach().to(dev, torch.float32).contiguous()
        self.vgg_b = [b.detach().to(dev, torch.float32).contiguous() for b in vgg_b[:9]]
        self.vgg_wpk = [None] + [pack_conv_weight(w.to(dev)) for w in vgg_w[1:9]]
        self.dec_b = [b.detach().to(dev, torch.float32).contiguous() for b in dec_b]
        self.dec_wpk = [pack_conv_weight(w.to(dev)) for w in dec_w[:8]] + [None]
        self.dec_wfold = {i: pack_conv_weight_fold(dec_w[i].to(dev)) for i in FOLD_LAYERS}
        self.fold = True      # evaluate the three post-upsample convs on the low-res map (AST_EPI_UPFOLD)
        # conv1_1 + conv1_2 (+ pool) in one kernel (ast_conv12_fused) where its shape rules hold; AST_CONV12_FUSED=0
        # keeps the two separate launches (A/B reference)
        self.fuse12 = os.environ.get("AST_CONV12_FUSED", "1") != "0"
        self.dec_w_last = dec_w[8].detach().to(dev, torch.float32).contiguous()
        self.dec_wpk_last = pack_conv_weight(self.dec_w_last, cout_pad=16)
        self.buf = _Buffers()
        self._f32buf = {}
        self._ws = None

    # ---- encoder: models.py:230-240 with content_layers=['relu_9'] ------------------------------
    def encode(self, img: torch.Tensor, key: str = "c") -> torch.Tensor:
        """(N,3,H,W) fp32 -> relu4_1 in native layout bf16 [N][H/8+2][W/8+2][512] (zero halo)."""
        lib = L.load()
        L.require_cuda(img)
        img = img.float().contiguous()
        N, c3, H, W = img.shape
        if c3 != 3:
            raise L.AstError("images must have 3 channels")
        dev = img.device
        st = L.stream_ptr(dev)
        h, w = H, W
        first = 1
        if self.fused12_ok(H, W):
            x = self.buf.get("enc1", N, H // 2, W // 2, 64, dev, True)
            with _Span("enc_conv12"):
                conv12_fused(img, self.vgg_w0, self.vgg_b[0], self.vgg_wpk[1], self.vgg_b[1], x)
            h, w, first = H // 2, W // 2, 2
        else:
            x = self.buf.get("enc0", N, H, W, 64, dev, True)
            with _Span("enc_conv1"):
                conv3x3_first(img, self.vgg_w0, self.vgg_b[0], x, impl=self.impl_edge)
        for i in range(first, 9):
            cin, cout, pool = self.plan[i]
            ho, wo = (h // 2, w // 2) if pool else (h, w)
            last = i == 8
            y = self.buf.get(f"enc{i}" + (key if last else ""), N, ho, wo, cout, dev, True)
            with _Span(f"enc_conv{i + 1}"):
                conv3x3(x, self.vgg_wpk[i], self.vgg_b[i], y, N=N, H=h, W=w, cin=cin, cout=cout,
                        relu=True, epilogue=L.EPI_POOL2 if pool else L.EPI_PLAIN, halo=L.HALO_KEEP,
                        impl=self.impl)
            x, h, w = y, ho, wo
        return x

    # ---- AdaIN on the native layout ---------------------------------------------------------------
    def adain(self, fc: torch.Tensor, fs: Sequence[torch.Tensor], weights, alpha=1.0,
              canonical=False) -> torch.Tensor:
        lib = L.load()
        N, Hp, Wp, Cc = fc.shape
        K = len(fs)
        for s in fs:
            if s.shape[0] != N or s.shape[3] != Cc:
                raise L.AstError("style feature maps must share the content's batch size and channel count")
        Hs = (C.c_int * K)(*[s.shape[1] - 2 for s in fs])
        Ws = (C.c_int * K)(*[s.shape[2] - 2 for s in fs])
        wsb = lib.ast_adain_native_ws_bytes(N, Cc, K)
        if self._ws is None or self._ws.numel() < wsb or self._ws.device != fc.device:
            self._ws = torch.empty(wsb, device=fc.device, dtype=torch.uint8)
        out = self.buf.get("adain", N, Hp - 2, Wp - 2, Cc, fc.device, False)
        sp = (C.c_void_p * K)(*[s.data_ptr() for s in fs])
        span = _Span("adain_native")
        span.__enter__()
        L.check(lib.ast_adain_native_fwd(fc.data_ptr(), sp, L.float_array(weights), K,
                                         out.data_ptr(), N, Cc, Hp - 2, Wp - 2, Hs, Ws,
                                         float(alpha), 0.0, L.F_CANONICAL if canonical else 0,
                                         L.HALO_REFLECT, self._ws.data_ptr(), self._ws.numel(),
                                         L.stream_ptr(fc.device)), "ast_adain_native_fwd")
        span.__exit__()
        return out

    # ---- decoder: models.py:598-628 ------------------------------------------------------------------
    def decode(self, t: torch.Tensor, clamp01: bool = False, out: torch.Tensor | None = None):
        """native bf16 [N][h+2][w+2][512] with reflection halo -> (N,3,8h,8w) fp32."""
        return run_decoder(self.buf, t, self.dec_wpk, self.dec_wfold, self.dec_b, self.dec_w_last, self.dec_wpk_last,
                           clamp01, out, self.impl, self.impl_edge, self.fold)

    # ---- full path ------------------------------------------------------------------------------
    def stylize(self, content: torch.Tensor, styles, alpha: float = 1.0, style_weights=None,
                canonical: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
        """content (N,3,H,W); styles: tensor (N,3,Hs,Ws) or list of K such tensors."""
        if isinstance(styles, torch.Tensor):
            styles = [styles]
        K = len(styles)
        if style_weights is None:
            style_weights = [1.0 / K] * K
        fc = self.encode(content, "c")
        fs = [self.encode(s, f"s{k}") for k, s in enumerate(styles)]
        t = self.adain(fc, fs, style_weights, alpha, canonical)
        return self.decode(t, out=out)

    def stylize_u8(self, content_u8: torch.Tensor, styles_u8, alpha: float = 1.0, style_weights=None,
                   canonical: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
        """The same path with byte images at the boundary: uint8 (N,H,W,3) content / style(s)
        as PIL or the loader hold them -> uint8 (N,H,W,3) stylised images.  Equals ``nchw_to_u8(stylize(u8/255))``
        exactly: the byte <-> float conversions run on the device, so images cross PCIe at 3 bytes per pixel."""
        if isinstance(styles_u8, torch.Tensor):
            styles_u8 = [styles_u8]
        N, H, W, _ = content_u8.shape
        dev = content_u8.device
        c = u8_to_nchw(content_u8, self._f32("u8c", N, H, W, dev))
        ss = [u8_to_nchw(s, self._f32(f"u8s{k}", s.shape[0], s.shape[1], s.shape[2], dev))
              for k, s in enumerate(styles_u8)]
        img = self.stylize(c, ss, alpha, style_weights, canonical, out=self._f32("u8o", N, H, W, dev))
        return nchw_to_u8(img, out)

    def _f32(self, key, N, H, W, dev):
        k = (key, N, H, W, str(dev))
        t = self._f32buf.get(k)
        if t is None:
            t = torch.empty(N, 3, H, W, device=dev, dtype=torch.float32)
            self._f32buf[k] = t
        return t

    def fused12_ok(self, H: int, W: int) -> bool:
        """Whether encode() runs the first two VGG layers as one kernel at this image size."""
        return (self.fuse12 and self.impl == L.CONV_AUTO and self.impl_edge == L.CONV_AUTO and W % 4 == 0
                and H % 2 == 0 and H >= 2 and W >= 4 and self.plan[1] == (64, 64, True))

    def launches_per_stylize(self, K: int = 1, H: int = 512, W: int = 512) -> int:
        """Kernels of libast_b200 launched by one stylize() call (for bench.py's gpu_launches)."""
        return (1 + K) * (8 if self.fused12_ok(H, W) else 9) + 3 + 9


class HostPipeline:
    """Streaming host API for batch inference (BASELINE config 4): pinned host batches in, pinned
    host images out, with the PCIe copies of step i+1 / i-1 overlapped with the kernels of step i.

    Three streams: H2D copies, compute (the caller's current stream), D2H copies; two slots of
    device input / output buffers.  ``submit`` only enqueues work; ``synchronize`` (or the next
    ``submit`` that reuses a slot) makes the corresponding ``out_host`` valid.  Every step still
    moves its own inputs and its own result across PCIe -- only the waiting is overlapped.
    """

    def __init__(self, engine: StyleTransferEngine, N: int, H: int, W: int, slots: int = 2, dtype: str = "f32"):
        """``dtype="f32"``: host tensors are the reference's fp32 (N,3,H,W) images.  ``dtype="u8"``: host tensors are
        uint8 (N,H,W,3) -- what the loader holds before ToTensor and what an image writer wants -- and the
        byte <-> float conversions run on the device: 4x fewer bytes over PCIe and through host memory."""
        self.eng = engine
        dev = engine.device
        self.dev = dev
        self.slots = slots
        if dtype not in ("f32", "u8"):
            raise L.AstError("HostPipeline dtype must be 'f32' or 'u8'")
        self.dtype = dtype
        shape, dt = ((N, 3, H, W), torch.float32) if dtype == "f32" else ((N, H, W, 3), torch.uint8)
        self.c = [torch.empty(shape, device=dev, dtype=dt) for _ in range(slots)]
        self.s = [torch.empty(shape, device=dev, dtype=dt) for _ in range(slots)]
        self.o = [torch.empty(shape, device=dev, dtype=dt) for _ in range(slots)]
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]      # inputs landed
        self.ev_done = [torch.cuda.Event() for _ in range(slots)]    # kernels finished
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]     # result copied out
        self.i = 0

    def submit(self, content_host: torch.Tensor, style_host: torch.Tensor, out_host: torch.Tensor,
               alpha: float = 1.0):
        if not (content_host.is_pinned() and style_host.is_pinned() and out_host.is_pinned()):
            raise L.AstError("HostPipeline needs pinned host tensors (torch.Tensor.pin_memory())")
        if content_host.dtype != self.c[0].dtype or content_host.shape != self.c[0].shape:
            raise L.AstError(f"HostPipeline(dtype={self.dtype!r}) expects host tensors of shape "
                             f"{tuple(self.c[0].shape)} and dtype {self.c[0].dtype}")
        k = self.i % self.slots
        compute = torch.cuda.current_stream(self.dev)
        if self.i >= self.slots:
            # slot reuse: its previous kernels must have consumed the inputs, its result must be out
            self.h2d.wait_event(self.ev_done[k])
            compute.wait_event(self.ev_out[k])
        with torch.cuda.stream(self.h2d):
            self.c[k].copy_(content_host, non_blocking=True)
            self.s[k].copy_(style_host, non_blocking=True)
            self.ev_in[k].record(self.h2d)
        compute.wait_event(self.ev_in[k])
        if self.dtype == "u8":
            self.eng.stylize_u8(self.c[k], self.s[k], alpha=alpha, out=self.o[k])
        else:
            self.eng.stylize(self.c[k], self.s[k], alpha=alpha, out=self.o[k])
        self.ev_done[k].record(compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.ev_done[k])
            out_host.copy_(self.o[k], non_blocking=True)
            self.ev_out[k].record(self.d2h)
        self.i += 1

    def synchronize(self):
        self.h2d.synchronize()
        self.d2h.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
