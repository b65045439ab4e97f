"""Drop-in for the AdaIN hot-path part of the reference's models.py.

Same class / function names, constructor and ``forward`` signatures, return conventions and
state-dict keys as the reference (paths relative to /root/reference):
  AdaIN                models.py:37-51     (incl. the swapped style-statistics unpack at :44)
  calc_mean_std        models.py:54-62
  mean_variance_norm   models.py:64-68
  PretrainedEncoder    models.py:186-240   (VGG-19 features, taps by name, early return)
  ClassicDecoder       models.py:598-628   (the commented nn.Sequential spec, same key indices)
Parameters stay ordinary fp32 OIHW ``nn.Parameter``s, so optimisers, clipping and checkpoints
written for the reference work unchanged; packed bf16 kernel weights are derived caches that are
rebuilt whenever a parameter's version counter changes.  All device work goes through
libast_b200.so; CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import engine as E
from . import functional as Fn
from .model_util import channel_stats  # re-exported like the reference's star import

__all__ = ["AdaIN", "calc_mean_std", "mean_variance_norm", "PretrainedEncoder", "ClassicDecoder",
           "StyleTransferNet", "channel_stats", "AdaAttN", "AST", "Encoder", "Decoder", "DecoderBlock", "AutoEncoder"]


class AdaIN(nn.Module):
    """models.py:37-51.  ``forward(content_map, style_map)``.

    The reference binds ``style_std, style_mean = channel_stats(style_map)`` (models.py:44), i.e.
    it scales by the style MEAN and shifts by the style STD; that is the default here.
    ``canonical=True`` gives the Huang & Belongie form.  ``alpha`` / multi-style interpolation
    (models.py:471 and BASELINE config 5) are fused into the same kernel pass."""

    def __init__(self, canonical: bool = False):
        super().__init__()
        self.canonical = canonical

    def forward(self, content_map, style_map, alpha: float = 1.0, style_weights=None):
        styles = list(style_map) if isinstance(style_map, (list, tuple)) else [style_map]
        if torch.is_grad_enabled() and (content_map.requires_grad or any(s.requires_grad for s in styles)):
            return _adain_autograd(content_map, styles, style_weights, alpha, self.canonical)
        return Fn.adain_forward(content_map, styles, style_weights, alpha, self.canonical)


def _adain_autograd(content, styles, weights, alpha, canonical):
    """Differentiable AdaIN: the fused forward kernel with a backward made of kernels only
    (functional._AdaIN: ast_adain_bwd + ast_channel_stats_bwd); used when a feature map requires grad."""
    return Fn.adain_autograd(content, styles, weights, alpha, canonical)


def calc_mean_std(feat, eps=1e-5):
    """models.py:54-62: unbiased var + eps -> sqrt; mean.  Returns (mean, std), each (N,C,1,1)."""
    size = feat.size()
    assert (len(size) == 4)
    N, C = size[:2]
    mean, std = Fn.channel_stats_flat(feat, eps=eps, biased=False)
    return mean.view(N, C, 1, 1).to(feat.dtype), std.view(N, C, 1, 1).to(feat.dtype)


def mean_variance_norm(feat):
    """models.py:64-68: (feat - mean) / sqrt(var + 1e-5), one fused pass, differentiable."""
    assert feat.dim() == 4
    return Fn.mean_variance_norm(feat, eps=1e-5)


class _Named(nn.Module):
    """Placeholder for the non-parametric VGG layers (norm / relu / pool); keeps ModuleList
    indices -- and therefore state-dict keys -- identical to the reference."""

    def __init__(self, name):
        super().__init__()
        self.name = name


class _WeightCache:
    def __init__(self):
        self._c = {}

    def packed(self, p: torch.Tensor, flip=False, cout_pad=0):
        key = (id(p), flip, cout_pad)
        ver = (p.data_ptr(), p._version, str(p.device))
        hit = self._c.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, E.pack_conv_weight(p, flip, cout_pad))
            self._c[key] = hit
        return hit[1]

    def folded(self, p: torch.Tensor):
        key = (id(p), "fold")
        ver = (p.data_ptr(), p._version, str(p.device))
        hit = self._c.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, E.pack_conv_weight_fold(p))
            self._c[key] = hit
        return hit[1]


class PretrainedEncoder(nn.Module):
    """models.py:186-240.  VGG-19 ``features`` behind ImageNet normalisation; layers are named
    conv_i / relu_i / pool_i (i = running conv index); ``forward`` returns the outputs of the
    layers named in ``content_layers`` in network order and stops once all are collected.

    The reference hard-codes ``pretrained=True`` (models.py:192); there is no network here, so
    weights are torchvision's default VGG init and ``load_state_dict`` accepts the reference's
    keys (``_vgg_layers.{1,3,6,...}.{weight,bias}``)."""

    def __init__(self, content_layers=['conv_1', 'conv_3', 'conv_5', 'conv_9', 'conv_13', 'relu_15'],
                 weights_path=None):
        super().__init__()
        self._content_layers = set(content_layers)
        layers = [_Named("norm")]
        cin, i = 3, 0
        for v in E.VGG19_CFG:
            if v == "M":
                layers.append(_Named(f"pool_{i}"))
            else:
                i += 1
                conv = nn.Conv2d(cin, v, kernel_size=3, padding=1)
                nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(conv.bias, 0)
                conv.name = f"conv_{i}"
                layers.append(conv)
                layers.append(_Named(f"relu_{i}"))
                cin = v
        self._vgg_layers = nn.ModuleList(layers)
        self._cache = _WeightCache()
        self._buf = E._Buffers()
        self.conv_impl = L.CONV_AUTO
        if weights_path is not None:
            self.load_vgg19_weights(weights_path)

    def load_vgg19_weights(self, path_or_state):
        """Offline replacement for ``models.vgg19(pretrained=True)`` (models.py:192): load the 16 conv weights / biases
        from a file saved on a machine that has them -- either a torchvision ``vgg19().state_dict()`` (keys
        ``features.{0,2,5,...}.{weight,bias}``; ``classifier.*`` is ignored), the ``features`` sub-module's own state
        dict (``{0,2,5,...}.weight``), or this module's / the reference's keys (``_vgg_layers.{1,3,6,...}.weight``).
        torchvision's ``features[i]`` is ``_vgg_layers[i + 1]`` here (the reference puts Normalization at index 0)."""
        sd = path_or_state
        if not isinstance(sd, dict):
            sd = torch.load(path_or_state, map_location="cpu")
        if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
            sd = sd["state_dict"]
        own = self.state_dict()
        new, used = {}, 0
        for k, v in sd.items():
            parts = k.split(".")
            if parts[0] == "_vgg_layers":
                tgt = k
            elif parts[0] == "features" and len(parts) == 3 and parts[1].isdigit():
                tgt = f"_vgg_layers.{int(parts[1]) + 1}.{parts[2]}"
            elif len(parts) == 2 and parts[0].isdigit():
                tgt = f"_vgg_layers.{int(parts[0]) + 1}.{parts[1]}"
            else:
                continue
            if tgt in own:
                if tuple(own[tgt].shape) != tuple(v.shape):
                    raise L.AstError(f"VGG-19 weight {k}: shape {tuple(v.shape)} != {tuple(own[tgt].shape)}")
                new[tgt] = v
                used += 1
        if used != len(own):
            raise L.AstError(f"VGG-19 weights file supplies {used} of the {len(own)} conv tensors")
        self.load_state_dict(new, strict=True)
        return self

    def _convs(self):
        return [m for m in self._vgg_layers if isinstance(m, nn.Conv2d)]

    def forward(self, x):
        lib = L.load()
        L.require_cuda(x)
        wanted = self._content_layers
        names = [m.name for m in self._vgg_layers]
        last_needed = max((k for k, nm in enumerate(names) if nm in wanted), default=-1)
        if last_needed < 0:
            return []
        if torch.is_grad_enabled() and x.requires_grad:
            return self._forward_autograd(x, names, last_needed)
        x = x.float().contiguous()
        N, _, H, W = x.shape
        dev = x.device
        st = L.stream_ptr(dev)
        outs = []
        convs = self._convs()
        cur, h, w, c = None, H, W, 3
        k = 1  # index into _vgg_layers (0 is the norm layer, fused into conv_1)
        ci = 0
        while k <= last_needed:
            conv = self._vgg_layers[k]
            assert isinstance(conv, nn.Conv2d)
            cname, rname = names[k], names[k + 1]
            has_pool = k + 2 < len(names) and names[k + 2].startswith("pool_")
            pname = names[k + 2] if has_pool else None
            want_c, want_r = cname in wanted, rname in wanted
            if want_c and want_r:
                raise L.AstError("tapping both conv_i and relu_i of the same layer is not supported")
            cout = conv.out_channels
            tap = (torch.empty(N, cout, h, w, device=dev, dtype=torch.float32)
                   if (want_c or want_r) else None)
            more = last_needed > k + 1  # anything after this conv+relu pair?
            fuse_pool = has_pool and more
            ho, wo = (h // 2, w // 2) if fuse_pool else (h, w)
            out = self._buf.get(f"v{ci}", N, ho, wo, cout, dev, True) if more else None
            if ci == 0:
                E.conv3x3_first(x, conv.weight, conv.bias, out, tap=tap, tap_prerelu=want_c)
                if fuse_pool:
                    raise L.AstError("VGG-19 never pools right after conv_1")
            else:
                E.conv3x3(cur, self._cache.packed(conv.weight), conv.bias, out, N=N, H=h, W=w,
                          cin=c, cout=cout, relu=True,
                          epilogue=L.EPI_POOL2 if fuse_pool else L.EPI_PLAIN, halo=L.HALO_KEEP,
                          impl=self.conv_impl, tap=tap, tap_prerelu=want_c)
            if tap is not None:
                outs.append(tap)
            cur, h, w, c = out, ho, wo, cout
            ci += 1
            k += 2
            if fuse_pool:
                if pname in wanted:
                    outs.append(E.native_to_nchw(cur))
                k += 1
        return outs


def _encoder_autograd(self, x, names, last_needed):
    """Training-mode walk (input requires grad): same taps, recorded by train_ops.EncoderFn, which
    back-propagates to the image only -- the VGG weights are a frozen loss network in every flow
    of the reference (train.py:55-56, train_autoencoder.py:24-25 build optimisers over the model
    only), so no weight gradient is produced for them."""
    from .train_ops import EncoderFn
    wanted = self._content_layers
    plan, wb = [], []
    k = 1
    while k <= last_needed:
        conv = self._vgg_layers[k]
        cname, rname = names[k], names[k + 1]
        has_pool = k + 2 < len(names) and names[k + 2].startswith("pool_")
        if has_pool and names[k + 2] in wanted:
            raise L.AstError("pool_i taps are not supported when the encoder input requires grad")
        if cname in wanted and rname in wanted:
            raise L.AstError("tapping both conv_i and relu_i of the same layer is not supported")
        tap = "pre" if cname in wanted else ("post" if rname in wanted else None)
        plan.append((conv.in_channels, conv.out_channels, has_pool, tap))
        wb += [conv.weight, conv.bias]
        k += 3 if has_pool else 2
    return list(EncoderFn.apply(x, plan, *wb))


PretrainedEncoder._forward_autograd = _encoder_autograd


class ClassicDecoder(nn.Sequential):
    """The classic mirrored decoder the reference keeps as a commented ``nn.Sequential``
    (models.py:598-628; channel list conf.py:9): [ReflectionPad2d(1), Conv2d(3x3), ReLU] x 9 (no
    ReLU after the last), nearest x2 upsample after convs 1, 5 and 7.  The module list is built
    exactly as that spec so the state-dict keys ('1.weight', '5.weight', ...) match; ``forward``
    runs the tcgen05 kernels on the native layout instead of calling the sub-modules."""

    def __init__(self, exporting: bool = False):
        mods = []
        for cin, cout, relu, up in E.DECODER_SPEC:
            mods.append(nn.ReflectionPad2d((1, 1, 1, 1)))
            mods.append(nn.Conv2d(cin, cout, (3, 3)))
            if relu:
                mods.append(nn.ReLU())
            if up:
                mods.append(nn.Upsample(scale_factor=2, mode='nearest'))
        super().__init__(*mods)
        self.exporting = exporting
        self._cache = _WeightCache()
        self._buf = E._Buffers()
        self.conv_impl = L.CONV_AUTO
        self.fold = True     # post-upsample convs on the low-res map (engine.run_decoder)

    def _convs(self):
        return [m for m in self if isinstance(m, nn.Conv2d)]

    def forward(self, x):
        """x: (N, 512, h, w) fp32 NCHW features -> (N, 3, 8h, 8w) fp32 image.  With grad enabled and
        trainable parameters the call is recorded for autograd (train_ops.DecoderFn: forward, data
        and weight gradients all on the tensor-core kernels)."""
        L.require_cuda(x)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train_ops import DecoderFn
            params = []
            for c in self._convs():
                params += [c.weight, c.bias]
            return DecoderFn.apply(x, self.exporting, *params)
        t = E.nchw_to_native(x, reflect=True)
        return self.forward_native(t)

    def forward_native(self, t):
        convs = self._convs()
        wpk = [self._cache.packed(convs[i].weight) for i in range(8)] + [None]
        wfold = {i: self._cache.folded(convs[i].weight) for i in E.FOLD_LAYERS} if self.fold else {}
        last = convs[8]
        return E.run_decoder(self._buf, t, wpk, wfold, [c.bias for c in convs], last.weight,
                             self._cache.packed(last.weight, cout_pad=16), self.exporting, None, self.conv_impl,
                             L.CONV_AUTO, self.fold, key="d")


class StyleTransferNet(nn.Module):
    """f = relu4_1(img) ('relu_9'); t = AdaIN(f_c, f_s); alpha blend
    (models.py:471); img = decoder(t).  ``forward(content, style, alpha=1.0)`` like AST.forward
    (models.py:425).  Inference runs end to end in the native layout (StyleTransferEngine)."""

    def __init__(self, encoder: PretrainedEncoder | None = None, decoder: ClassicDecoder | None = None,
                 canonical: bool = False):
        super().__init__()
        self._enc = encoder if encoder is not None else PretrainedEncoder(['relu_9'])
        self._dec = decoder if decoder is not None else ClassicDecoder()
        self.ada_in = AdaIN(canonical)
        self._engine = None
        self._engine_ver = None

    def engine(self) -> E.StyleTransferEngine:
        convs = self._enc._convs()[:9]
        dconvs = self._dec._convs()
        params = [c.weight for c in convs] + [c.bias for c in convs] + \
                 [c.weight for c in dconvs] + [c.bias for c in dconvs]
        ver = tuple((p.data_ptr(), p._version) for p in params)
        if self._engine is None or ver != self._engine_ver:
            dev = convs[0].weight.device
            self._engine = E.StyleTransferEngine([c.weight for c in convs], [c.bias for c in convs],
                                                 [c.weight for c in dconvs], [c.bias for c in dconvs],
                                                 device=dev)
            self._engine_ver = ver
        return self._engine

    @torch.no_grad()
    def forward(self, content_img, style_img, alpha: float = 1.0, style_weights=None):
        return self.engine().stylize(content_img, style_img, alpha=alpha,
                                     style_weights=style_weights, canonical=self.ada_in.canonical)


@torch.no_grad()
def calibrate_encoder_bias(enc: PretrainedEncoder, n_convs: int = 9, size: int = 128,
                           seed: int = 1234) -> None:
    """Synthetic-weight helper (no reference counterpart: the reference downloads pretrained
    weights, models.py:192).  With random-init VGG weights tens of relu4_1 channels are dead and
    the epsilon-free AdaIN (models.py:47) turns them into NaN; this sets, conv by conv,
    ``bias_c = -mean(pre-activation_c)`` on a seeded uniform batch so every channel stays alive
    (SURVEY.md section 8d).  Runs on the GPU through this package's own encoder taps."""
    dev = enc._convs()[0].weight.device
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(2, 3, size, size, generator=g).to(dev)
    saved = enc._content_layers
    try:
        for i, conv in enumerate(enc._convs()[:n_convs]):
            conv.bias.zero_()
            enc._content_layers = {f"conv_{i + 1}"}
            pre = enc(x)[0]
            m, _ = Fn.channel_stats_flat(pre.transpose(0, 1).reshape(1, pre.shape[1], -1, 1))
            conv.bias.copy_(-m.view(-1))
    finally:
        enc._content_layers = saved


# The rest of the reference's ``models`` surface (``from models import *``, train.py:15): the MobileNet-style
# networks (models.py:140-338) and the attention network (models.py:70-115, 393-582) live in their own modules.
from .mobilenet import Encoder, Decoder, DecoderBlock, AutoEncoder  # noqa: E402,F401
from .attention import AdaAttN, AST  # noqa: E402,F401
