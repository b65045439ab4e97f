"""CPU restatement of the reference's MobileNet-style autoencoder path.  TEST INFRASTRUCTURE ONLY.

Rows a7-a9 of SURVEY.md section 8: ``Encoder`` (models.py:140-184), ``DecoderBlock`` / ``Decoder``
(models.py:242-320), ``AutoEncoder`` (models.py:322-338), their building blocks ``conv_3x3_bn`` /
``SELayer`` / ``DepthWiseConv`` (mobilenetv2.py:38-43, 63-81, 95-181) and the training step of
train_autoencoder.py:111-148.  Written as FUNCTIONS over a flat state dict with the reference's
key names (``encoder.mob_net.1._layers.3.weight`` ...), torch fp32 ATen ops on CPU.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
file; the product package never does.

Pinning: the reference has no tests or goldens, so this restatement is pinned against the GENUINE
reference classes executed by ``oracle/make_golden.py`` (fixtures ``tests/golden/autoencoder.npz``,
checked by ``tests/test_oracle_ae_golden.py``; live re-check in ``tests/test_oracle_vs_reference.py``
while /root/reference is present).  Arithmetic below the reference (conv2d, batch_norm, hardswish,
adaptive_avg_pool2d, linear, hardtanh, huber_loss) is third-party torch 2.11.0 ATen.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import restate as R

# ---- topology constants: conf.py:71-113 ---------------------------------------------------------
EXPAND_RATIO = 3
ENC_SHAPES = [(3, 16, 1, 3, 1), (16, 16, 1, 3, 6), (16, 24, 2, 3, 6), (24, 24, 1, 3, 6), (24, 40, 2, 5, 6),
              (40, 40, 1, 5, 4), (40, 40, 1, 5, 4), (40, 80, 2, 3, 4), (80, 80, 1, 3, 4), (80, 80, 1, 3, 4),
              (80, 96, 1, 5, 4), (96, 96, 1, 5, 3), (96, 128, 1, 3, 3), (128, 128, 1, 3, 3),
              (128, 128, 1, 3, 3)]
DEC_SHAPES = [(128, 128, 1, 3, 3), (128, 128, 1, 3, 3), (128, 96, 1, 3, 3), (96, 96, 1, 5, 3),
              (96, 80, 1, 5, 4), (80, 80, 1, 3, 4), (80, 80, 1, 3, 4), (80, 40, 1, 3, 4), (40, 40, 1, 5, 4),
              (40, 40, 1, 5, 4), (40, 24, 1, 5, 6), (24, 24, 1, 3, 6), (24, 16, 1, 3, 6), (16, 16, 1, 3, 6),
              (16, 3, 1)]
ENC_OUT_LAYERS = (12, 14)
ENC_OUT_CHANNELS = 128
BN_EPS, BN_MOMENTUM = 1e-5, 0.1          # nn.BatchNorm2d defaults (mobilenetv2.py:108 etc.)


def make_divisible(v, divisor, min_value=None):
    """mobilenetv2.py:16-35."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def encoder_block_specs():
    """[(inp, oup, stride, t, k)] for Encoder.mob_net[1:], models.py:148-155.  The last block re-uses
    the loop variables of the final iteration (in_ch, out_ch, stride) with EXPAND_RATIO and k = 3."""
    specs = [(i, o, s, t, k) for (i, o, s, k, t) in ENC_SHAPES[1:-1]]
    i, o, s = ENC_SHAPES[-2][:3]
    specs.append((i, o, s, EXPAND_RATIO, 3))
    return specs


def decoder_block_specs():
    """[(inp, oup, stride, t, k, upsample)] for Decoder._decoder_blocks, models.py:280-292."""
    out = []
    for idx, (i, o, s, k, t) in enumerate(DEC_SHAPES[:-1]):
        out.append((i, o, s, t, k, (i != o and idx + 6 < len(DEC_SHAPES))))
    return out


# ---- one DepthWiseConv block: mobilenetv2.py:95-165 ----------------------------------------------
def _layer_index(expand: bool, norm: bool):
    """positions inside ``_layers`` (mobilenetv2.py:103-150) -> dict of indices."""
    idx, i = {}, 0
    if expand:
        idx["pw1"] = i; i += 1
        if norm:
            idx["bn1"] = i; i += 1
        i += 1                                  # Hardswish
    else:
        i += 1                                  # ReflectionPad2d
    idx["dw"] = i; i += 1
    if norm:
        idx["bn2"] = i; i += 1
    i += 1                                      # Hardswish
    idx["se"] = i; i += 1
    idx["pw2"] = i; i += 1
    if norm:
        idx["bn3"] = i; i += 1
    return idx


def _bn(P, key, x, training, momentum=BN_MOMENTUM):
    """nn.BatchNorm2d(affine, track_running_stats): batch statistics + in-place running-stat update in
    training mode, running statistics in eval mode."""
    rm, rv = P[key + ".running_mean"], P[key + ".running_var"]
    y = F.batch_norm(x, rm, rv, P[key + ".weight"], P[key + ".bias"], training, momentum, BN_EPS)
    if training and (key + ".num_batches_tracked") in P:
        P[key + ".num_batches_tracked"] += 1
    return y


def se_layer(P, key, x):
    """SELayer.forward, mobilenetv2.py:73-81: x * Hardtanh(0,1)(W2 relu(W1 avgpool(x) + b1) + b2)."""
    b, c = x.shape[:2]
    y = F.adaptive_avg_pool2d(x, 1).view(b, c)
    y = F.relu(F.linear(y, P[key + ".fc.0.weight"], P[key + ".fc.0.bias"]))
    y = F.hardtanh(F.linear(y, P[key + ".fc.2.weight"], P[key + ".fc.2.bias"]), 0.0, 1.0)
    return x * y.view(b, c, 1, 1)


def depthwise_block(P, prefix, x, inp, oup, stride, t, k=3, norm=False, use_identity=True, training=False,
                    bn_momentum=BN_MOMENTUM):
    """DepthWiseConv.forward (mobilenetv2.py:152-165) for the block whose parameters live under
    ``prefix + '._layers.'``."""
    expand = t != 1
    hidden = round(inp * t)
    ix = _layer_index(expand, norm)
    L = prefix + "._layers."
    org = x
    if expand:
        x = F.conv2d(x, P[f"{L}{ix['pw1']}.weight"])
        if norm:
            x = _bn(P, f"{L}{ix['bn1']}", x, training, bn_momentum)
        x = F.hardswish(x)
        pad = (k - 1) // 2                                                  # mobilenetv2.py:133
    else:
        pad = 1                                                             # mobilenetv2.py:105 (always 1)
    x = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    x = F.conv2d(x, P[f"{L}{ix['dw']}.weight"], stride=stride, groups=hidden)
    if norm:
        x = _bn(P, f"{L}{ix['bn2']}", x, training, bn_momentum)
    x = F.hardswish(x)
    x = se_layer(P, f"{L}{ix['se']}", x)
    x = F.conv2d(x, P[f"{L}{ix['pw2']}.weight"])
    if norm:
        x = _bn(P, f"{L}{ix['bn3']}", x, training, bn_momentum)
    if stride == 1 and inp == oup and use_identity:                         # mobilenetv2.py:99, 161
        x = x + org
    return x


# ---- Encoder / Decoder / AutoEncoder ---------------------------------------------------------------
def encoder_forward(P, x, out_layers=(), auto_enc=False, training=False, prefix="encoder", bn_momentum=BN_MOMENTUM):
    """Encoder.forward, models.py:158-184.  Block 0 = reflect-padded 3x3 conv, no bias, Hardswish
    (conv_3x3_bn, mobilenetv2.py:38-43: no BatchNorm despite the name)."""
    outs = []
    x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), P[f"{prefix}.mob_net.0.0.weight"])
    x = F.hardswish(x)
    if 0 in out_layers:
        outs.append(x)
    for i, (inp, oup, s, t, k) in enumerate(encoder_block_specs(), start=1):
        x = depthwise_block(P, f"{prefix}.mob_net.{i}", x, inp, oup, s, t, k, norm=True, training=training,
                            bn_momentum=bn_momentum)
        if i in out_layers:
            outs.append(x)
    return x if auto_enc else outs


def decoder_forward(P, x, exporting=False, prefix="decoder"):
    """Decoder.forward, models.py:306-320 (+ DecoderBlock.forward, models.py:256-272)."""
    for i, (inp, oup, s, t, k, up) in enumerate(decoder_block_specs()):
        b = f"{prefix}._decoder_blocks.{i}"
        x = depthwise_block(P, b + "._conv", x, inp, oup, s, t, k, norm=False)
        if up:
            x = F.interpolate(x, scale_factor=2, mode="nearest")                       # _upsample_3
            x = depthwise_block(P, b + "._upsample_2", x, oup, oup, 1, 1, 3, norm=False)
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")                                          # _ref_out
    x = F.conv2d(x, P[f"{prefix}._img_out.weight"], P[f"{prefix}._img_out.bias"])
    if exporting:
        x = F.hardtanh(x, 0.0, 1.0)                                                     # models.py:315-316
    return x


def autoencoder_forward(P, x, training=False, bn_momentum=BN_MOMENTUM):
    """AutoEncoder.forward, models.py:329-338."""
    e = encoder_forward(P, x, ENC_OUT_LAYERS, training=training, bn_momentum=bn_momentum)
    z = depthwise_block(P, "ada_out", torch.cat((e[0], e[1]), dim=1), ENC_OUT_CHANNELS * 2, ENC_OUT_CHANNELS,
                        1, EXPAND_RATIO, 3, norm=False, use_identity=False)
    return decoder_forward(P, z)


def ae_losses(P, x, vgg_w, vgg_b, recon_lam=100.0, perp_lam=0.01, training=True):
    """Loss of one train_autoencoder.py step (:111-139): Huber(recon, x) and the sum of Huber losses
    between the default PretrainedEncoder taps of recon and x.  Returns (loss, recon_loss, perp, recon)."""
    recon = autoencoder_forward(P, x, training=training)
    recon_loss = F.huber_loss(recon, x)                                     # nn.HuberLoss(), :28, :113
    with torch.no_grad():
        cm = R.vgg_forward(x, vgg_w, vgg_b)
    rm = R.vgg_forward(recon, vgg_w, vgg_b)
    perp = None
    for a, b in zip(rm, cm):
        l = R.compute_content_loss(a, b.detach())
        perp = l if perp is None else perp + l
    return recon_lam * recon_loss + perp_lam * perp, recon_loss, perp, recon


# ---- seeded synthetic weights: same RNG draws, in the same order, as ``AutoEncoder()`` -------------
def _init_block(sd, prefix, inp, oup, t, k, stride, norm):
    """Replays DepthWiseConv.__init__ + _initialize_weights (mobilenetv2.py:96-181): every layer is
    first constructed with torch's default init (which consumes RNG), then re-drawn in module order."""
    expand = t != 1
    hidden = round(inp * t)
    ix = _layer_index(expand, norm)
    mods = {}
    if expand:
        mods["pw1"] = nn.Conv2d(inp, hidden, 1, bias=False)
        if norm:
            mods["bn1"] = nn.BatchNorm2d(hidden)
    mods["dw"] = nn.Conv2d(hidden, hidden, k, stride, 0, groups=hidden, bias=False)
    if norm:
        mods["bn2"] = nn.BatchNorm2d(hidden)
    sq = make_divisible(hidden // 4, 8)
    mods["se.fc.0"] = nn.Linear(hidden, sq)
    mods["se.fc.2"] = nn.Linear(sq, hidden)
    mods["pw2"] = nn.Conv2d(hidden, oup, 1, bias=False)
    if norm:
        mods["bn3"] = nn.BatchNorm2d(oup)
    for name, m in mods.items():                      # construction order == modules() order
        if isinstance(m, nn.Conv2d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2. / n))
        elif isinstance(m, nn.Linear):
            m.weight.data.normal_(0, 0.01)
            m.bias.data.zero_()
    for name, m in mods.items():
        if name.startswith("se."):
            key = f"{prefix}._layers.{ix['se']}.{name[3:]}"
        else:
            key = f"{prefix}._layers.{ix[name]}"
        for pn, p in list(m.named_parameters()) + list(m.named_buffers()):
            sd[f"{key}.{pn}"] = p.detach().clone()


def make_ae_state(seed: int = 2):
    """State dict of ``AutoEncoder()`` constructed under ``torch.manual_seed(seed)`` (SURVEY.md 8d,
    config 3), reproduced draw for draw without the reference classes."""
    torch.manual_seed(seed)
    sd = {}
    stem = nn.Conv2d(3, 16, 3, 1, 1, bias=False, padding_mode="reflect")   # default init only
    sd["encoder.mob_net.0.0.weight"] = stem.weight.detach().clone()
    for i, (inp, oup, s, t, k) in enumerate(encoder_block_specs(), start=1):
        _init_block(sd, f"encoder.mob_net.{i}", inp, oup, t, k, s, True)
    _init_block(sd, "ada_out", 256, 128, EXPAND_RATIO, 3, 1, False)
    for i, (inp, oup, s, t, k, up) in enumerate(decoder_block_specs()):
        _init_block(sd, f"decoder._decoder_blocks.{i}._conv", inp, oup, t, k, s, False)
        if up:
            _init_block(sd, f"decoder._decoder_blocks.{i}._upsample_2", oup, oup, 1, 3, 1, False)
    head = nn.Conv2d(16, 3, (3, 3))
    sd["decoder._img_out.weight"] = head.weight.detach().clone()
    sd["decoder._img_out.bias"] = head.bias.detach().clone()
    return sd


def clone_state(sd, requires_grad=False):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


# parameters / buffers whose full gradient / value the golden fixture stores (the rest: norms only)
GOLDEN_GRAD_KEYS = (
    "encoder.mob_net.0.0.weight",
    "encoder.mob_net.1._layers.0.weight", "encoder.mob_net.1._layers.1.weight", "encoder.mob_net.1._layers.1.bias",
    "encoder.mob_net.4._layers.3.weight", "encoder.mob_net.4._layers.4.weight",
    "encoder.mob_net.7._layers.6.fc.0.weight", "encoder.mob_net.7._layers.6.fc.2.bias",
    "encoder.mob_net.14._layers.7.weight", "encoder.mob_net.14._layers.8.bias",
    "ada_out._layers.0.weight", "ada_out._layers.2.weight",
    "decoder._decoder_blocks.2._upsample_2._layers.1.weight", "decoder._decoder_blocks.2._upsample_2._layers.4.weight",
    "decoder._decoder_blocks.8._conv._layers.2.weight", "decoder._decoder_blocks.13._conv._layers.5.weight",
    "decoder._decoder_blocks.13._conv._layers.4.fc.0.bias",
    "decoder._img_out.weight", "decoder._img_out.bias",
)
GOLDEN_BUFFER_KEYS = (
    "encoder.mob_net.1._layers.1.running_mean", "encoder.mob_net.1._layers.1.running_var",
    "encoder.mob_net.1._layers.1.num_batches_tracked",
    "encoder.mob_net.7._layers.4.running_mean", "encoder.mob_net.7._layers.4.running_var",
    "encoder.mob_net.14._layers.8.running_mean", "encoder.mob_net.14._layers.8.running_var",
)


def activate_gates(sd, gate_bias: float = 0.5, hidden_bias: float = 0.25):
    """A NON-DEGENERATE variant of a seeded state, for parity fixtures.

    The reference's initialisation is degenerate for this network: (1) SELayer Linear weights N(0, 0.01) with
    zero biases (mobilenetv2.py:178-181) leave every gate at Hardtanh(~1e-4); (2) depthwise weights are drawn
    with std sqrt(2 / (k*k*C)) (mobilenetv2.py:171-172), a gain of sqrt(2/C) ~ 0.07 per depthwise conv, and the
    decoder has no BatchNorm to undo it.  A freshly constructed AutoEncoder therefore outputs EXACTLY its head
    bias in fp32 and most gradients are exactly zero: that state pins names, shapes, the encoder and the first
    blocks, but exercises nothing downstream.  The fixtures are therefore ALSO made on this variant of the same
    seeded state: SE biases moved into the gates' linear region (fc.2.bias = 0.5, fc.0.bias = 0.25) and every
    depthwise weight rescaled to unit gain (x sqrt(C/2); the encoder's too, so that eval mode with fresh running
    statistics -- BatchNorm ~ identity -- keeps a signal as well).
    Returns a new state dict."""
    out = clone_state(sd)
    for k in out:
        if k.endswith(".fc.2.bias"):
            out[k].fill_(gate_bias)
        elif k.endswith(".fc.0.bias"):
            out[k].fill_(hidden_bias)
        elif out[k].dim() == 4 and out[k].shape[1] == 1 and out[k].shape[2] > 1:      # every depthwise weight
            out[k].mul_(math.sqrt(out[k].shape[0] / 2.0))
    return out


def calibrate_running_stats(P, x):
    """Set every BatchNorm's running statistics to the batch statistics of ``x`` (one training-mode forward with
    momentum 1.0, i.e. ``bn.momentum = 1.0`` on the reference modules), so that the eval-mode forward that
    follows normalises a non-degenerate signal.  In place; returns P."""
    with torch.no_grad():
        autoencoder_forward(P, x, training=True, bn_momentum=1.0)
    return P


# ------------------------------------------------------------------------------------------------------
# The 16-bit STORAGE CONTRACT of the CUDA inference path, restated on CPU.
# Same mathematics as depthwise_block(..., training=False), but every tensor the kernels keep in HBM is rounded
# to the storage format -- IEEE fp16 since round 2 (fmt="fp16"; "bf16" restates the round-1 kernels and shows what the
# 8-bit significand cost: 5 % at the deepest encoder tap at 256 x 256, 0.6 % with fp16) -- at exactly the point where
# they round it: the folded pointwise weights, the Hardswish'ed expand
# output, the Hardswish'ed depthwise output, the SE-scaled per-sample pointwise weights and the block output
# (after bias and residual).  Accumulation, biases, BatchNorm folding, SE and the depthwise weights stay fp32.
# Comparing the fp32 restatement with this one shows what bf16 storage costs on a given state (it is what the
# CUDA path is allowed to differ by); comparing the CUDA path with this one checks the kernels themselves.
# ------------------------------------------------------------------------------------------------------
_FMT = {"fp16": torch.float16, "bf16": torch.bfloat16}
_fmt = ["fp16"]


def _bf(t):
    """round-trip through the storage format of the contract being restated (fp16 unless set otherwise)"""
    return t.to(_FMT[_fmt[0]]).float()


def _fold_bn_eval(P, conv_key, bn_key):
    w = P[conv_key + ".weight"]
    if bn_key is None:
        return w, None
    inv = (P[bn_key + ".running_var"] + BN_EPS).rsqrt() * P[bn_key + ".weight"]
    return w * inv.view(-1, 1, 1, 1), P[bn_key + ".bias"] - P[bn_key + ".running_mean"] * inv


def depthwise_block_bf16(P, prefix, x, inp, oup, stride, t, k=3, norm=False, use_identity=True, up2=False):
    """x: bf16-representable (N,inp,H,W) -> bf16-representable block output, eval mode."""
    expand = t != 1
    hidden = round(inp * t)
    ix = _layer_index(expand, norm)
    L = prefix + "._layers."
    h = x
    if expand:
        w1, b1 = _fold_bn_eval(P, f"{L}{ix['pw1']}", f"{L}{ix['bn1']}" if norm else None)
        a = F.conv2d(x, _bf(w1))
        if b1 is not None:
            a = a + b1.view(1, -1, 1, 1)
        h = _bf(F.hardswish(a))
        pad = (k - 1) // 2
    else:
        pad = 1
    if up2:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
    wd, bd = _fold_bn_eval(P, f"{L}{ix['dw']}", f"{L}{ix['bn2']}" if norm else None)
    y = F.conv2d(F.pad(h, (pad, pad, pad, pad), mode="reflect"), wd, stride=stride, groups=hidden)
    if bd is not None:
        y = y + bd.view(1, -1, 1, 1)
    y = _bf(F.hardswish(y))
    se = f"{L}{ix['se']}"
    g = F.hardtanh(F.linear(F.relu(F.linear(y.mean(dim=(2, 3)), P[se + ".fc.0.weight"], P[se + ".fc.0.bias"])),
                            P[se + ".fc.2.weight"], P[se + ".fc.2.bias"]), 0.0, 1.0)
    w2, b2 = _fold_bn_eval(P, f"{L}{ix['pw2']}", f"{L}{ix['bn3']}" if norm else None)
    outs = []
    for n in range(x.shape[0]):
        w2s = _bf(w2.view(oup, hidden) * g[n].view(1, -1)).view(oup, hidden, 1, 1)
        outs.append(F.conv2d(y[n:n + 1], w2s))
    o = torch.cat(outs)
    if b2 is not None:
        o = o + b2.view(1, -1, 1, 1)
    if stride == 1 and inp == oup and use_identity:
        o = o + (F.interpolate(x, scale_factor=2, mode="nearest") if up2 else x)
    return _bf(o)


def autoencoder_forward_contract(P, x, want=(), fmt="fp16"):
    """AutoEncoder.forward in eval mode under the 16-bit storage contract of the CUDA path (``fmt``: "fp16" = the
    kernels as they are, "bf16" = the round-1 kernels).  Returns (image, {name: tensor}) with the intermediate
    tensors named in ``want`` ('enc<i>', 'code')."""
    _fmt[0] = fmt
    try:
        return _autoencoder_forward_contract(P, x, want)
    finally:
        _fmt[0] = "fp16"


def _autoencoder_forward_contract(P, x, want=()):
    keep = {}
    h = _bf(F.hardswish(F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), P["encoder.mob_net.0.0.weight"])))
    if "enc0" in want:
        keep["enc0"] = h
    taps = {}
    for i, (inp, oup, s, t, k) in enumerate(encoder_block_specs(), start=1):
        h = depthwise_block_bf16(P, f"encoder.mob_net.{i}", h, inp, oup, s, t, k, norm=True)
        if f"enc{i}" in want:
            keep[f"enc{i}"] = h
        if i in ENC_OUT_LAYERS:
            taps[i] = h
    z = depthwise_block_bf16(P, "ada_out", torch.cat((taps[12], taps[14]), dim=1), 256, 128, 1, EXPAND_RATIO, 3,
                             norm=False, use_identity=False)
    if "code" in want:
        keep["code"] = z
    h = z
    for i, (inp, oup, s, t, k, up) in enumerate(decoder_block_specs()):
        b = f"decoder._decoder_blocks.{i}"
        h = depthwise_block_bf16(P, b + "._conv", h, inp, oup, s, t, k, norm=False)
        if up:
            h = depthwise_block_bf16(P, b + "._upsample_2", h, oup, oup, 1, 1, 3, norm=False, up2=True)
    img = F.conv2d(F.pad(h, (1, 1, 1, 1), mode="reflect"), P["decoder._img_out.weight"], P["decoder._img_out.bias"])
    return img, keep
