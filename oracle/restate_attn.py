"""CPU restatement of the reference's AdaAttN layer and the AST network built on it.  TEST INFRASTRUCTURE ONLY.

SURVEY.md section 8 row f1 (the first "next" row): ``AdaAttN`` (models.py:70-115) and ``AST`` (models.py:393-575).
Functions over a flat state dict with the reference's key names (``ada_att_1.W_q.weight``, ``_enc.mob_net...``,
``_dec._decoder_blocks...``, ``ada_out._layers...``), torch fp32 ATen ops on CPU.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file; the product
package never does.

The reference's ``AST`` is half-deleted as shipped: ``__init__`` never creates
``self.ada_att_2`` / ``self.ada_out`` (commented out at models.py:407, 410) although ``encode`` / ``forward`` /
train.py use both.  The restatement -- like the fixture generator -- restores exactly those two commented lines:
``ada_att_2 = AdaAttN(enc_out_channels)`` and ``ada_out = DepthWiseConv(enc_out_channels*2, enc_out_channels, 1,
EXPAND_RATIO, use_norm=False, use_identity=False)`` (identical to ``AutoEncoder.ada_out``, models.py:326, which
train.py:143 copies into it).

Pinning: the reference has no tests or goldens; ``oracle/make_golden.py`` executes the GENUINE ``AdaAttN`` class and
the genuine ``AST`` (with the two attributes restored and the models.py:459 token fix) on seeded inputs ->
``tests/golden/adaattn.npz``; ``tests/test_oracle_attn_golden.py`` checks this file against them and
``tests/test_oracle_vs_reference.py`` re-checks live while /root/reference is present.  Arithmetic below the
reference (conv2d, instance_norm, softmax, bmm) is third-party torch 2.11.0 ATen.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import restate_ae as A

IN_EPS = 1e-5      # nn.InstanceNorm2d default (models.py:78-80): no affine, no running stats, biased variance


def instance_norm(x, eps: float = IN_EPS):
    """nn.InstanceNorm2d(C) as constructed at models.py:78-80: (x - mean) / sqrt(biased var + eps) per (n, c)."""
    return F.instance_norm(x, eps=eps)      # the ATen op nn.InstanceNorm2d.forward itself calls


def adaattn(P, prefix, content_map, style_map):
    """AdaAttN.forward, models.py:83-115.  Returns ``std * IN(content) + mean`` with the attention-weighted mean
    and standard deviation of V = W_v(style) under softmax(Q K^T), Q = W_q(IN(content)), K = W_k(IN(style))."""
    b, c, h, w = content_map.shape
    q = F.conv2d(instance_norm(content_map), P[prefix + ".W_q.weight"])          # :87
    k = F.conv2d(instance_norm(style_map), P[prefix + ".W_k.weight"])            # :88
    v = F.conv2d(style_map, P[prefix + ".W_v.weight"])                           # :89
    q = q.view(b, c, -1).permute(0, 2, 1)                                        # (b, HW, c)      :92
    k = k.view(b, c, -1)                                                         # (b, c, HWs)     :93
    v = v.view(b, c, -1).permute(0, 2, 1)                                        # (b, HWs, c)     :94
    att = torch.softmax(torch.bmm(q, k), dim=-1)                                 # :97-99
    mean = torch.bmm(att, v)                                                     # :101
    std = torch.sqrt(torch.relu(torch.bmm(att, v ** 2) - mean ** 2))             # :103
    std = std.view(b, -1, w, c).permute(0, 3, 1, 2)                              # :105
    mean = mean.view(b, -1, w, c).permute(0, 3, 1, 2)                            # :106
    return std * instance_norm(content_map) + mean                               # :115


def ada_out(P, x):
    """The restored ``ada_out`` block (models.py:410 / :326)."""
    return A.depthwise_block(P, "ada_out", x, A.ENC_OUT_CHANNELS * 2, A.ENC_OUT_CHANNELS, 1, A.EXPAND_RATIO, 3,
                             norm=False, use_identity=False)


def ast_encode(P, content_img, style_img, detach=False, return_maps=False, training=True):
    """AST.encode, models.py:535-572.  ``detach=True`` runs the encoder in eval mode on both images and detaches
    the taps (:539-547); otherwise the encoder runs in the module's current mode (``training``)."""
    enc_training = False if detach else training
    cm = A.encoder_forward(P, content_img, A.ENC_OUT_LAYERS, training=enc_training, prefix="_enc")
    sm = A.encoder_forward(P, style_img, A.ENC_OUT_LAYERS, training=enc_training, prefix="_enc")
    if detach:
        cm = [t.detach() for t in cm]
        sm = [t.detach() for t in sm]
    s1 = adaattn(P, "ada_att_1", cm[0], sm[0])                                   # :554
    s2 = adaattn(P, "ada_att_2", cm[1], sm[1])                                   # :555
    t = ada_out(P, torch.cat((s1, s2), dim=1))                                   # :565-566
    return (s1, s2, t) if return_maps else t


def ast_forward(P, content_img, style_img, alpha=1.0, exporting=False, training=True):
    """AST.forward, models.py:425-533 (with the :459 token fix).  Not exporting: ``(t_cs, t_return, org_out)``;
    exporting: ``t_cs`` (the decoder then ends in Hardtanh(0, 1), models.py:315-316).

    Note the mode bookkeeping of the reference: ``encode(detach=True)`` leaves ``_enc`` in TRAIN mode (:547), so
    the second encoder pass (:467) uses batch statistics and updates the running statistics whenever the module
    is used the way train.py uses it; ``training=False`` restates ``ast.eval()`` being called AFTER construction
    and the encode call flipping the encoder back to train mode -- i.e. pass ``training`` = the mode ``_enc`` is
    in when :467 executes, which is True for every non-exporting call of the reference."""
    if exporting:
        t = ast_encode(P, content_img, style_img, training=training)
        return A.decoder_forward(P, t, exporting=True, prefix="_dec")
    s1, _, t = ast_encode(P, content_img, style_img, detach=True, return_maps=True)
    t_return = s1
    cm = A.encoder_forward(P, content_img, A.ENC_OUT_LAYERS, training=training, prefix="_enc")   # :467
    content_map = ada_out(P, torch.cat((cm[0], cm[1]), dim=1))                                    # :468-469
    t = alpha * t + (1 - alpha) * content_map                                                     # :471
    org_out = A.decoder_forward(P, content_map, prefix="_dec")                                    # :476
    t_cs = A.decoder_forward(P, t, prefix="_dec")                                                 # :506
    return t_cs, t_return, org_out


def make_ast_state(seed: int = 3):
    """A seeded AST state with the reference's key names: encoder / ada_out / decoder drawn exactly like
    ``restate_ae.make_ae_state(seed)`` (then renamed ``encoder.* -> _enc.*``, ``decoder.* -> _dec.*``, which is what
    train.py:142-144 ``load_ae`` does with a trained autoencoder), plus the two AdaAttN layers' 1x1 convolutions
    with nn.Conv2d's default initialisation drawn after them."""
    sd = A.make_ae_state(seed)
    out = {}
    for k, v in sd.items():
        if k.startswith("encoder."):
            out["_enc." + k[len("encoder."):]] = v
        elif k.startswith("decoder."):
            out["_dec." + k[len("decoder."):]] = v
        else:
            out[k] = v
    c = A.ENC_OUT_CHANNELS
    for name in ("ada_att_1", "ada_att_2"):
        for wn in ("W_q", "W_k", "W_v"):
            conv = torch.nn.Conv2d(c, c, 1, 1, 0, bias=False)
            out[f"{name}.{wn}.weight"] = conv.weight.detach().clone()
    return out


def calibrate_encoder(P, x):
    """Set every BatchNorm of ``_enc`` to the batch statistics of ``x`` (one training-mode pass with momentum 1.0,
    i.e. ``bn.momentum = 1.0`` on the reference modules).  With FRESH running statistics the eval-mode encoder of
    ``AST.encode(detach=True)`` emits taps of ~5e-5, InstanceNorm's eps dominates and the attention is uniform; the
    network fixtures therefore use a calibrated state.  In place; returns P."""
    with torch.no_grad():
        A.encoder_forward(P, x, A.ENC_OUT_LAYERS, training=True, bn_momentum=1.0, prefix="_enc")
    return P


def sharpen_attention(sd, gain: float = 6.0):
    """Fixture helper: default-initialised W_q / W_k give logits of ~1e-1, i.e. a nearly uniform attention in which
    softmax is not exercised; scaling both by ``sqrt(gain)`` spreads the logits to O(gain) so that the fixtures
    pin a peaked attention too.  Returns a new state dict."""
    out = A.clone_state(sd)
    for k in out:
        if k.endswith(".W_q.weight") or k.endswith(".W_k.weight"):
            out[k].mul_(gain ** 0.5)
    return out


# parameters whose full gradient the golden fixture stores (the rest: norms only)
GOLDEN_GRAD_KEYS = (
    "ada_att_1.W_q.weight", "ada_att_1.W_k.weight", "ada_att_1.W_v.weight",
    "ada_att_2.W_q.weight", "ada_att_2.W_v.weight",
    "ada_out._layers.0.weight", "ada_out._layers.5.weight",
    "_dec._img_out.weight", "_dec._decoder_blocks.8._conv._layers.2.weight",
    "_enc.mob_net.14._layers.7.weight", "_enc.mob_net.1._layers.1.weight", "_enc.mob_net.0.0.weight",
)


# ------------------------------------------------------------------------------------------------------
# The STORAGE / PRECISION CONTRACT of the CUDA AdaAttN forward path (csrc/bgemm_tc.cu, attn.cu), restated on CPU:
# fp32 instance norms; W_q, W_k and Q K^T as two-term bf16 splits (three of the four cross terms, fp32 accumulation);
# softmax in fp32, weights rounded to bf16 and the moments divided by the sum of the ROUNDED weights; v = W_v(style)
# from bf16-rounded operands, rounded to bf16; v^2 exact (hi + lo); output rounded to bf16.  Comparing the fp32
# restatement with this one shows what the contract costs; comparing the CUDA path with this one checks the kernels.
# ------------------------------------------------------------------------------------------------------
def _bf(t):
    return t.to(torch.bfloat16).float()


def _split_matmul(a, b):
    """a @ b^T with both operands as hi + lo bf16 splits, dropping lo*lo (what one K = 3C GEMM of [hi|lo|hi] by
    [hi|hi|lo] computes)."""
    ah, bh = _bf(a), _bf(b)
    al, bl = _bf(a - ah), _bf(b - bh)
    return ah @ bh.transpose(-1, -2) + al @ bh.transpose(-1, -2) + ah @ bl.transpose(-1, -2)


def adaattn_contract(P, prefix, content_map, style_map):
    b, c, h, w = content_map.shape
    cn = instance_norm(content_map).flatten(2).transpose(1, 2)                  # (b, HW, c) fp32
    sn = instance_norm(style_map).flatten(2).transpose(1, 2)
    xs = _bf(style_map).flatten(2).transpose(1, 2)
    wq, wk, wv = (P[f"{prefix}.{n}.weight"].view(c, c) for n in ("W_q", "W_k", "W_v"))
    q = _split_matmul(cn, wq.unsqueeze(0))                                      # fp32, ~2^-17
    k = _split_matmul(sn, wk.unsqueeze(0))
    v = _bf(xs @ _bf(wv).t())
    att = _bf(torch.softmax(_split_matmul(q, k), dim=-1))
    lsum = att.sum(-1, keepdim=True)
    mean = (att @ v) / lsum
    m2 = (att @ (v * v)) / lsum                                                 # v^2 = hi + lo is exact for bf16 v
    std = torch.sqrt(torch.relu(m2 - mean * mean))
    out = _bf(std * _bf(cn) + mean)
    return out.transpose(1, 2).reshape(b, c, h, w)
