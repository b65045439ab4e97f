"""CPU restatement of the reference AdaIN hot path.  TEST INFRASTRUCTURE ONLY.

This file is the parity ORACLE: a plain CPU (torch fp32 / numpy fp64) restatement of what
rwickman/ArbitraryStyleTransfer computes on the path BASELINE.json's north_star names.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product package never does: it fails loudly without its CUDA library.

Pinning status: the reference ships NO tests, golden vectors or fixtures, so
this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container
by ``oracle/make_golden.py`` through ``oracle/ref_loader.py`` (the genuine reference classes, read
from /root/reference at run time) and committed under ``tests/golden/``; and, while the reference
tree is present, checked live by ``tests/test_oracle_vs_reference.py``.  The arithmetic below the
reference (conv2d, max_pool2d, mean/std/var, bmm, huber_loss) lives in third-party torch 2.11.0
ATen / torchvision 0.26.0 (unpinned by the reference: it has no requirements file); the
restatement calls the same ATen ops on CPU, and additionally restates the statistics in numpy
fp64 (``*_np``) for small cases.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# a1  channel_stats                                                    model_util.py:3-8
# --------------------------------------------------------------------------------------

def channel_stats(img: torch.Tensor):
    """mean and UNBIASED std over (H, W), keepdim, no epsilon.  model_util.py:3-8."""
    img_mean = img.mean(dim=(2, 3), keepdim=True)
    img_std = img.std(dim=(2, 3), keepdim=True)
    return img_mean, img_std


def channel_stats_np(img: np.ndarray):
    """fp64 numpy restatement of model_util.py:3-8 (small cases)."""
    x = img.astype(np.float64)
    n, c = x.shape[:2]
    x = x.reshape(n, c, -1)
    mean = x.mean(axis=2)
    hw = x.shape[2]
    var = ((x - mean[..., None]) ** 2).sum(axis=2) / (hw - 1) if hw > 1 else np.full_like(mean, np.nan)
    return mean.reshape(n, c, 1, 1), np.sqrt(var).reshape(n, c, 1, 1)


# --------------------------------------------------------------------------------------
# a2  AdaIN.forward                                                      models.py:43-51
# --------------------------------------------------------------------------------------

def adain(content_map: torch.Tensor, style_map: torch.Tensor) -> torch.Tensor:
    """The reference AdaIN, INCLUDING its swapped unpack at models.py:44:
    ``style_std, style_mean = channel_stats(style_map)`` binds style_std := mean(style),
    style_mean := std(style); so out = (c - mu_c)/sigma_c * mu_s + sigma_s."""
    style_std, style_mean = channel_stats(style_map)           # models.py:44 (swapped names)
    content_mean, content_std = channel_stats(content_map)     # models.py:45
    content_map = (content_map - content_mean) / content_std   # models.py:47
    content_map = content_map * style_std + style_mean         # models.py:50
    return content_map


def adain_canonical(content_map, style_map):
    """Huang & Belongie form (sigma_s scale, mu_s shift); NOT what the reference computes."""
    s_mean, s_std = channel_stats(style_map)
    c_mean, c_std = channel_stats(content_map)
    return (content_map - c_mean) / c_std * s_std + s_mean


# --------------------------------------------------------------------------------------
# a3  alpha blend                                                           models.py:471
# --------------------------------------------------------------------------------------

def alpha_blend(t: torch.Tensor, content_map: torch.Tensor, alpha: float) -> torch.Tensor:
    """``t = alpha * t + (1 - alpha) * content_map``  models.py:471."""
    return alpha * t + (1 - alpha) * content_map


def adain_multi(content_map, style_maps, weights, alpha=1.0, canonical=False):
    """K-style interpolation.  NOT in the reference: defined through
    AdaIN's linearity in the style statistics,
        out = z_c * sum_k w_k A_k + sum_k w_k B_k,  z_c = (c - mu_c)/sigma_c,
    with (A, B) = (mu_s, sigma_s) in reference (swapped) mode and (sigma_s, mu_s) canonical,
    followed by the models.py:471 alpha blend.  K = 1, w = (1,) reduces to models.py:43-51."""
    c_mean, c_std = channel_stats(content_map)
    z = (content_map - c_mean) / c_std
    A = torch.zeros_like(c_mean)
    B = torch.zeros_like(c_mean)
    for w, s in zip(weights, style_maps):
        s_mean, s_std = channel_stats(s)
        if canonical:
            A = A + w * s_std
            B = B + w * s_mean
        else:
            A = A + w * s_mean
            B = B + w * s_std
    t = z * A + B
    if alpha != 1.0:
        t = alpha_blend(t, content_map, alpha)
    return t


# --------------------------------------------------------------------------------------
# a4  calc_mean_std / mean_variance_norm                                 models.py:54-68
# --------------------------------------------------------------------------------------

def calc_mean_std(feat: torch.Tensor, eps: float = 1e-5):
    """unbiased var + eps -> sqrt; mean.  models.py:54-62."""
    size = feat.size()
    assert len(size) == 4
    N, C = size[:2]
    feat_var = feat.view(N, C, -1).var(dim=2) + eps
    feat_std = feat_var.sqrt().view(N, C, 1, 1)
    feat_mean = feat.view(N, C, -1).mean(dim=2).view(N, C, 1, 1)
    return feat_mean, feat_std


def mean_variance_norm(feat: torch.Tensor) -> torch.Tensor:
    """(feat - mean) / std with the eps-guarded std.  models.py:64-68."""
    mean, std = calc_mean_std(feat)
    return (feat - mean.expand(feat.size())) / std.expand(feat.size())


# --------------------------------------------------------------------------------------
# a5  PretrainedEncoder (VGG-19 features)                               models.py:186-240
# --------------------------------------------------------------------------------------

# torchvision.models.vgg19 configuration "E" (third-party, torchvision 0.26.0 vgg.py cfgs['E']):
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M",
             512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
IMAGENET_MEAN = (0.485, 0.456, 0.406)   # models.py:189
IMAGENET_STD = (0.229, 0.224, 0.225)    # models.py:190
DEFAULT_TAPS = ("conv_1", "conv_3", "conv_5", "conv_9", "conv_13", "relu_15")  # models.py:187


def vgg_conv_shapes():
    """[(cin, cout)] of the 16 VGG-19 convs in order."""
    shapes, cin = [], 3
    for v in VGG19_CFG:
        if v != "M":
            shapes.append((cin, v))
            cin = v
    return shapes


def vgg_forward(x: torch.Tensor, weights, biases, content_layers=DEFAULT_TAPS):
    """PretrainedEncoder.forward, models.py:230-240, with the layer naming of models.py:198-224:
    Normalization (models.py:129-131) then Conv3x3(zero pad 1)+bias named conv_i, out-of-place
    ReLU named relu_i, MaxPool2x2 named pool_i (i = running conv index); outputs of layers whose
    name is in ``content_layers`` are collected in network order and the walk returns early once
    all are collected (models.py:237-238)."""
    wanted = set(content_layers)
    outs = []
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype).view(-1, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype).view(-1, 1, 1)
    x = (x - mean) / std                                     # models.py:131
    i = 0

    def tap(name, t):
        if name in wanted:
            outs.append(t)
        return len(outs) == len(wanted)

    for v in VGG19_CFG:
        if v == "M":
            x = F.max_pool2d(x, kernel_size=2, stride=2)
            if tap(f"pool_{i}", x):
                return outs
        else:
            x = F.conv2d(x, weights[i], biases[i], stride=1, padding=1)
            i += 1
            if tap(f"conv_{i}", x):
                return outs
            x = F.relu(x)
            if tap(f"relu_{i}", x):
                return outs
    return outs


def vgg_relu4_1(x, weights, biases):
    """relu4_1 == the reference's 'relu_9' (SURVEY.md section 0.5)."""
    return vgg_forward(x, weights, biases, ("relu_9",))[0]


# --------------------------------------------------------------------------------------
# a6  classic mirrored decoder (commented nn.Sequential)                models.py:598-628
# --------------------------------------------------------------------------------------

# (cin, cout, relu_after, upsample_after) for the 9 convs of models.py:598-628
DECODER_SPEC = [
    (512, 256, True, True),
    (256, 256, True, False),
    (256, 256, True, False),
    (256, 256, True, False),
    (256, 128, True, True),
    (128, 128, True, False),
    (128, 64, True, True),
    (64, 64, True, False),
    (64, 3, False, False),
]


def decoder_forward(t: torch.Tensor, weights, biases) -> torch.Tensor:
    """[ReflectionPad2d(1), Conv3x3(p=0, bias), ReLU] x9 (last has no ReLU), nearest x2
    upsample after convs 1, 5, 7.  models.py:598-628."""
    x = t
    for (cin, cout, relu, up), w, b in zip(DECODER_SPEC, weights, biases):
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        x = F.conv2d(x, w, b)
        if relu:
            x = F.relu(x)
        if up:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
    return x


# --------------------------------------------------------------------------------------
# a10 / a11  losses                                                    losses.py:105-139
# --------------------------------------------------------------------------------------

def compute_content_loss(inp, tgt):
    """F.huber_loss, delta 1.0, mean.  losses.py:124-126."""
    return F.huber_loss(inp, tgt)


def huber_np(inp: np.ndarray, tgt: np.ndarray) -> float:
    d = inp.astype(np.float64) - tgt.astype(np.float64)
    a = np.abs(d)
    return float(np.where(a < 1.0, 0.5 * d * d, a - 0.5).mean())


def tv_loss(img):
    """sum (x[..., :-1] - x[..., 1:])^2 + sum (x[..., :-1, :] - x[..., 1:, :])^2, un-normalised.  losses.py:90-103."""
    w_variance = torch.sum(torch.pow(img[:, :, :, :-1] - img[:, :, :, 1:], 2))
    h_variance = torch.sum(torch.pow(img[:, :, :-1, :] - img[:, :, 1:, :], 2))
    return h_variance + w_variance


HIST_K = 256                      # losses.py:42-46 (HistLayerBase): K bins of width L = 1/K over [0, 1], W = L / 2.5
HIST_L = 1.0 / HIST_K
HIST_W = HIST_L / 2.5


def soft_histogram(x, chunk: int = 1 << 14):
    """SingleDimHistLayer.forward, losses.py:49-56 (+ compute_pj / phi_k, :24-37): per image, over ALL C*H*W
    elements, hist_k = sum_i [sigmoid((x_i - mu_k + L/2)/W) - sigmoid((x_i - mu_k - L/2)/W)] / N with
    N = x.size(1) * x.size(2) = C*H (the reference's normaliser, not C*H*W).  Same arithmetic as the reference, but
    the (B, 256, C*H*W) tensor it materialises is walked in chunks of elements."""
    B = x.size(0)
    N = x.size(1) * x.size(2)
    mu_k = (HIST_L * (torch.arange(HIST_K, dtype=x.dtype) + 0.5)).view(1, -1, 1)
    flat = x.reshape(B, 1, -1)
    hist = torch.zeros(B, HIST_K, dtype=x.dtype)
    for i in range(0, flat.size(2), chunk):
        d = flat[:, :, i:i + chunk] - mu_k
        hist = hist + (torch.sigmoid((d + HIST_L / 2) / HIST_W) - torch.sigmoid((d - HIST_L / 2) / HIST_W)).sum(dim=2)
    return hist / N


def compute_hist_loss(t_cs, style_map):
    """losses.py:82-87: squared earth mover's distance between the two soft histograms -- sum over bins of the
    squared difference of their cumulative sums (EarthMoversDistanceLoss, :8-22: matmul with the upper-triangular
    ones matrix) -- averaged over the batch."""
    hx, hy = soft_histogram(t_cs), soft_histogram(style_map)
    tt = torch.triu(torch.ones(HIST_K, HIST_K, dtype=hx.dtype))          # tt[s][t] = (t >= s)
    return torch.sum(torch.square(hx @ tt - hy @ tt), dim=1).mean()


def gram_matrix(tensor):
    """X X^T / (C*H*W), X = (B, C, HW).  losses.py:105-109."""
    B, C, H, W = tensor.shape
    x = tensor.view(B, C, H * W)
    x_t = x.transpose(1, 2)
    return torch.bmm(x, x_t) / (C * H * W)


def compute_style_loss(t_cs_map, style_map):
    """1.25 huber(mean) + 1.25 huber(std) + 10 huber(gram).  losses.py:128-139."""
    enc_mean, enc_std = channel_stats(t_cs_map)
    style_mean, style_std = channel_stats(style_map)
    mean_loss = F.huber_loss(enc_mean, style_mean) * 1.25
    std_loss = F.huber_loss(enc_std, style_std) * 1.25
    g_c = gram_matrix(t_cs_map)
    g_s = gram_matrix(style_map)
    gram_loss = F.huber_loss(g_c, g_s) * 10
    return mean_loss + std_loss + gram_loss


# --------------------------------------------------------------------------------------
# Full classic path: configs 1, 4, 5
# --------------------------------------------------------------------------------------

def stylize(content_img, style_imgs, vgg_w, vgg_b, dec_w, dec_b, alpha=1.0,
            style_weights=None, canonical=False):
    """f_c = relu4_1(content); f_s = relu4_1(style_k); t = AdaIN(+K-style mix, +alpha blend);
    img = decoder(t).  ``style_imgs`` is a tensor (N,3,H,W) (K = 1) or a list of K tensors."""
    if isinstance(style_imgs, torch.Tensor):
        style_imgs = [style_imgs]
    if style_weights is None:
        style_weights = [1.0 / len(style_imgs)] * len(style_imgs)
    f_c = vgg_relu4_1(content_img, vgg_w, vgg_b)
    f_s = [vgg_relu4_1(s, vgg_w, vgg_b) for s in style_imgs]
    if len(f_s) == 1 and not canonical and style_weights[0] == 1.0:
        t = adain(f_c, f_s[0])
        if alpha != 1.0:
            t = alpha_blend(t, f_c, alpha)
    else:
        t = adain_multi(f_c, f_s, style_weights, alpha, canonical)
    return decoder_forward(t, dec_w, dec_b)


# --------------------------------------------------------------------------------------
# Synthetic weight recipe.  The reference hard-codes pretrained=True
# (models.py:192) which needs a network; there are no checkpoints offline, so parity runs on
# seeded random-init weights with a bias calibration that keeps every relu4_1 channel alive
# (the reference AdaIN has no epsilon, models.py:47: a dead channel gives 0/0 = NaN).
# --------------------------------------------------------------------------------------

def make_vgg_weights(seed: int = 0):
    """He-normal (fan_out, relu) conv weights like torchvision's VGG._initialize_weights,
    drawn from an explicit seeded generator, conv by conv in network order; zero biases."""
    g = torch.Generator().manual_seed(seed)
    ws, bs = [], []
    for cin, cout in vgg_conv_shapes():
        std = (2.0 / (cout * 9)) ** 0.5
        ws.append(torch.randn(cout, cin, 3, 3, generator=g) * std)
        bs.append(torch.zeros(cout))
    return ws, bs


def calibrate_vgg_bias(ws, calib_seed: int = 1234, size: int = 128):
    """bias_c <- -mean_{n,h,w}(pre-activation_c), conv by conv, on a 2x3xSxS uniform batch."""
    g = torch.Generator().manual_seed(calib_seed)
    x = torch.rand(2, 3, size, size, generator=g)
    mean = torch.tensor(IMAGENET_MEAN).view(-1, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(-1, 1, 1)
    x = (x - mean) / std
    bs, i = [], 0
    with torch.no_grad():
        for v in VGG19_CFG:
            if v == "M":
                x = F.max_pool2d(x, 2, 2)
            else:
                pre = F.conv2d(x, ws[i], None, padding=1)
                b = -pre.mean(dim=(0, 2, 3))
                bs.append(b)
                x = F.relu(pre + b.view(1, -1, 1, 1))
                i += 1
    return bs


def make_decoder_weights(seed: int = 1):
    """He-normal (fan_in) weights and small uniform biases for the 9 decoder convs."""
    g = torch.Generator().manual_seed(seed)
    ws, bs = [], []
    for cin, cout, _, _ in DECODER_SPEC:
        std = (2.0 / (cin * 9)) ** 0.5
        ws.append(torch.randn(cout, cin, 3, 3, generator=g) * std)
        bs.append((torch.rand(cout, generator=g) - 0.5) * 0.2)
    return ws, bs


def rand_image(n, size, seed, w=None):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, size, w or size, generator=g)


def psnr(a: torch.Tensor, ref: torch.Tensor) -> float:
    """PSNR with range = max(ref) - min(ref) (random-init outputs are not in [0,1])."""
    mse = torch.mean((a.double() - ref.double()) ** 2).item()
    rng = (ref.max() - ref.min()).item()
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(rng * rng / mse)
