"""Generate tests/golden/*.npz by EXECUTING THE GENUINE REFERENCE CODE (read from /root/reference
at run time through oracle/ref_loader.py) on seeded inputs.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the GPU box has no /root/reference):
    python -m oracle.make_golden
The fixtures pin oracle/restate.py (tests/test_oracle_golden.py) and, through it, the CUDA path
(tests/test_gpu_*.py).  Every file records the torch / torchvision versions that produced it,
because the arithmetic below the reference lives in ATen / torchvision (unpinned by the
reference, which has no requirements file).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader, restate as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _versions():
    import torchvision
    return {"torch_version": np.array(torch.__version__), "torchvision_version": np.array(torchvision.__version__)}


def _t(seed, *shape, scale=1.0, shift=0.0, relu=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g) * scale + shift
    return torch.relu(x) if relu else x


def golden_stats_adain(M, mu, out):
    """AdaIN / channel_stats / calc_mean_std / mean_variance_norm from the genuine reference."""
    cases = {
        "a": ((2, 8, 7, 9), (2, 8, 5, 6)),        # ragged, HW not a multiple of 4
        "b": ((1, 16, 32, 32), (1, 16, 32, 32)),  # cfg-1 spatial size, few channels
        "c": ((3, 4, 1, 2), (3, 4, 2, 1)),        # HW = 2: smallest size with a finite unbiased std
        "d": ((1, 6, 64, 48), (1, 6, 16, 80)),    # different content / style sizes
    }
    for k, (cs, ss) in cases.items():
        c = _t(10 + ord(k), *cs, scale=3.0, shift=1.0, relu=True)
        s = _t(20 + ord(k), *ss, scale=2.0, shift=3.0)
        out[f"adain_{k}_content"] = c.numpy()
        out[f"adain_{k}_style"] = s.numpy()
        with torch.no_grad():
            out[f"adain_{k}_out"] = M.AdaIN()(c, s).numpy()                  # models.py:43-51
            m, sd = mu.channel_stats(c)                                      # model_util.py:3-8
            out[f"adain_{k}_cmean"], out[f"adain_{k}_cstd"] = m.numpy(), sd.numpy()
            m, sd = M.calc_mean_std(c)                                       # models.py:54-62
            out[f"adain_{k}_cms_mean"], out[f"adain_{k}_cms_std"] = m.numpy(), sd.numpy()
            out[f"adain_{k}_mvn"] = M.mean_variance_norm(c).numpy()          # models.py:64-68
            t = M.AdaIN()(c, s)
            out[f"adain_{k}_blend06"] = (0.6 * t + (1 - 0.6) * c).numpy()     # models.py:471
    # a dead (constant) channel: the reference divides 0/0 (no epsilon) -> NaN
    c = _t(31, 1, 3, 4, 4)
    c[:, 1] = 0.0
    s = _t(32, 1, 3, 4, 4)
    with torch.no_grad():
        out["adain_dead_content"], out["adain_dead_style"] = c.numpy(), s.numpy()
        out["adain_dead_out"] = M.AdaIN()(c, s).numpy()
    # backward of mean_variance_norm and channel_stats (autograd of the reference functions)
    c = _t(41, 2, 5, 6, 7, scale=2.0, shift=0.5).requires_grad_(True)
    gy = _t(42, 2, 5, 6, 7)
    M.mean_variance_norm(c).backward(gy)
    out["mvn_bwd_x"], out["mvn_bwd_gy"], out["mvn_bwd_gx"] = c.detach().numpy(), gy.numpy(), c.grad.numpy()
    c2 = c.detach().clone().requires_grad_(True)
    m, sd = mu.channel_stats(c2)
    gm, gs = _t(43, 2, 5, 1, 1), _t(44, 2, 5, 1, 1)
    (m * gm).sum().add((sd * gs).sum()).backward()
    out["cs_bwd_gm"], out["cs_bwd_gs"], out["cs_bwd_gx"] = gm.numpy(), gs.numpy(), c2.grad.numpy()


def golden_losses(Ls, out):
    """compute_content_loss / gram_matrix / compute_style_loss values and input gradients."""
    a = _t(51, 2, 6, 9, 11, scale=1.5).requires_grad_(True)
    b = _t(52, 2, 6, 9, 11, scale=1.5, shift=0.3)
    out["loss_a"], out["loss_b"] = a.detach().numpy(), b.numpy()
    l = Ls.compute_content_loss(a, b)                                     # losses.py:124-126
    l.backward()
    out["content_loss"], out["content_loss_ga"] = l.detach().numpy(), a.grad.numpy()
    a.grad = None
    g = Ls.gram_matrix(a)                                                 # losses.py:105-109
    out["gram_a"] = g.detach().numpy()
    gg = _t(53, 2, 6, 6)
    (g * gg).sum().backward()
    out["gram_gg"], out["gram_ga"] = gg.numpy(), a.grad.numpy()
    a.grad = None
    l = Ls.compute_style_loss(a, b)                                       # losses.py:128-139
    l.backward()
    out["style_loss"], out["style_loss_ga"] = l.detach().numpy(), a.grad.numpy()
    a.grad = None
    l = Ls.tv_loss(a)                                                     # losses.py:90-103
    (l * 0.37).backward()
    out["tv_loss"], out["tv_loss_ga"] = l.detach().numpy(), a.grad.numpy()


def golden_networks(M, out):
    """Genuine PretrainedEncoder (models.py:186-240) and the commented classic decoder
    (models.py:598-628) with the seeded synthetic weight recipe of oracle/restate.py."""
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)

    def load_vgg(enc):
        convs = [l for l in enc._vgg_layers if isinstance(l, torch.nn.Conv2d)]
        assert len(convs) == 16
        with torch.no_grad():
            for c, w, b in zip(convs, vw, vb):
                c.weight.copy_(w)
                c.bias.copy_(b)

    enc_default = M.PretrainedEncoder().eval()
    load_vgg(enc_default)
    enc_r9 = M.PretrainedEncoder(['relu_9']).eval()
    load_vgg(enc_r9)
    dec = ref_loader.build_reference_classic_decoder()
    dconvs = [l for l in dec if isinstance(l, torch.nn.Conv2d)]
    with torch.no_grad():
        for c, w, b in zip(dconvs, dw, db):
            c.weight.copy_(w)
            c.bias.copy_(b)
    out["vgg_state_keys"] = np.array(sorted(enc_default.state_dict().keys()))
    out["dec_state_keys"] = np.array(sorted(dec.state_dict().keys()))

    x = R.rand_image(1, 32, 61)
    with torch.no_grad():
        taps = enc_default(x)
    out["vgg_x32"] = x.numpy()
    for i, t in enumerate(taps):
        out[f"vgg_x32_tap{i}"] = t.numpy()
    # config-1 shaped run at reduced size (64x64) and the real config 1 (256x256, summary + crop)
    adain = M.AdaIN()
    for size, tag in ((64, "s64"), (256, "cfg1")):
        c, s = R.rand_image(1, size, 101), R.rand_image(1, size, 102)
        with torch.no_grad():
            fc, fs = enc_r9(c)[0], enc_r9(s)[0]
            t = adain(fc, fs)
            img = dec(t)
        if size == 64:
            out[f"{tag}_fc"], out[f"{tag}_t"], out[f"{tag}_img"] = fc.numpy(), t.numpy(), img.numpy()
        else:
            out[f"{tag}_img_crop"] = img[:, :, 96:160, 96:160].numpy()
            out[f"{tag}_img_sub4"] = img[:, :, ::4, ::4].numpy()
            out[f"{tag}_img_stats"] = np.array([img.mean().item(), img.std().item(),
                                                img.min().item(), img.max().item()])
            out[f"{tag}_fc_sub"] = fc[:, ::8, ::2, ::2].numpy()
    # decoder alone on a tiny feature map
    f = _t(71, 1, 512, 4, 4, scale=1.0, shift=0.5, relu=True)
    with torch.no_grad():
        out["dec_in"], out["dec_out"] = f.numpy(), dec(f).numpy()


def golden_autoencoder(M, out):
    """Genuine ``AutoEncoder`` (models.py:322-338) under ``torch.manual_seed(2)``:
    weight checksums, eval forward, train-mode forward (batch statistics + running-stat update) and the
    gradients of one train_autoencoder.py:111-139 loss, at 2 x 3 x 32 x 32."""
    import torch.nn.functional as F
    from oracle import restate_ae as A
    torch.manual_seed(2)
    ae = M.AutoEncoder()
    sd = ae.state_dict()
    keys = sorted(sd.keys())
    out["ae_state_keys"] = np.array(keys)
    out["ae_state_sum"] = np.array([sd[k].double().sum().item() for k in keys])
    out["ae_state_abs"] = np.array([sd[k].double().abs().sum().item() for k in keys])
    x = R.rand_image(2, 32, 301)
    out["ae_x"] = x.numpy()
    ae.eval()
    with torch.no_grad():
        out["ae_eval_recon_fresh"] = ae(x).numpy()
        taps = ae.encoder(x, out_layers=[0, 2, 12, 14])
        for i, t in zip((0, 2, 12, 14), taps):
            out[f"ae_eval_enc{i}"] = t.numpy()
        out["ae_eval_autoenc"] = ae.encoder(x, auto_enc=True).numpy()
    # one training step's loss and gradients (train_autoencoder.py:111-139), VGG = oracle recipe
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    enc = M.PretrainedEncoder().eval()
    convs = [l for l in enc._vgg_layers if isinstance(l, torch.nn.Conv2d)]
    with torch.no_grad():
        for c, w, b in zip(convs, vw, vb):
            c.weight.copy_(w)
            c.bias.copy_(b)
    ae.train()
    recon = ae(x)
    recon_loss = torch.nn.HuberLoss()(recon, x)
    cm, rm = enc(x), enc(recon)
    perp = None
    for a, b in zip(rm, cm):
        l = F.huber_loss(a, b.detach())
        perp = l if perp is None else perp + l
    loss = 100.0 * recon_loss + 0.01 * perp
    loss.backward()
    out["ae_train_recon"] = recon.detach().numpy()
    out["ae_train_losses"] = np.array([loss.item(), recon_loss.item(), perp.item()])
    named = dict(ae.named_parameters())
    gkeys = sorted(named.keys())
    out["ae_grad_keys"] = np.array(gkeys)
    out["ae_grad_norm"] = np.array([named[k].grad.double().norm().item() for k in gkeys])
    for k in A.GOLDEN_GRAD_KEYS:
        out["ae_grad::" + k] = named[k].grad.numpy()
    sd = ae.state_dict()
    for k in A.GOLDEN_BUFFER_KEYS:
        out["ae_buf::" + k] = sd[k].numpy()
    # eval forward with the updated running statistics
    ae.eval()
    with torch.no_grad():
        out["ae_eval_recon_after_step"] = ae(x).numpy()

    # ---- the same on the NON-DEGENERATE variant of the state (restate_ae.activate_gates explains why) ----
    torch.manual_seed(2)
    ae2 = M.AutoEncoder()
    ae2.load_state_dict(A.activate_gates(ae2.state_dict()), strict=True)
    ae2.train()
    recon = ae2(x)
    recon_loss = torch.nn.HuberLoss()(recon, x)
    cm, rm = enc(x), enc(recon)
    perp = None
    for a, b in zip(rm, cm):
        l = F.huber_loss(a, b.detach())
        perp = l if perp is None else perp + l
    loss = 100.0 * recon_loss + 0.01 * perp
    loss.backward()
    out["act_train_recon"] = recon.detach().numpy()
    out["act_train_losses"] = np.array([loss.item(), recon_loss.item(), perp.item()])
    named = dict(ae2.named_parameters())
    out["act_grad_norm"] = np.array([named[k].grad.double().norm().item() for k in gkeys])
    for k in A.GOLDEN_GRAD_KEYS:
        out["act_grad::" + k] = named[k].grad.numpy()
    sd2 = ae2.state_dict()
    for k in A.GOLDEN_BUFFER_KEYS:
        out["act_buf::" + k] = sd2[k].numpy()
    # eval mode with running statistics calibrated on x (bn.momentum = 1.0 for one training-mode forward)
    torch.manual_seed(2)
    ae3 = M.AutoEncoder()
    ae3.load_state_dict(A.activate_gates(ae3.state_dict()), strict=True)
    bns = [m for m in ae3.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = 1.0
    ae3.train()
    with torch.no_grad():
        ae3(x)
    for m in bns:
        m.momentum = 0.1
    ae3.eval()
    with torch.no_grad():
        out["act_eval_recon"] = ae3(x).numpy()
        taps = ae3.encoder(x, out_layers=[0, 2, 12, 14])
        for i, t in zip((0, 2, 12, 14), taps):
            out[f"act_eval_enc{i}"] = t.numpy()
        z = ae3.ada_out(torch.cat((taps[2], taps[3]), dim=1))
        out["act_eval_code"] = z.numpy()
        out["act_eval_dec_of_code"] = ae3.decoder(z).numpy()


def _load_vgg_into(enc, vw, vb):
    convs = [l for l in enc._vgg_layers if isinstance(l, torch.nn.Conv2d)]
    assert len(convs) == 16
    with torch.no_grad():
        for c, w, b in zip(convs, vw, vb):
            c.weight.copy_(w)
            c.bias.copy_(b)


def golden_bigsizes(M, out):
    """The classic path AT THE BENCHMARKED SIZES, from the genuine reference pieces (models.py:186-240 encoder,
    :43-51 AdaIN, :471 alpha blend, :598-628 decoder): BASELINE config 4 (one 512x512 pair, seeds 401 / 402) and
    config 5 (2048x2048 content seed 501, four 2048x2048 styles seed 502, weights (.4,.3,.2,.1), alpha 1.0 and 0.6).
    The K-style mix is not in the reference; by AdaIN's linearity in the style statistics it is
    sum_k w_k * AdaIN(f_c, f_s_k) with the GENUINE AdaIN module (weights sum to 1).  Stored: a 64x64 crop, a stride-8
    subsample, image statistics, and a subsample of the relu4_1 content map."""
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)
    enc = M.PretrainedEncoder(['relu_9']).eval()
    _load_vgg_into(enc, vw, vb)
    dec = ref_loader.build_reference_classic_decoder()
    with torch.no_grad():
        for c, w, b in zip([l for l in dec if isinstance(l, torch.nn.Conv2d)], dw, db):
            c.weight.copy_(w)
            c.bias.copy_(b)
    adain = M.AdaIN()

    def store(tag, img, fc):
        H = img.shape[2]
        a = H // 2 - 32
        out[f"{tag}_img_crop"] = img[:, :, a:a + 64, a:a + 64].numpy()
        out[f"{tag}_img_corner"] = img[:, :, :48, H - 48:].numpy()        # includes two image borders (reflection pad)
        out[f"{tag}_img_sub8"] = img[:, :, ::8, ::8].numpy()
        out[f"{tag}_img_stats"] = np.array([img.mean().item(), img.std().item(), img.min().item(), img.max().item()])
        out[f"{tag}_fc_sub"] = fc[:, ::16, ::4, ::4].numpy()

    with torch.no_grad():
        c, s = R.rand_image(1, 512, 401), R.rand_image(1, 512, 402)
        fc, fs = enc(c)[0], enc(s)[0]
        img = dec(adain(fc, fs))
        store("cfg4", img, fc)
        del fs, img
        c = R.rand_image(1, 2048, 501)
        styles = R.rand_image(4, 2048, 502)
        w = (0.4, 0.3, 0.2, 0.1)
        fc = enc(c)[0]
        t = None
        for k in range(4):
            tk = adain(fc, enc(styles[k:k + 1])[0]) * w[k]
            t = tk if t is None else t + tk
        out["cfg5_t_sub"] = t[:, ::16, ::4, ::4].numpy()
        for alpha, tag in ((1.0, "cfg5_a10"), (0.6, "cfg5_a06")):
            tb = t if alpha == 1.0 else alpha * t + (1 - alpha) * fc        # models.py:471
            store(tag, dec(tb), fc)


def golden_autoencoder256(M, out):
    """The genuine ``AutoEncoder`` at BASELINE config 3's resolution (256x256, batch 2) on the non-degenerate seeded
    state (restate_ae.activate_gates): eval-mode forward with running statistics calibrated on the batch, and one
    train_autoencoder.py:111-139 training step (loss terms, gradient norms of every parameter, full gradients of the
    GOLDEN_GRAD_KEYS, running statistics).  Tensors are stored subsampled."""
    import torch.nn.functional as F
    from oracle import restate_ae as A
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    enc = M.PretrainedEncoder().eval()
    _load_vgg_into(enc, vw, vb)
    x = R.rand_image(2, 256, 301)
    torch.manual_seed(2)
    ae = M.AutoEncoder()
    ae.load_state_dict(A.activate_gates(ae.state_dict()), strict=True)
    # ---- one training step from the seeded state ----
    ae.train()
    recon = ae(x)
    recon_loss = torch.nn.HuberLoss()(recon, x)
    cm, rm = enc(x), enc(recon)
    perp = None
    for a, b in zip(rm, cm):
        l = F.huber_loss(a, b.detach())
        perp = l if perp is None else perp + l
    loss = 100.0 * recon_loss + 0.01 * perp
    loss.backward()
    out["t256_recon_sub4"] = recon.detach()[:, :, ::4, ::4].numpy()
    out["t256_recon_stats"] = np.array([recon.mean().item(), recon.std().item(), recon.min().item(), recon.max().item()])
    out["t256_losses"] = np.array([loss.item(), recon_loss.item(), perp.item()])
    named = dict(ae.named_parameters())
    gkeys = sorted(named.keys())
    out["t256_grad_keys"] = np.array(gkeys)
    out["t256_grad_norm"] = np.array([named[k].grad.double().norm().item() for k in gkeys])
    for k in A.GOLDEN_GRAD_KEYS:
        out["t256_grad::" + k] = named[k].grad.numpy()
    sd = ae.state_dict()
    for k in A.GOLDEN_BUFFER_KEYS:
        out["t256_buf::" + k] = sd[k].numpy()
    # ---- eval mode, running statistics calibrated on x (momentum 1.0 for one training-mode forward) ----
    torch.manual_seed(2)
    ae3 = M.AutoEncoder()
    ae3.load_state_dict(A.activate_gates(ae3.state_dict()), strict=True)
    bns = [m for m in ae3.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = 1.0
    ae3.train()
    with torch.no_grad():
        ae3(x)
    for m in bns:
        m.momentum = 0.1
    ae3.eval()
    with torch.no_grad():
        rec = ae3(x)
        out["e256_recon_sub4"] = rec[:, :, ::4, ::4].numpy()
        out["e256_recon_crop"] = rec[:, :, 96:160, 96:160].numpy()
        taps = ae3.encoder(x, out_layers=[0, 2, 4, 7, 12, 14])
        for i, t in zip((0, 2, 4, 7, 12, 14), taps):
            st = max(1, t.shape[2] // 32)
            out[f"e256_enc{i}_sub"] = t[:, :, ::st, ::st].numpy()
        z = ae3.ada_out(torch.cat((taps[4], taps[5]), dim=1))
        out["e256_code"] = z.numpy()
        out["e256_dec_of_code_sub4"] = ae3.decoder(z)[:, :, ::4, ::4].numpy()


def golden_hist(Ls, out):
    """compute_hist_loss (losses.py:8-87) value and input gradients from the genuine reference: image-like inputs with
    a few values outside [0, 1] (a stylised image is not clamped, models.py:315)."""
    for tag, shape, seed in (("a", (2, 3, 24, 20), 61), ("b", (3, 3, 16, 16), 63)):
        g = torch.Generator().manual_seed(seed)
        x = (torch.rand(*shape, generator=g) * 1.2 - 0.1).requires_grad_(True)
        y = (torch.rand(*shape, generator=g) ** 2).requires_grad_(True)
        l = Ls.compute_hist_loss(x, y)
        (l * 0.5).backward()
        out[f"hist_{tag}_x"], out[f"hist_{tag}_y"] = x.detach().numpy(), y.detach().numpy()
        out[f"hist_{tag}_loss"] = l.detach().numpy()
        out[f"hist_{tag}_gx"], out[f"hist_{tag}_gy"] = x.grad.numpy(), y.grad.numpy()


def restore_ast(M, ast):
    """Restore the two attributes AST.__init__ leaves commented out (models.py:407, 410) although encode / forward
    / train.py use them -- with exactly the commented constructor calls."""
    ast.ada_att_2 = M.AdaAttN(M.enc_out_channels)
    ast.ada_out = M.DepthWiseConv(M.enc_out_channels * 2, M.enc_out_channels, 1, M.EXPAND_RATIO, use_norm=False,
                                  use_identity=False)
    return ast


def golden_adaattn(M, out):
    """Genuine ``AdaAttN`` (models.py:70-115) forward + gradients, and the genuine ``AST`` (models.py:393-575, with
    ada_att_2 / ada_out restored) forward + one loss's gradients, on seeded inputs."""
    from oracle import restate_ae as A, restate_attn as T
    # ---- the layer alone: default init (flat attention) and sharpened (peaked attention), ragged style size ----
    for tag, gain, (h, w, hs, ws) in (("flat", 1.0, (8, 8, 8, 8)), ("sharp", 36.0, (8, 12, 6, 10))):
        torch.manual_seed(11)
        layer = M.AdaAttN(32)
        with torch.no_grad():
            layer.W_q.weight.mul_(gain ** 0.5)
            layer.W_k.weight.mul_(gain ** 0.5)
        c = _t(601, 2, 32, h, w, scale=1.5, shift=0.5).requires_grad_(True)
        s = _t(602, 2, 32, hs, ws, scale=2.0, shift=1.0).requires_grad_(True)
        y = layer(c, s)
        gy = _t(603, *y.shape)
        y.backward(gy)
        out[f"{tag}_content"], out[f"{tag}_style"], out[f"{tag}_gy"] = c.detach().numpy(), s.detach().numpy(), gy.numpy()
        for n in ("W_q", "W_k", "W_v"):
            out[f"{tag}_{n}"] = getattr(layer, n).weight.detach().numpy()
            out[f"{tag}_g{n}"] = getattr(layer, n).weight.grad.numpy()
        out[f"{tag}_out"] = y.detach().numpy()
        out[f"{tag}_gcontent"], out[f"{tag}_gstyle"] = c.grad.numpy(), s.grad.numpy()
    # ---- the network: non-degenerate seeded state (restate_ae.activate_gates) whose encoder running statistics are
    # calibrated on the two images (momentum 1.0 for one training-mode pass): with FRESH running statistics the
    # eval-mode encoder of AST.encode(detach=True) emits taps of ~5e-5, InstanceNorm's eps dominates their variance
    # and the attention is uniform -- a fixture that pins nothing in the layer ----
    sd = A.activate_gates(T.make_ast_state(3))
    torch.manual_seed(5)
    ast = restore_ast(M, M.AST())
    ast.load_state_dict(sd, strict=True)
    out["ast_state_keys"] = np.array(sorted(ast.state_dict().keys()))
    c, s = R.rand_image(2, 64, 701), R.rand_image(2, 64, 702)
    out["ast_content"], out["ast_style"] = c.numpy(), s.numpy()
    bns = [m for m in ast._enc.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = 1.0
    ast._enc.train()
    with torch.no_grad():
        ast._enc(torch.cat((c, s)), out_layers=M.enc_out_layers)
    for m in bns:
        m.momentum = 0.1
    sd = {k: v.detach().clone() for k, v in ast.state_dict().items()}
    out["ast_buf_cal::_enc.mob_net.14._layers.8.running_var"] = sd["_enc.mob_net.14._layers.8.running_var"].numpy()
    ast.train()
    t_cs, t_ret, org = ast(c, s, alpha=0.75)
    loss = torch.nn.functional.huber_loss(t_cs, s) + 0.5 * torch.nn.functional.huber_loss(org, c) + 0.1 * t_ret.mean()
    loss.backward()
    out["ast_t_cs"], out["ast_t_return"], out["ast_org_out"] = t_cs.detach().numpy(), t_ret.detach().numpy(), org.detach().numpy()
    out["ast_loss"] = np.array(loss.item())
    named = dict(ast.named_parameters())
    gk = sorted(k for k, p in named.items() if p.grad is not None)
    out["ast_grad_keys"] = np.array(gk)
    out["ast_grad_norm"] = np.array([named[k].grad.double().norm().item() for k in gk])
    for k in T.GOLDEN_GRAD_KEYS:
        out["ast_grad::" + k] = named[k].grad.numpy()
    out["ast_buf::_enc.mob_net.1._layers.1.running_mean"] = ast.state_dict()["_enc.mob_net.1._layers.1.running_mean"].numpy()
    # exporting network: encode without detach + Hardtanh head (models.py:304, 315-316, 478-479)
    torch.manual_seed(5)
    ast_e = restore_ast(M, M.AST(exporting=True))
    ast_e.load_state_dict(sd, strict=True)
    ast_e.eval()
    with torch.no_grad():
        out["ast_export_t_cs"] = ast_e(c, s).numpy()


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not present: golden vectors can only be made in the build container")
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    M = ref_loader.load_reference_models()
    mu = ref_loader.load_reference_module("model_util")
    Ls = ref_loader.load_reference_module("losses")
    os.makedirs(OUT, exist_ok=True)
    for name, fn, args in (("stats_adain", golden_stats_adain, (M, mu)),
                           ("losses", golden_losses, (Ls,)),
                           ("networks", golden_networks, (M,)),
                           ("autoencoder", golden_autoencoder, (M,)),
                           ("adaattn", golden_adaattn, (M,)),
                           ("hist", golden_hist, (Ls,)),
                           ("bigsizes", golden_bigsizes, (M,)),
                           ("autoencoder256", golden_autoencoder256, (M,))):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        d = dict(_versions())
        fn(*args, d)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **d)
        print(f"wrote {path}: {len(d)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
