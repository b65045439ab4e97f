"""The reference's trainer code paths on the drop-in modules, on the GPU.

/root/reference does not exist on the GPU box, so the loop body of ``AutoencoderTrainer.train``
(train_autoencoder.py:111-148) and the helpers ``interpolate`` / ``get_distr`` (:150-179) are restated here line for
line (each line cites the reference) and run through the SHIM modules under dropin/ -- ``from models import ...``,
``from conf import *``, ``from losses import compute_content_loss`` exactly as train_autoencoder.py:10-14 imports them.
(The genuine trainer class is exercised on the CPU by tests/test_dropin_surface.py.)  Every number is compared with
the CPU oracle (oracle/restate_ae.py) running the same steps with torch's Adam."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import restate as R, restate_ae as A

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = ("models", "losses", "model_util", "mobilenetv2", "conf")


@pytest.fixture()
def shims():
    saved_path = list(sys.path)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in SHIMS}
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        yield {n: importlib.import_module(n) for n in SHIMS}
    finally:
        for k in SHIMS:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
        sys.path[:] = saved_path


def test_autoencoder_trainer_loop_three_iterations_vs_oracle(shims):
    ns = {}
    exec("from models import AutoEncoder, Encoder, PretrainedEncoder\nfrom conf import *\n"
         "from losses import compute_content_loss\nimport torch.optim as optim\nimport torch.nn as nn", ns)
    device = ns["device"]
    assert device == "cuda"
    args = types.SimpleNamespace(lr=2e-4, recon_lam=100.0, perp_lam=0.01, batch_size=2)
    sd = A.activate_gates(A.make_ae_state(2))
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    # ---- AutoencoderTrainer.__init__, train_autoencoder.py:23-28 ----
    model = ns["AutoEncoder"]().to(device)
    model.load_state_dict(sd, strict=True)
    pretrained_mobnet = ns["PretrainedEncoder"]().to(device).eval()
    with torch.no_grad():
        for conv, w, b in zip(pretrained_mobnet._convs(), vw, vb):
            conv.weight.copy_(w)
            conv.bias.copy_(b)
    ae_optim = ns["optim"].Adam(model.parameters(), lr=args.lr, betas=[0.9, 0.99], eps=1e-7)
    loss_fn = ns["nn"].HuberLoss()
    compute_content_loss = ns["compute_content_loss"]
    batches = [R.rand_image(2, 64, 310 + i) for i in range(3)]
    content_iter = iter(batches)
    got = []
    for cur_iter in range(3):
        content_imgs = next(content_iter).to(device)                               # :111
        recon_imgs = model(content_imgs)                                           # :112
        ae_optim.zero_grad()                                                       # :113 (set_to_none=True)
        recon_loss = loss_fn(recon_imgs, content_imgs)                             # :114
        content_maps = pretrained_mobnet(content_imgs)                             # :117
        recon_maps = pretrained_mobnet(recon_imgs)                                 # :118
        for i in range(len(content_maps)):                                         # :122-134
            content_weight = 1.0
            if i == 0:
                content_loss = compute_content_loss(recon_maps[i], content_maps[i].detach()) * content_weight
            else:
                content_loss = content_loss + compute_content_loss(recon_maps[i], content_maps[i].detach()) * content_weight
        loss = args.recon_lam * recon_loss + args.perp_lam * content_loss         # :140
        loss.backward()                                                            # :142
        ns["nn"].utils.clip_grad.clip_grad_norm_(model.parameters(), 10.0)         # :143
        assert model.encoder.mob_net[0][0].weight.grad is not None                 # :145
        assert model.decoder._img_out.weight.grad is not None                      # :146
        ae_optim.step()                                                            # :148
        got.append((recon_loss.item(), content_loss.item(), loss.item()))
    # ---- the same three steps on the CPU oracle ----
    P = A.clone_state(sd, requires_grad=True)
    train = [P[k] for k in sorted(P) if P[k].requires_grad]
    opt = torch.optim.Adam(train, lr=args.lr, betas=[0.9, 0.99], eps=1e-7)
    want = []
    for x in batches:
        opt.zero_grad()
        loss, recon_loss, perp, _ = A.ae_losses(P, x, vw, vb)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(train, 10.0)
        opt.step()
        want.append((recon_loss.item(), perp.item(), loss.item()))
    print("trainer loop (recon, perceptual, total): cuda", got, "oracle", want)
    for g, w in zip(got, want):
        assert g[0] == pytest.approx(w[0], rel=2e-2)
        assert g[1] == pytest.approx(w[1], rel=5e-2)
        assert g[2] == pytest.approx(w[2], rel=2e-2)
    # the weights after three Adam steps moved the same way
    named = dict(model.named_parameters())
    for k in ("decoder._img_out.weight", "ada_out._layers.0.weight", "encoder.mob_net.1._layers.0.weight"):
        d_got = (named[k].detach().cpu() - sd[k]).flatten().double()
        d_ref = (P[k].detach() - sd[k]).flatten().double()
        cos = torch.nn.functional.cosine_similarity(d_got, d_ref, dim=0).item()
        assert cos > 0.9, (k, cos)      # Adam normalises each step to ~lr * sign: direction agreement, element by element


def test_trainer_helpers_interpolate_and_get_distr_vs_oracle(shims):
    """train_autoencoder.py:150-179 restated on the drop-in modules, and the package's own helpers
    (arbitrarystyletransfer_b200.trainer_util), against the oracle."""
    from arbitrarystyletransfer_b200 import trainer_util as TU
    AutoEncoder = shims["models"].AutoEncoder
    x1, x2 = R.rand_image(2, 64, 321), R.rand_image(2, 64, 322)
    Q = A.calibrate_running_stats(A.clone_state(A.activate_gates(A.make_ae_state(2))), torch.cat((x1, x2)))
    model = AutoEncoder().cuda()
    model.load_state_dict({k: v for k, v in Q.items()}, strict=True)
    model.eval()
    alpha = 0.3
    with torch.no_grad():
        # reference body, :167-177
        img_enc_1 = model.encoder(x1.cuda(), auto_enc=True)
        img_enc_2 = model.encoder(x2.cuda(), auto_enc=True)
        img_enc_inter = alpha * img_enc_1 + (1 - alpha) * img_enc_2
        img_inter = model.decoder(img_enc_inter).cpu()
        mine = TU.interpolate(model, x1.cuda(), x2.cuda(), alpha).cpu()
        e1, e2 = A.encoder_forward(Q, x1, auto_enc=True), A.encoder_forward(Q, x2, auto_enc=True)
        ref = A.decoder_forward(Q, alpha * e1 + (1 - alpha) * e2)
    assert R.psnr(img_inter, ref) >= 40.0 and R.psnr(mine, ref) >= 40.0
    assert ((mine - img_inter).norm() / img_inter.norm()).item() < 5e-3
    # get_distr, :150-164
    bs, ns_ = 2, 3
    batches = [R.rand_image(bs, 64, 330 + i) for i in range(ns_)]
    got = TU.get_distr(model, iter(batches), bs, ns_).cpu()
    with torch.no_grad():
        enc_sum = None
        for b in batches:
            e = A.encoder_forward(Q, b, auto_enc=True).sum(axis=0)
            enc_sum = e if enc_sum is None else enc_sum + e
        want = (enc_sum / (bs * ns_)).sum(axis=0)
    assert got.shape == want.shape == (8, 8)
    # a sum over 128 channels of the deepest feature (8 x 8 maps here): cancellation amplifies the per-element error
    assert ((got - want).norm() / want.norm()).item() < 0.15
    assert not model.training
