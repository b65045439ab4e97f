"""Trainer-side helpers of the reference's AutoencoderTrainer that run the networks:
``interpolate`` (train_autoencoder.py:166-179) and ``get_distr`` (train_autoencoder.py:150-164), as free functions
over an ``AutoEncoder``.  The reference's own methods work unchanged on the drop-in modules (they only call
``model.encoder(x, auto_enc=True)`` and ``model.decoder(z)``); these versions blend with the ``ast_axpby`` kernel
(one pass) instead of three ATen passes and never compute on the host."""
from __future__ import annotations

import torch

from . import _lib as L


def _axpby(x, y, a, b):
    """a * x + b * y on fp32 tensors of one shape (the alpha blend of models.py:471 / train_autoencoder.py:174)."""
    lib = L.load()
    x, y = x.float().contiguous(), y.float().contiguous()
    out = torch.empty_like(x)
    L.check(lib.ast_axpby(x.data_ptr(), y.data_ptr(), float(a), float(b), out.data_ptr(), x.numel(),
                          L.stream_ptr(x.device)), "ast_axpby")
    return out


@torch.no_grad()
def interpolate(model, img_1: torch.Tensor, img_2: torch.Tensor, alpha: float = 0.5) -> torch.Tensor:
    """train_autoencoder.py:166-179: decode ``alpha * enc(img_1) + (1 - alpha) * enc(img_2)`` with the final encoder
    feature (``auto_enc=True``, models.py:183).  Images (N,3,H,W) fp32 in [0,1] -> (N,3,H,W) fp32."""
    L.require_cuda(img_1, img_2)
    e1 = model.encoder(img_1, auto_enc=True)
    e2 = model.encoder(img_2, auto_enc=True)
    return model.decoder(_axpby(e1, e2, alpha, 1.0 - alpha))


@torch.no_grad()
def get_distr(model, content_iter, batch_size: int, num_samples: int = 16) -> torch.Tensor:
    """train_autoencoder.py:150-164: the mean final-encoder feature over ``num_samples`` batches, summed over the
    channel axis -- ``(sum_batches sum_n enc / (batch_size * num_samples)).sum(axis=0)``, shape (H/8, W/8).  The model
    is put in eval mode as the reference does (:152)."""
    model.eval()
    enc_sum = None
    for _ in range(num_samples):
        x = next(content_iter)
        x = x.to(next(model.parameters()).device)
        e = model.encoder(x, auto_enc=True).sum(dim=0)
        enc_sum = e if enc_sum is None else enc_sum + e
    return (enc_sum / (batch_size * num_samples)).sum(dim=0)
