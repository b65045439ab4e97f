"""Drop-in for the hot-path part of the reference's losses.py (losses.py:90-139).

Same names, argument meaning and return convention (0-dim tensors with autograd history); the
device work is libast_b200's Huber / Gram / channel-statistics kernels."""
from __future__ import annotations

import torch

from . import functional as Fn
from .model_util import channel_stats


def gram_matrix(tensor: torch.Tensor) -> torch.Tensor:
    """X X^T / (C*H*W), X = (B, C, H*W) -- losses.py:105-109."""
    return Fn.gram_matrix(tensor)


def compute_content_loss(inp: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """F.huber_loss(inp, tgt): delta 1, mean reduction -- losses.py:124-126."""
    return Fn.huber_loss(inp, tgt)


def compute_style_loss(t_cs_map: torch.Tensor, style_map: torch.Tensor) -> torch.Tensor:
    """1.25 huber(mean) + 1.25 huber(std) + 10 huber(gram) -- losses.py:128-139."""
    enc_mean, enc_std = channel_stats(t_cs_map)
    style_mean, style_std = channel_stats(style_map)
    mean_loss = Fn.huber_loss(enc_mean, style_mean, 1.25)
    std_loss = Fn.huber_loss(enc_std, style_std, 1.25)
    g_c = gram_matrix(t_cs_map)
    g_s = gram_matrix(style_map)
    gram_loss = Fn.huber_loss(g_c, g_s, 10.0)
    return mean_loss + std_loss + gram_loss


def tv_loss(img: torch.Tensor) -> torch.Tensor:
    """Total variation: sum of squared differences between horizontal and vertical neighbours, un-normalised --
    losses.py:90-103 (train.py:266 weights it with args.tv_lam)."""
    return Fn.tv_loss(img)


def compute_hist_loss(t_cs: torch.Tensor, style_map: torch.Tensor) -> torch.Tensor:
    """Squared earth mover's distance between the soft 256-bin histograms of the two tensors (all C*H*W values of
    each sample in one histogram, normalised by C*H as the reference does), averaged over the batch --
    losses.py:8-87 (train.py:261 weights it with 1e-5).  One fused pass per tensor; the reference's
    (B, 256, C*H*W) intermediates are never formed."""
    return Fn.hist_loss(t_cs, style_map)
