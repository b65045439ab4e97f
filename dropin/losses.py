"""Drop-in for the reference's losses.py (losses.py:83-143): the loss functions the trainers call, on the CUDA
kernels of arbitrarystyletransfer_b200, plus the names its star import re-exports (torch, nn, F, conf, channel_stats)."""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F

from conf import *  # noqa: F401,F403     (losses.py:4)
from model_util import channel_stats  # noqa: F401   (losses.py:6)
from arbitrarystyletransfer_b200.losses import (compute_content_loss, compute_style_loss, gram_matrix,  # noqa: F401
                                                tv_loss, compute_hist_loss)


def discriminator_loss(output, label):
    """losses.py:142-143.  Every call site in the reference is commented out (train.py:175-204); kept as the plain
    ATen call so the name resolves -- the discriminator is outside the hot path."""
    return F.binary_cross_entropy(output, label)
