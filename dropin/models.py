"""Drop-in for the reference's models.py: ``from models import *`` (train.py:15) and
``from models import AutoEncoder, Encoder, PretrainedEncoder`` (train_autoencoder.py:11) bind to the B200 package.

The reference's star import also re-exports everything models.py itself imported (models.py:1-11): ``torch``, ``nn``,
``F``, ``transforms``, ``random``, the whole of ``conf`` (``device``, ``enc_out_layers``, ``enc_out_channels``, ...)
and of ``losses`` (``compute_content_loss``, ``compute_style_loss``, ``tv_loss``, ``compute_hist_loss``, ...), and
``channel_stats``; train.py relies on those arriving this way.  This module has no ``__all__`` for the same reason."""
import random  # noqa: F401

import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from model_util import channel_stats, rgb2lab, lab2rgb  # noqa: F401      (models.py:4)
from conf import *  # noqa: F401,F403                                     (models.py:5)
from losses import *  # noqa: F401,F403                                   (models.py:6)
from mobilenetv2 import DepthWiseConv, conv_3x3_bn  # noqa: F401          (models.py:7)

try:                                                                      # models.py:8-9 (host-side, optional here)
    import torchvision.models as models  # noqa: F401
    import torchvision.transforms as transforms  # noqa: F401
except Exception:                                                         # pragma: no cover
    pass

from arbitrarystyletransfer_b200.models import (AdaIN, calc_mean_std, mean_variance_norm, PretrainedEncoder,  # noqa: F401
                                                ClassicDecoder, StyleTransferNet, AdaAttN, AST, Encoder, Decoder,
                                                DecoderBlock, AutoEncoder)
