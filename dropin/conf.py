"""Drop-in for the reference's conf.py: the names `from conf import *` gives the trainers and the model
modules (conf.py:1-7, 71-113).  Topology tables come from the package (arbitrarystyletransfer_b200.mobilenet keeps
them verbatim: they are inputs to the architecture); the dataset directories of the reference are the author's local
paths (conf.py:122-123) and default to empty lists here -- set them before building a loader."""
import torch

from arbitrarystyletransfer_b200.mobilenet import (EXPAND_RATIO, enc_conv_shapes, decoder_conv_shapes,  # noqa: F401
                                                   enc_out_layers, enc_out_channels)

device = "cuda" if torch.cuda.is_available() else "cpu"          # conf.py:3
img_sizes = [96, 128, 160]                                       # conf.py:4: per-batch training resolutions
imsize = 320 if torch.cuda.is_available() else 128               # conf.py:8
expand_ratios = [1, 6, 6, 6, 6, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4]    # conf.py:72 (unused by the live code)
kernel_sizes = [3, 3, 3, 3, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5]       # conf.py:73 (unused by the live code)
content_dir = []
style_dir = []
