"""Drop-in for the part of the reference's mobilenetv2.py the AdaIN / AutoEncoder networks are built from
(mobilenetv2.py:16-43, 63-81, 95-181).  MobileNetV2 / InvertedResidual serve only the reference's discriminator
(models.py:371), which is outside the hot path and is not provided."""
from arbitrarystyletransfer_b200.mobilenet import (DepthWiseConv, SELayer, conv_3x3_bn,  # noqa: F401
                                                   _make_divisible)
