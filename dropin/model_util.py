"""Drop-in for the reference's model_util.py: ``channel_stats`` (model_util.py:3-8) on the fused CUDA kernel, and the
Lab colour conversions the reference's data_loader.py imports (model_util.py:13-139).  The colour code is host-side
data preparation (out of the hot path): plain torch, written in matrix form."""
import torch

from arbitrarystyletransfer_b200.model_util import channel_stats  # noqa: F401

_RGB2XYZ = ((0.412453, 0.357580, 0.180423), (0.212671, 0.715160, 0.072169), (0.019334, 0.119193, 0.950227))
_XYZ2RGB = ((3.24048134, -1.53715152, -0.49853633), (-0.96925495, 1.87599, 0.04155593),
            (0.05564664, -0.20404134, 1.05731107))
_WHITE = (0.95047, 1.0, 1.08883)     # D65


def _mix(m, x):
    w = torch.tensor(m, dtype=x.dtype, device=x.device)
    return torch.einsum("oc,nchw->nohw", w, x)


def rgb2xyz(rgb):
    """sRGB in [0, 1] (N,3,H,W) -> CIE XYZ (model_util.py:13-36)."""
    lin = torch.where(rgb > 0.04045, ((rgb + 0.055) / 1.055) ** 2.4, rgb / 12.92)
    return _mix(_RGB2XYZ, lin)


def xyz2rgb(xyz):
    """model_util.py:38-59."""
    rgb = _mix(_XYZ2RGB, xyz).clamp_min(0)
    return torch.where(rgb > 0.0031308, 1.055 * rgb ** (1.0 / 2.4) - 0.055, 12.92 * rgb)


def xyz2lab(xyz):
    """model_util.py:65-88."""
    s = xyz / torch.tensor(_WHITE, dtype=xyz.dtype, device=xyz.device).view(1, 3, 1, 1)
    f = torch.where(s > 0.008856, s ** (1.0 / 3.0), 7.787 * s + 16.0 / 116.0)
    return torch.stack((116.0 * f[:, 1] - 16.0, 500.0 * (f[:, 0] - f[:, 1]), 200.0 * (f[:, 1] - f[:, 2])), dim=1)


def lab2xyz(lab):
    """model_util.py:90-115."""
    y = (lab[:, 0] + 16.0) / 116.0
    f = torch.stack((lab[:, 1] / 500.0 + y, y, (y - lab[:, 2] / 200.0).clamp_min(0)), dim=1)
    out = torch.where(f > 0.2068966, f ** 3.0, (f - 16.0 / 116.0) / 7.787)
    return out * torch.tensor(_WHITE, dtype=lab.dtype, device=lab.device).view(1, 3, 1, 1)


def rgb2lab(rgb):
    """model_util.py:117-128: Lab rescaled to roughly [0, 1] as (lab / 100 + 1) / 2."""
    return (xyz2lab(rgb2xyz(rgb)) / 100 + 1) / 2


def lab2rgb(lab_rs):
    """model_util.py:130-139."""
    return xyz2rgb(lab2xyz((lab_rs * 2 - 1) * 100))
