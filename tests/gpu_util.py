"""Helpers shared by the -m gpu parity tests (CUDA path vs the CPU oracle)."""
import torch
import torch.nn.functional as F


def bf16r(x):
    """round-trip through bf16 (what the native layout stores)."""
    return x.to(torch.bfloat16).float()


def oracle_conv_native(x, w, b, relu, epilogue, pad_mode):
    """CPU fp32 conv on bf16-rounded operands, then the epilogue, then bf16 rounding: the exact
    arithmetic contract of ast_conv3x3_fwd (fp32 accumulation order aside).
    x (N,Cin,H,W) fp32 (already bf16-representable), w OIHW fp32, epilogue in {0 plain, 1 pool, 2 up}."""
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect" if pad_mode == "reflect" else "constant")
    y = F.conv2d(xp, bf16r(w), b)
    pre = y
    if relu:
        y = F.relu(y)
    post = y
    if epilogue == 1:
        y = F.max_pool2d(y, 2, 2)
    elif epilogue == 2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    return bf16r(y), pre, post


def native_to_padded_nchw(t):
    """bf16 [N][H+2][W+2][C] -> fp32 (N,C,H+2,W+2) including the halo."""
    return t.float().permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
