"""ctypes binding of libast_b200.so (C ABI declared in include/ast_b200.h).

There is NO CPU fallback: if the shared library is missing the import fails loudly, and every
wrapper refuses non-CUDA tensors.  PyTorch only provides device memory, streams and autograd.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# AST_B200_LIB: developer override for A/B runs against another build of the same ABI (tools/); never a fallback
LIB_PATH = os.environ.get("AST_B200_LIB") or os.path.join(HERE, "libast_b200.so")

# mirrors of the header constants
ABI_VERSION = 3
MAX_STYLES = 8
F_CANONICAL, F_BIASED, F_BF16 = 0x1, 0x2, 0x4
EPI_PLAIN, EPI_POOL2, EPI_UP2, EPI_UPFOLD = 0, 1, 2, 4
HALO_KEEP, HALO_REFLECT, HALO_CLAMP = 0, 1, 2
CONV_AUTO, CONV_TC, CONV_DIRECT, CONV_TC_TAPBOX = 0, 1, 2, 3
DT_BF16, DT_F16 = 0, 1     # 16-bit storage formats of the MobileNet-style path: gradients / activations


class AstError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("N", "H", "W", "Cin", "Cout", "relu", "epilogue", "halo", "impl", "tap_prerelu")]


_vp, _i, _i64, _f, _u, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint, C.c_size_t
_d = C.c_double
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol include/ast_b200.h declares
PROTOTYPES = {
    "ast_abi_version": (_i, []),
    "ast_error_string": (C.c_char_p, [_i]),
    "ast_device_info": (_i, [C.POINTER(_i)] * 3),
    "ast_adain_fwd": (_i, [_vp, C.POINTER(_vp), C.POINTER(_i64), _fp, _i, _vp, _vp, _i, _i, _i64,
                           _f, _f, _u, _vp]),
    "ast_channel_stats_fwd": (_i, [_vp, _vp, _vp, _i64, _i64, _f, _u, _vp]),
    "ast_channel_stats_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _u, _vp]),
    "ast_mvn_fwd": (_i, [_vp, _vp, _vp, _i64, _i64, _f, _u, _vp]),
    "ast_mvn_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _u, _vp]),
    "ast_adain_bwd": (_i, [_vp, _vp, _vp, _fp, _i, _f, _vp, _vp, _i64, _i64, _u, _vp]),
    "ast_huber_ws_bytes": (_sz, [_i64]),
    "ast_huber_fwd": (_i, [_vp, _vp, _vp, _i64, _f, _vp, _sz, _vp]),
    "ast_huber_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _vp]),
    "ast_gram_bwd_tc_ws_bytes": (_sz, [_i, _i, _i64]),
    "ast_gram_bwd_tc": (_i, [_vp, _vp, _vp, _i, _i, _i64, _vp, _sz, _vp]),
    "ast_hist_ws_bytes": (_sz, [_i]),
    "ast_hist_loss_fwd": (_i, [_vp, _vp, _i, _i64, _i64, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "ast_hist_loss_bwd": (_i, [_vp, _vp, _vp, _f, _f, _vp, _i, _i64, _vp]),
    "ast_tv_fwd": (_i, [_vp, _vp, _i64, _i, _i, _vp, _sz, _vp]),
    "ast_tv_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp]),
    "ast_gram_fwd": (_i, [_vp, _vp, _i, _i, _i64, _vp]),
    "ast_gram_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i64, _vp]),
    "ast_gram_fwd_tf32": (_i, [_vp, _vp, _i, _i, _i64, _vp]),
    "ast_conv3x3_fwd": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp]),
    "ast_pack_conv_weight": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ast_pack_conv_weight_fold": (_i, [_vp, _vp, _i, _i, _vp]),
    "ast_conv12_fused": (_i, [_vp, _vp, _vp, _fp, _fp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ast_conv3x3_first": (_i, [_vp, _vp, _vp, _fp, _fp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_conv3x3_last": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_nchw_to_native": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ast_native_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ast_u8hwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ast_nchw_to_u8hwc": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ast_adain_native_ws_bytes": (_sz, [_i, _i, _i]),
    "ast_adain_native_fwd": (_i, [_vp, C.POINTER(_vp), _fp, _i, _vp, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), _f, _f,
                                  _u, _i, _vp, _sz, _vp]),
    "ast_pack_conv_weight_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ast_nchw_to_native_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_native_to_nchw_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ast_maxpool2_native": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ast_vgg_bwd_prep": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_dec_bwd_fold": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_native_to_planar": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_conv3x3_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_unpack_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i64, _i, _vp]),
    "ast_zero_halo": (_i, [_vp, _i, _i, _i, _i, _i, _vp]),
    "ast_conv3x3_wgrad_native": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ast_set_act_format": (_i, [_i]),
    "ast_get_act_format": (_i, []),
    "ast_pw_conv": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i64, _i, _i, _vp, _i, _i, _i, _vp]),
    "ast_dw_conv": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_se_fc": (_i, [_vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ast_scale_weights": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ast_stem_conv": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ast_head_conv": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_nhwc_to_nchw": (_i, [_vp, _i, _vp, _i, _i, _i64, _i, _vp]),
    "ast_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i64, _i, _vp]),
    "ast_cvt_f16_to_bf16": (_i, [_vp, _i64, _vp, _i64, _i64, _i, _vp]),
    "ast_bn_stats": (_i, [_vp, _i, _vp, _i, _i, _i64, _vp]),
    "ast_bn_finalize": (_i, [_vp, _d, _vp, _vp, _vp, _vp, _f, _f, _vp, _i, _vp, _vp]),
    "ast_affine_act": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i64, _vp]),
    "ast_dw_bwd_reduce": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i64, _vp]),
    "ast_se_bn_combine": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _d, _vp]),
    "ast_dw_bwd_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp]),
    "ast_bn_bwd_reduce": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i64, _vp]),
    "ast_bn_bwd_finalize": (_i, [_vp, _d, _vp, _vp, _vp, _i, _vp]),
    "ast_bn_bwd_apply": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i64, _vp]),
    "ast_dw_conv_dgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_dw_conv_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ast_se_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ast_pw_wgrad": (_i, [_vp, _i, _i, _vp, _i, _i, _i64, _vp, _i64, _i64, _vp]),
    "ast_stem_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ast_stem_dgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ast_hardtanh01_bwd": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "ast_head_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ast_head_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ast_prep_weight": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ast_prep_block_weights": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ast_bgemm": (_i, [_vp, _i, _i, _i64, _vp, _i, _i, _i64, _vp, _i, _i64, _i64, _i, _i, _i, _i, _vp]),
    "ast_attn_softmax": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _vp]),
    "ast_attn_softmax_bwd": (_i, [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i, _vp]),
    "ast_attn_vv3": (_i, [_vp, _i64, _vp, _i64, _i, _vp]),
    "ast_attn_vv5": (_i, [_vp, _i64, _vp, _i64, _i, _vp]),
    "ast_attn_out_fwd": (_i, [_vp, _vp, _vp, _i64, _vp, _i64, _i, _vp]),
    "ast_attn_out_bwd": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _vp]),
    "ast_split3_rows": (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp]),
    "ast_split3_nchw": (_i, [_vp, _vp, _i, _i, _i64, _i, _vp]),
    "ast_attn_dv": (_i, [_vp, _vp, _i64, _vp, _i64, _i, _vp]),
    "ast_axpby": (_i, [_vp, _vp, _f, _f, _vp, _i64, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libast_b200.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise AstError(
            f"{LIB_PATH} not found: build it with `python -m arbitrarystyletransfer_b200._build` "
            "(there is no CPU or PyTorch fallback for the AdaIN hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.ast_abi_version() != ABI_VERSION:
        raise AstError(f"libast_b200.so ABI {lib.ast_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ast_error_string(rc).decode()
        raise AstError(f"{what or 'libast_b200'} failed with code {rc}: {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise AstError("arbitrarystyletransfer_b200 runs on CUDA tensors only (no CPU fallback); "
                           f"got a tensor on {t.device}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def float_array(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])
