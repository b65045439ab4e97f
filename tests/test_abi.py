"""The C-ABI shared library: loads, exports every symbol include/ast_b200.h declares, and the
ctypes binding covers exactly that set.  No compute calls (CPU only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ast_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ast_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from arbitrarystyletransfer_b200 import _build
    return _build.build()


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("ast_adain_fwd", "ast_conv3x3_fwd", "ast_channel_stats_fwd", "ast_huber_fwd",
                 "ast_gram_fwd", "ast_abi_version"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in ast_b200.h but not exported"


def test_binding_matches_header(lib_path):
    from arbitrarystyletransfer_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    lib = _lib.load()
    assert lib.ast_abi_version() == _lib.ABI_VERSION
    assert b"success" in lib.ast_error_string(0)
    assert b"workspace" in lib.ast_error_string(-6)


def test_argument_errors_without_gpu(lib_path):
    """Argument validation happens before any CUDA call, so it is checkable on CPU."""
    from arbitrarystyletransfer_b200 import _lib
    lib = _lib.load()
    assert lib.ast_adain_fwd(None, None, None, None, 1, None, None, 1, 1, 1, 1.0, 0.0, 0, None) == -1
    assert lib.ast_gram_fwd(None, None, 1, 1, 1, None) == -1
    assert lib.ast_huber_ws_bytes(10) >= 4
    # K6 / hist-loss entry points
    assert lib.ast_bgemm(None, 0, 8, 0, None, 0, 8, 0, None, 0, 8, 0, 1, 1, 1, 1, None) == -1
    assert lib.ast_attn_softmax(None, 8, None, 8, None, 1, 8, None) == -1
    assert lib.ast_hist_ws_bytes(4) >= 2 * 4 * 257 * 8
    assert lib.ast_hist_loss_fwd(None, None, 1, 1, 1, 1.0, 1.0, None, None, None, 0, None) == -1
    assert lib.ast_split3_rows(None, 8, None, 1, 8, 0, None) == -1
    # round 2: the native-layout weight gradient (K2wn) and the halo-ring zeroing
    assert lib.ast_conv3x3_wgrad_native(None, 64, 2, None, None, None, 1, 8, 8, 64, 64, None) == -1
    assert lib.ast_zero_halo(None, 1, 8, 4, 4, 1, None) == -1


def test_no_cpu_fallback():
    import torch
    from arbitrarystyletransfer_b200 import _lib, models
    with pytest.raises(_lib.AstError):
        models.AdaIN()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(_lib.AstError):
        models.mean_variance_norm(torch.zeros(1, 2, 4, 4))


def test_sass_is_blackwell_native(lib_path):
    """The conv kernel must contain tcgen05 MMA, TMEM loads and TMA loads (SASS mnemonics)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    # round 2: CTA-pair MMAs with multicast commits, TMA loads signalling the leader's barrier, and a TMA store
    assert "UTCHMMA.2CTA" in sass and "UTCBAR.2CTA.MULTICAST" in sass and ".2CTA" in sass and "UTMASTG" in sass
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    # the fused conv1_1 + conv1_2 kernel: ptxas 12.9 once dropped the high word of its A-operand descriptor (stride,
    # version, 128-byte-swizzle layout type = 0x40004050) when it was built as one 64-bit sum; every instantiation
    # must carry the constant (csrc/conv12_fused.cuh)
    parts = sass.split("Function : ")
    fused = [p for p in parts if p.startswith("_ZN3ast2tc24conv12_fused_pair_kernel")]
    assert len(fused) >= 2
    for body in fused:
        assert "0x40004050" in body and "UTCHMMA.2CTA" in body
    assert any("USETMAXREG" in body for body in fused)        # the default variant rebalances registers per role
    # K2wn: MN-major tcgen05 GEMM fed by 4-D TMA boxes, vector reductions into the packed gradient
    wn = [p for p in parts if p.startswith("_ZN3ast2tc18wgrad3x3_mn_kernel")]
    assert len(wn) == 1 and "UTCHMMA" in wn[0] and "UTMALDG.4D" in wn[0] and "REDG.E.ADD.F32x4" in wn[0]
