"""K2 parity: tcgen05 implicit-GEMM conv (and the CUDA-core kernels) vs the CPU oracle arithmetic
(torch fp32 conv on bf16-rounded operands = what models.py:186-240 / 598-628 compute, at the
storage precision of the native layout).  Tolerance: one bf16 ulp (2^-8 relative) on outputs that
sit on a rounding boundary, i.e. rtol 8e-3 elementwise and < 2e-3 relative L2."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import bf16r, oracle_conv_native, native_to_padded_nchw, rel_err

pytestmark = pytest.mark.gpu


def _run(N, H, W, cin, cout, relu, epi, halo_reflect, impl, seed=0, with_tap=None):
    from arbitrarystyletransfer_b200 import _lib as L, engine as E
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(N, cin, H, W, generator=g))
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    pad_mode = "reflect" if halo_reflect else "zeros"
    exp, pre, post = oracle_conv_native(x, w, b, relu, epi, pad_mode)
    dev = "cuda"
    xin = E.nchw_to_native(x.to(dev), reflect=halo_reflect)
    wpk = E.pack_conv_weight(w.to(dev))
    Ho, Wo = (H // 2, W // 2) if epi == 1 else ((2 * H, 2 * W) if epi == 2 else (H, W))
    out = torch.full((N, Ho + 2, Wo + 2, cout), 7.0, device=dev, dtype=torch.bfloat16)
    tap = None
    if with_tap is not None:
        tap = torch.full((N, cout, H, W), -99.0, device=dev, dtype=torch.float32)
    E.conv3x3(xin, wpk, b.to(dev), out, N=N, H=H, W=W, cin=cin, cout=cout, relu=relu, epilogue=epi,
              halo=L.HALO_REFLECT if halo_reflect else L.HALO_KEEP, impl=impl, tap=tap,
              tap_prerelu=bool(with_tap == "pre"))
    torch.cuda.synchronize()
    got = native_to_padded_nchw(out).cpu()
    inner = got[:, :, 1:-1, 1:-1]
    assert rel_err(inner, exp) < 2e-3, f"interior rel err {rel_err(inner, exp)}"
    torch.testing.assert_close(inner, exp, rtol=8e-3, atol=2e-3)
    if halo_reflect:
        torch.testing.assert_close(got, F.pad(exp, (1, 1, 1, 1), mode="reflect"), rtol=8e-3, atol=2e-3)
    else:  # halo untouched
        assert (got[:, :, 0, :] == 7.0).all() and (got[:, :, :, -1] == 7.0).all()
    if tap is not None:
        ref = pre if with_tap == "pre" else post
        torch.testing.assert_close(tap.cpu(), ref, rtol=2e-3, atol=2e-3)
    return inner, exp


SHAPES_TC = [
    # N, H, W, cin, cout
    (1, 16, 16, 64, 64),
    (2, 8, 16, 64, 128),
    (1, 32, 32, 128, 256),
    (1, 24, 40, 64, 64),      # partial tiles in both directions
    (1, 12, 20, 128, 128),    # 96/160-pixel images at /8 (conf.py:4 img_sizes)
    (1, 16, 32, 256, 512),
    (3, 10, 18, 64, 256),
]


KERNELS = {"kwbox": 1, "tapbox": 3}   # AST_CONV_TC (default kernel), AST_CONV_TC_TAPBOX


@pytest.mark.parametrize("kern", list(KERNELS))
@pytest.mark.parametrize("shape", SHAPES_TC)
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_tc_conv_zero_pad(shape, epi, kern):
    N, H, W, cin, cout = shape
    _run(N, H, W, cin, cout, True, epi, False, KERNELS[kern], seed=epi)


@pytest.mark.parametrize("kern", list(KERNELS))
@pytest.mark.parametrize("shape", SHAPES_TC[:5])
@pytest.mark.parametrize("epi", [0, 2])
def test_tc_conv_reflect_halo(shape, epi, kern):
    N, H, W, cin, cout = shape
    _run(N, H, W, cin, cout, True, epi, True, KERNELS[kern], seed=3 + epi)


def _fold_reference(x, w, b, relu, wq=None):
    """Upsample(x2, nearest) -> ReflectionPad2d(1) -> Conv2d(3x3) (models.py:602-604) in fp32 on the CPU; with ``wq``
    (a rounding function) the same thing as the four parity-specific 2x2 convs with pre-summed, rounded weights --
    the arithmetic contract of AST_EPI_UPFOLD."""
    if wq is None:
        y = F.conv2d(F.pad(F.interpolate(x, scale_factor=2, mode="nearest"), (1, 1, 1, 1), mode="reflect"), w, b)
    else:
        N, _, H, W = x.shape
        xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
        y = torch.empty(N, w.shape[0], 2 * H, 2 * W)
        rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
        for py in (0, 1):
            for px in (0, 1):
                wf = torch.zeros(w.shape[0], w.shape[1], 2, 2)
                for a in (0, 1):
                    for bb in (0, 1):
                        for kh in rows[py][a]:
                            for kw in rows[px][bb]:
                                wf[:, :, a, bb] += w[:, :, kh, kw]
                yp = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], wq(wf), b)
                y[:, :, py::2, px::2] = yp
    return F.relu(y) if relu else y


def _run_fold(N, H, W, cin, cout, relu, out_halo, impl, seed=0):
    from arbitrarystyletransfer_b200 import _lib as L, engine as E
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(N, cin, H, W, generator=g))
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    exact = bf16r(_fold_reference(x, w, b, relu, wq=bf16r))
    ref32 = _fold_reference(x, w, b, relu)
    assert rel_err(exact, ref32) < 4e-3          # the fold itself (pre-summed bf16 weights) vs the fp32 sequence
    xin = F.pad(x, (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    wf = E.pack_conv_weight_fold(w.cuda())
    out = torch.full((N, 2 * H + 2, 2 * W + 2, cout), 7.0, device="cuda", dtype=torch.bfloat16)
    E.conv3x3(xin, wf, b.cuda(), out, N=N, H=H, W=W, cin=cin, cout=cout, relu=relu, epilogue=L.EPI_UPFOLD,
              halo=out_halo, impl=impl)
    torch.cuda.synchronize()
    got = native_to_padded_nchw(out).cpu()
    inner = got[:, :, 1:-1, 1:-1]
    assert rel_err(inner, exact) < 2e-3, rel_err(inner, exact)
    torch.testing.assert_close(inner, exact, rtol=8e-3, atol=2e-3)
    assert rel_err(inner, ref32) < 5e-3
    mode = {L.HALO_REFLECT: "reflect", L.HALO_CLAMP: "replicate"}.get(out_halo)
    if mode:
        torch.testing.assert_close(got, F.pad(exact, (1, 1, 1, 1), mode=mode), rtol=8e-3, atol=2e-3)
    else:
        assert (got[:, :, 0, :] == 7.0).all() and (got[:, :, :, -1] == 7.0).all()


FOLD_SHAPES = [
    # N, H, W, cin, cout   (low-res input)
    (1, 16, 8, 64, 64),       # one tile, resident weights
    (2, 20, 12, 64, 64),      # partial tiles
    (1, 16, 16, 128, 128),    # BN = 128: one row parity per tile
    (1, 8, 8, 256, 256),      # BN = 256: one parity per tile
    (3, 12, 20, 128, 64),     # two channel blocks, not resident
    (1, 24, 40, 64, 128),
    (2, 9, 7, 256, 128),      # ragged
]


@pytest.mark.parametrize("shape", FOLD_SHAPES)
def test_upsample_fold_conv(shape):
    """Upsample -> ReflectionPad -> Conv3x3 on the low-res map (conv3x3_fold_kernel)."""
    from arbitrarystyletransfer_b200 import _lib as L
    N, H, W, cin, cout = shape
    _run_fold(N, H, W, cin, cout, True, L.HALO_REFLECT, L.CONV_TC, seed=H + cin)


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_upsample_fold_forced_n_block_many_tiles(bn):
    from arbitrarystyletransfer_b200 import _lib as L
    _run_fold(2, 64, 48, 128, 256, True, L.HALO_REFLECT, bn, seed=bn)       # > 148 tiles: stage / phase wrap
    _run_fold(1, 16, 8, 64, 256, False, L.HALO_KEEP, bn, seed=bn + 1)


def test_upsample_fold_many_tiles_resident():
    from arbitrarystyletransfer_b200 import _lib as L
    _run_fold(4, 96, 64, 64, 64, True, L.HALO_REFLECT, L.CONV_TC, seed=77)   # 192 tiles, weights resident
    _run_fold(1, 16, 16, 64, 64, True, L.HALO_CLAMP, L.CONV_TC, seed=78)


@pytest.mark.parametrize("epi", [0, 1])
def test_tc_conv_clamp_halo(epi):
    """AST_HALO_CLAMP: the producer of a folded conv writes the replicate halo of its (low-res) output."""
    from arbitrarystyletransfer_b200 import _lib as L, engine as E
    g = torch.Generator().manual_seed(5)
    N, H, W, cin, cout = 2, 24, 20, 128, 64
    x = bf16r(torch.randn(N, cin, H, W, generator=g))
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    exp, _, _ = oracle_conv_native(x, w, b, True, epi, "reflect")
    xin = E.nchw_to_native(x.cuda(), reflect=True)
    Ho, Wo = (H // 2, W // 2) if epi == 1 else (H, W)
    out = torch.full((N, Ho + 2, Wo + 2, cout), 7.0, device="cuda", dtype=torch.bfloat16)
    E.conv3x3(xin, E.pack_conv_weight(w.cuda()), b.cuda(), out, N=N, H=H, W=W, cin=cin, cout=cout, relu=True,
              epilogue=epi, halo=L.HALO_CLAMP, impl=L.CONV_TC)
    got = native_to_padded_nchw(out).cpu()
    torch.testing.assert_close(got, F.pad(exp, (1, 1, 1, 1), mode="replicate"), rtol=8e-3, atol=2e-3)


def test_decoder_fold_matches_unfolded(tmp_path):
    """The whole classic decoder with and without the folded upsample convs (same weights, same input)."""
    from arbitrarystyletransfer_b200 import models as M
    torch.manual_seed(1)
    dec = M.ClassicDecoder().cuda()
    f = torch.relu(torch.randn(2, 512, 12, 20, generator=torch.Generator().manual_seed(3))).cuda()
    with torch.no_grad():
        dec.fold = True
        a = dec(f).clone()
        dec.fold = False
        b = dec(f)
    assert rel_err(a.cpu(), b.cpu()) < 5e-3


@pytest.mark.parametrize("bn", [64, 128, 256, 1064, 1128, 1256])
def test_tc_conv_forced_n_block(bn):
    _run(2, 16, 32, 128, 256, True, 0, True, bn, seed=11)


@pytest.mark.parametrize("bn", [64, 128])
def test_tc_conv_resident_weights(bn):
    """Cin = 64 with one N block: the nine weight tiles stay resident in shared memory."""
    _run(3, 40, 24, 64, bn, True, 1, False, bn, seed=14)
    _run(2, 32, 48, 64, bn, True, 2, True, bn, seed=15)


def test_tc_conv_no_relu_and_taps():
    from arbitrarystyletransfer_b200 import _lib as L
    _run(1, 16, 16, 64, 64, False, 0, False, L.CONV_TC, seed=5, with_tap="pre")
    _run(1, 16, 32, 64, 128, True, 1, False, L.CONV_TC, seed=6, with_tap="pre")
    _run(1, 16, 16, 128, 64, True, 0, False, L.CONV_TC, seed=7, with_tap="post")


@pytest.mark.parametrize("epi", [0, 1, 2])
def test_direct_conv(epi):
    from arbitrarystyletransfer_b200 import _lib as L
    _run(1, 6, 10, 16, 24, True, epi, epi != 1, L.CONV_DIRECT, seed=20 + epi)
    _run(2, 8, 8, 64, 64, True, epi, False, L.CONV_DIRECT, seed=30 + epi, with_tap="pre")


def test_tc_matches_direct_bitwise_mostly():
    """Same operands, same fp32 accumulate: the two device kernels agree to fp32 re-association."""
    from arbitrarystyletransfer_b200 import _lib as L
    a, _ = _run(1, 16, 16, 64, 64, True, 0, False, L.CONV_TC, seed=9)
    b, _ = _run(1, 16, 16, 64, 64, True, 0, False, L.CONV_DIRECT, seed=9)
    assert (a != b).float().mean().item() < 0.02   # only rounding-boundary flips


@pytest.mark.parametrize("kern", list(KERNELS))
def test_many_tiles_persistent_loop(kern):
    """More tiles than SMs and > 2 tiles per CTA: exercises stage / accumulator phase wrap."""
    _run(4, 64, 96, 64, 64, True, 1, False, KERNELS[kern], seed=12)
    _run(2, 64, 64, 128, 128, True, 0, True, KERNELS[kern], seed=13)
    _run(1, 128, 160, 256, 256, True, 0, False, KERNELS[kern], seed=16)


@pytest.mark.parametrize("impl", ["tc", "tapbox", "direct"])
def test_first_and_last_layers(impl):
    """conv_1 (Normalization + 3->64 + ReLU, models.py:129-131, 198-224) and the last decoder conv
    (64->3, reflect pad, models.py:626-627).  The tensor-core variants round their operands to bf16
    (the image after normalisation / the 64->3 weights): checked tightly against that arithmetic and
    loosely (1 % of the output scale) against the fp32 reference arithmetic."""
    from arbitrarystyletransfer_b200 import _lib as L, engine as E
    from oracle import restate as R
    im = {"tc": L.CONV_TC, "tapbox": L.CONV_TC_TAPBOX, "direct": L.CONV_DIRECT}[impl]
    impl = "tc" if impl == "tapbox" else impl
    g = torch.Generator().manual_seed(1)
    N, H, W = 2, 20, 28
    img = torch.rand(N, 3, H, W, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(64, generator=g) * 0.1
    mean = torch.tensor(R.IMAGENET_MEAN).view(-1, 1, 1)
    std = torch.tensor(R.IMAGENET_STD).view(-1, 1, 1)
    xn = (img - mean) / std
    pre32 = F.conv2d(xn, w, b, padding=1)
    pre = F.conv2d(bf16r(xn), bf16r(w), b, padding=1) if impl == "tc" else pre32
    out = torch.zeros(N, H + 2, W + 2, 64, device="cuda", dtype=torch.bfloat16)
    tap = torch.empty(N, 64, H, W, device="cuda")
    img_d, w_d, b_d = img.cuda(), w.cuda(), b.cuda()   # keep the device tensors alive
    E.conv3x3_first(img_d, w_d, b_d, out, tap=tap, tap_prerelu=True, impl=im)
    torch.cuda.synchronize()
    torch.testing.assert_close(tap.cpu(), pre, rtol=1e-3, atol=2e-3)
    assert (tap.cpu() - pre32).abs().max() < 1e-2 * pre32.abs().max()
    got = native_to_padded_nchw(out).cpu()
    torch.testing.assert_close(got[:, :, 1:-1, 1:-1], bf16r(F.relu(pre)), rtol=8e-3, atol=4e-3)
    assert (got[:, :, 0] == 0).all() and (got[:, :, :, 0] == 0).all()
    E.conv3x3_first(img_d, w_d, b_d, None, tap=tap, tap_prerelu=False, impl=im)   # tap only, post-ReLU
    torch.testing.assert_close(tap.cpu(), F.relu(pre), rtol=1e-3, atol=2e-3)
    # last layer: reflect-padded native input -> NCHW fp32
    x = bf16r(torch.randn(N, 64, H, W, generator=g))
    wl = torch.randn(3, 64, 3, 3, generator=g) * 0.05
    bl = torch.randn(3, generator=g) * 0.1
    exp32 = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), wl, bl)
    exp = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), bf16r(wl), bl) if impl == "tc" else exp32
    xin = E.nchw_to_native(x.cuda(), reflect=True)
    o = torch.empty(N, 3, H, W, device="cuda")
    wl_d, bl_d = wl.cuda(), bl.cuda()
    wpk16 = E.pack_conv_weight(wl_d, cout_pad=16)
    E.conv3x3_last(xin, wl_d, wpk16, bl_d, o, clamp01=False, impl=im)
    torch.testing.assert_close(o.cpu(), exp, rtol=1e-4, atol=1e-4)
    assert (o.cpu() - exp32).abs().max() < 1e-2 * exp32.abs().max()
    E.conv3x3_last(xin, wl_d, wpk16, bl_d, o, clamp01=True, impl=im)
    torch.testing.assert_close(o.cpu(), exp.clamp(0, 1), rtol=1e-4, atol=1e-4)


def test_first_layer_tc_many_tiles_and_ragged():
    from arbitrarystyletransfer_b200 import _lib as L, engine as E
    g = torch.Generator().manual_seed(4)
    for (N, H, W) in ((3, 96, 160), (1, 37, 53)):
        img = torch.rand(N, 3, H, W, generator=g).cuda()
        w = (torch.randn(64, 3, 3, 3, generator=g) * 0.3).cuda()
        b = (torch.randn(64, generator=g) * 0.1).cuda()
        a = torch.zeros(N, H + 2, W + 2, 64, device="cuda", dtype=torch.bfloat16)
        d = torch.zeros_like(a)
        E.conv3x3_first(img, w, b, a, impl=L.CONV_TC)
        E.conv3x3_first(img, w, b, d, impl=L.CONV_DIRECT)
        torch.testing.assert_close(a.float(), d.float(), rtol=2e-2, atol=3e-2)
        assert ((a.float() - d.float()).norm() / d.float().norm()).item() < 5e-3


def test_layout_roundtrip():
    from arbitrarystyletransfer_b200 import engine as E
    g = torch.Generator().manual_seed(2)
    x = bf16r(torch.randn(2, 40, 9, 13, generator=g))
    for reflect in (False, True):
        t = E.nchw_to_native(x.cuda(), reflect=reflect)
        assert torch.equal(E.native_to_nchw(t).cpu(), x)
        full = native_to_padded_nchw(t).cpu()
        ref = F.pad(x, (1, 1, 1, 1), mode="reflect" if reflect else "constant")
        assert torch.equal(full, ref)


def test_conv_argument_errors():
    from arbitrarystyletransfer_b200 import _lib as L, engine as E
    x = torch.zeros(1, 10, 10, 48, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(9, 64, 48, device="cuda", dtype=torch.bfloat16)
    o = torch.zeros(1, 10, 10, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(L.AstError):  # Cin % 64 != 0 cannot take the tensor-core path
        E.conv3x3(x, w, None, o, N=1, H=8, W=8, cin=48, cout=64, impl=L.CONV_TC)
    with pytest.raises(L.AstError):  # reflect halo of a 1-pixel output
        E.conv3x3(x, w, None, o, N=1, H=1, W=8, cin=48, cout=64, halo=L.HALO_REFLECT, impl=L.CONV_DIRECT)
