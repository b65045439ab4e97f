"""End-to-end parity of the classic VGG relu4_1 -> AdaIN -> decoder path (configs 1, 4-shaped, 5-shaped)
against golden outputs of the genuine reference and the CPU oracle.  north_star tolerance for the
bf16 pipeline: relative L2 <= 1e-2 and PSNR >= 40 dB (range = max(ref) - min(ref))."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def weights():
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)
    return vw, vb, dw, db


@pytest.fixture(scope="module")
def engine(weights):
    from arbitrarystyletransfer_b200.engine import StyleTransferEngine
    vw, vb, dw, db = weights
    return StyleTransferEngine(vw[:9], vb[:9], dw, db, device="cuda")


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_stylize_64_vs_reference_golden(engine, golden_networks):
    g = golden_networks
    c, s = R.rand_image(1, 64, 101).cuda(), R.rand_image(1, 64, 102).cuda()
    from arbitrarystyletransfer_b200 import engine as E
    fc = E.native_to_nchw(engine.encode(c)).cpu()
    # relu4_1 after 9 bf16-stored layers: ~3 % relative L2 (the north_star bound applies to the image)
    assert rel_l2(fc, T(g["s64_fc"])) < 5e-2
    img = engine.stylize(c, s).cpu()
    ref = T(g["s64_img"])
    assert torch.isfinite(img).all()
    assert rel_l2(img, ref) < 1e-2
    assert R.psnr(img, ref) >= 40.0


def test_config1_256_vs_reference_golden(engine, golden_networks):
    g = golden_networks
    c, s = R.rand_image(1, 256, 101).cuda(), R.rand_image(1, 256, 102).cuda()
    img = engine.stylize(c, s, alpha=1.0).cpu()
    crop, sub = T(g["cfg1_img_crop"]), T(g["cfg1_img_sub4"])
    assert R.psnr(img[:, :, 96:160, 96:160], crop) >= 40.0
    assert R.psnr(img[:, :, ::4, ::4], sub) >= 40.0
    assert rel_l2(img[:, :, ::4, ::4], sub) < 1e-2
    st = g["cfg1_img_stats"]
    assert img.mean().item() == pytest.approx(st[0], abs=2e-3)
    assert img.std().item() == pytest.approx(st[1], rel=2e-2)


@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_batch_and_alpha_vs_oracle(engine, weights, alpha):
    vw, vb, dw, db = weights
    c, s = R.rand_image(3, 96, 401), R.rand_image(3, 96, 402)     # ragged: 96/8 = 12 (partial tiles)
    with torch.no_grad():
        ref = R.stylize(c, s, vw, vb, dw, db, alpha=alpha)
    img = engine.stylize(c.cuda(), s.cuda(), alpha=alpha).cpu()
    assert R.psnr(img, ref) >= 40.0 and rel_l2(img, ref) < 1e-2


def test_multi_style_interpolation_vs_oracle(engine, weights):
    vw, vb, dw, db = weights
    c = R.rand_image(1, 128, 501)
    styles = [R.rand_image(1, 128, 502 + k) for k in range(4)]
    w = [0.4, 0.3, 0.2, 0.1]
    with torch.no_grad():
        ref = R.stylize(c, styles, vw, vb, dw, db, alpha=0.6, style_weights=w)
    img = engine.stylize(c.cuda(), [s.cuda() for s in styles], alpha=0.6, style_weights=w).cpu()
    assert R.psnr(img, ref) >= 40.0 and rel_l2(img, ref) < 1e-2


def _check_big(img, g, tag):
    """img (1,3,S,S) fp32 on the CPU vs the reference-made crop / corner / stride-8 subsample: north_star bar."""
    S = img.shape[2]
    a = S // 2 - 32
    views = {"crop": img[:, :, a:a + 64, a:a + 64], "corner": img[:, :, :48, S - 48:], "sub8": img[:, :, ::8, ::8]}
    out = {}
    for k, v in views.items():
        ref = T(g[f"{tag}_img_{k}"])
        out[k] = (R.psnr(v, ref), rel_l2(v, ref))
    print(f"{tag}: " + ", ".join(f"{k} PSNR {p:.1f} dB rel {r:.2e}" for k, (p, r) in out.items()))
    for k, (p, r) in out.items():
        assert p >= 40.0 and r <= 1e-2, (tag, k, p, r)
    st = g[f"{tag}_img_stats"]
    assert img.mean().item() == pytest.approx(st[0], abs=2e-3)
    assert img.std().item() == pytest.approx(st[1], rel=2e-2)


def test_config4_512_vs_reference_golden(engine, golden_big):
    """BASELINE config 4 AT ITS OWN SIZE (the benchmarked one): one 512x512 pair, seeds 401 / 402, against the output
    of the genuine reference modules (oracle/make_golden.py::golden_bigsizes)."""
    from arbitrarystyletransfer_b200 import engine as E
    c, s = R.rand_image(1, 512, 401).cuda(), R.rand_image(1, 512, 402).cuda()
    fc = E.native_to_nchw(engine.encode(c)).cpu()
    assert rel_l2(fc[:, ::16, ::4, ::4], T(golden_big["cfg4_fc_sub"])) < 5e-2
    _check_big(engine.stylize(c, s, alpha=1.0).cpu(), golden_big, "cfg4")
    # the same pair inside a batch of 3 (the bench runs 32 per GPU): position in the batch must not matter
    cb = torch.cat([R.rand_image(1, 512, 7).cuda(), c, R.rand_image(1, 512, 8).cuda()])
    sb = torch.cat([R.rand_image(1, 512, 9).cuda(), s, R.rand_image(1, 512, 10).cuda()])
    _check_big(engine.stylize(cb, sb, alpha=1.0)[1:2].cpu(), golden_big, "cfg4")


@pytest.mark.parametrize("alpha,tag", [(1.0, "cfg5_a10"), (0.6, "cfg5_a06")])
def test_config5_2048_vs_reference_golden(engine, golden_big, alpha, tag):
    """BASELINE config 5 at its own size: 2048x2048 content (seed 501), four 2048x2048 styles (seed 502), weights
    (.4,.3,.2,.1), alpha 1.0 / 0.6, against the genuine reference pieces."""
    c = R.rand_image(1, 2048, 501).cuda()
    styles = R.rand_image(4, 2048, 502)
    w = [0.4, 0.3, 0.2, 0.1]
    img = engine.stylize(c, [styles[k:k + 1].cuda() for k in range(4)], alpha=alpha, style_weights=w).cpu()
    _check_big(img, golden_big, tag)


def test_size_independent_properties_512(engine):
    """Full-size (512x512) checks that need no oracle run: batch-order equivariance, alpha = 0
    reproduces decoder(relu4_1(content)) regardless of the style, determinism."""
    c, s = R.rand_image(2, 512, 11).cuda(), R.rand_image(2, 512, 12).cuda()
    a = engine.stylize(c, s).clone()
    b = engine.stylize(c.flip(0), s.flip(0)).flip(0)
    assert torch.equal(a, b)
    assert torch.equal(a, engine.stylize(c, s))
    z1 = engine.stylize(c, s, alpha=0.0).clone()
    z2 = engine.stylize(c, s.flip(0), alpha=0.0)
    torch.testing.assert_close(z1, z2, rtol=0, atol=0)


def test_modules_drop_in(weights, golden_networks):
    """The nn.Module surface: state-dict keys of the reference, taps by name, NCHW fp32 in/out."""
    from arbitrarystyletransfer_b200 import models as M
    vw, vb, dw, db = weights
    g = golden_networks
    enc = M.PretrainedEncoder().cuda()
    assert sorted(enc.state_dict().keys()) == list(g["vgg_state_keys"])
    dec = M.ClassicDecoder().cuda()
    assert sorted(dec.state_dict().keys()) == list(g["dec_state_keys"])
    with torch.no_grad():
        for conv, w, b in zip(enc._convs(), vw, vb):
            conv.weight.copy_(w); conv.bias.copy_(b)
        for conv, w, b in zip(dec._convs(), dw, db):
            conv.weight.copy_(w); conv.bias.copy_(b)
        taps = enc(T(g["vgg_x32"]).cuda())
    assert len(taps) == 6
    for i, t in enumerate(taps):
        ref = T(g[f"vgg_x32_tap{i}"])
        assert t.shape == ref.shape and t.dtype == torch.float32
        assert rel_l2(t.cpu(), ref) < 5e-2, f"tap {i}"   # bf16-stored layers upstream of the tap
    with torch.no_grad():
        out = dec(T(g["dec_in"]).cuda()).cpu()
    assert R.psnr(out, T(g["dec_out"])) >= 40.0
    # full net through the module API
    enc9 = M.PretrainedEncoder(['relu_9']).cuda()
    enc9.load_state_dict(enc.state_dict())
    net = M.StyleTransferNet(enc9, dec)
    img = net(R.rand_image(1, 64, 101).cuda(), R.rand_image(1, 64, 102).cuda()).cpu()
    assert R.psnr(img, T(g["s64_img"])) >= 40.0


def test_config5_2048_four_style_interpolation_properties(engine):
    """BASELINE config 5 at full size (2048x2048, 4-style interpolation weights): properties that need
    no CPU run of 12 TFLOP -- finiteness, linearity of the interpolation in the style statistics
    (four copies of one style with any weights == that single style), alpha = 0 independence of the
    styles, and the K1 kernel agreeing with the in-pipeline AdaIN on the 256x256 relu4_1 maps."""
    from arbitrarystyletransfer_b200 import engine as E, functional as Fn
    c = R.rand_image(1, 2048, 501).cuda()
    styles = [R.rand_image(1, 2048, 502 + k).cuda() for k in range(4)]
    w = [0.4, 0.3, 0.2, 0.1]
    out = engine.stylize(c, styles, alpha=1.0, style_weights=w).clone()
    assert out.shape == (1, 3, 2048, 2048) and torch.isfinite(out).all()
    same = engine.stylize(c, [styles[0]] * 4, alpha=1.0, style_weights=w).clone()
    single = engine.stylize(c, [styles[0]], alpha=1.0, style_weights=[1.0])
    assert R.psnr(same.cpu(), single.cpu()) >= 60.0
    z1 = engine.stylize(c, styles, alpha=0.0, style_weights=w).clone()
    z2 = engine.stylize(c, styles[::-1], alpha=0.0, style_weights=w)
    torch.testing.assert_close(z1, z2, rtol=0, atol=0)
    # relu4_1 maps (1, 512, 256, 256): rows of 65 536 elements -> cluster-split K1 path in fp32
    fc = E.native_to_nchw(engine.encode(c, "c"))
    fs = [E.native_to_nchw(engine.encode(s, f"s{k}")) for k, s in enumerate(styles)]
    t_k1 = Fn.adain_forward(fc, fs, w, alpha=0.6)
    t_native = E.native_to_nchw(engine.adain(engine.encode(c, "c"),
                                             [engine.encode(s, f"s{k}") for k, s in enumerate(styles)], w, alpha=0.6))
    assert rel_l2(t_native, t_k1) < 1e-2      # bf16 output rounding of the native path
    ref = R.adain_multi(fc.cpu(), [f.cpu() for f in fs], w, alpha=0.6)
    torch.testing.assert_close(t_k1.cpu(), ref, rtol=1e-5, atol=1e-5 * ref.abs().max().item())


def test_host_pipeline_matches_direct_calls(engine):
    """engine.HostPipeline (pinned host in/out, overlapped copies) returns exactly what direct
    stylize() calls return, for more steps than slots and distinct batches per step."""
    from arbitrarystyletransfer_b200.engine import HostPipeline
    N, S, steps = 2, 96, 5
    pipe = HostPipeline(engine, N, S, S)
    cs = [R.rand_image(N, S, 600 + i).pin_memory() for i in range(steps)]
    ss = [R.rand_image(N, S, 700 + i).pin_memory() for i in range(steps)]
    outs = [torch.empty(N, 3, S, S).pin_memory() for _ in range(steps)]
    for i in range(steps):
        pipe.submit(cs[i], ss[i], outs[i])
    pipe.synchronize()
    for i in range(steps):
        ref = engine.stylize(cs[i].cuda(), ss[i].cuda()).cpu()
        assert torch.equal(outs[i], ref), f"step {i}"
    with pytest.raises(Exception):
        pipe.submit(torch.zeros(N, 3, S, S), ss[0], outs[0])      # not pinned


# ---- byte images at the boundary -------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 5, 7), (3, 96, 160), (1, 3, 3)])
def test_u8_converters_bit_exact(shape):
    """ast_u8hwc_to_nchw == transforms.ToTensor() (u8 / 255, data_loader.py:114); ast_nchw_to_u8hwc ==
    Hardtanh(0,1) then transforms.ToPILImage() (mul(255).byte(), train.py:18); ragged sizes take the scalar path."""
    from arbitrarystyletransfer_b200 import engine as E
    N, H, W = shape
    g = torch.Generator().manual_seed(H * W)
    u = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8)
    ref = u.permute(0, 3, 1, 2).float().div(255)
    got = E.u8_to_nchw(u.cuda()).cpu()
    assert torch.equal(got, ref)
    x = torch.rand(N, 3, H, W, generator=g) * 1.4 - 0.2
    x[0, 0, 0, 0] = float("nan")
    x[0, 1, 0, 0] = 1.0
    x[0, 2, 0, 0] = 255.0 / 255.0 - 1e-8
    want = torch.nan_to_num(x, nan=0.0).clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous()
    assert torch.equal(E.nchw_to_u8(x.cuda()).cpu(), want)
    # round trip of every byte value
    allv = torch.arange(256, dtype=torch.uint8).view(1, 16, 16, 1).expand(1, 16, 16, 3).contiguous()
    assert torch.equal(E.nchw_to_u8(E.u8_to_nchw(allv.cuda())).cpu(), allv)


def test_stylize_u8_equals_float_path_and_pipeline(engine):
    from arbitrarystyletransfer_b200 import engine as E
    from arbitrarystyletransfer_b200.engine import HostPipeline
    g = torch.Generator().manual_seed(9)
    N, S = 2, 96
    cu = torch.randint(0, 256, (N, S, S, 3), generator=g, dtype=torch.uint8)
    su = torch.randint(0, 256, (N, S, S, 3), generator=g, dtype=torch.uint8)
    out = engine.stylize_u8(cu.cuda(), su.cuda(), alpha=0.8).cpu()
    ref = E.nchw_to_u8(engine.stylize(E.u8_to_nchw(cu.cuda()), E.u8_to_nchw(su.cuda()), alpha=0.8)).cpu()
    assert out.dtype == torch.uint8 and out.shape == (N, S, S, 3) and torch.equal(out, ref)
    pipe = HostPipeline(engine, N, S, S, dtype="u8")
    outs = [torch.empty(N, S, S, 3, dtype=torch.uint8).pin_memory() for _ in range(3)]
    for o in outs:
        pipe.submit(cu.pin_memory(), su.pin_memory(), o, alpha=0.8)
    pipe.synchronize()
    for o in outs:
        assert torch.equal(o, ref)
    with pytest.raises(Exception):
        pipe.submit(torch.zeros(N, 3, S, S).pin_memory(), su.pin_memory(), outs[0])     # fp32 into a u8 pipeline


def test_adain_native_styles_of_different_sizes(engine, weights):
    """every style map carries its own size through ast_adain_native_fwd."""
    vw, vb, dw, db = weights
    c = R.rand_image(1, 64, 901)
    styles = [R.rand_image(1, 96, 902), R.rand_image(1, 64, 903, w=128)]
    w = [0.7, 0.3]
    with torch.no_grad():
        ref = R.stylize(c, styles, vw, vb, dw, db, alpha=1.0, style_weights=w)
    img = engine.stylize(c.cuda(), [s.cuda() for s in styles], alpha=1.0, style_weights=w).cpu()
    assert R.psnr(img, ref) >= 40.0 and rel_l2(img, ref) < 1e-2
