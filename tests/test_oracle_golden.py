"""The CPU oracle (oracle/restate.py) against golden vectors produced by the GENUINE reference code
(oracle/make_golden.py, executed in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import restate as R

CASES = ["a", "b", "c", "d"]


def T(a):
    return torch.from_numpy(np.asarray(a))


def same(a, b):
    """bit-exact, NaN == NaN (dead channels give 0/0 in the reference: models.py:47)."""
    return np.array_equal(a.detach().numpy(), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("k", CASES)
def test_channel_stats_and_adain(golden_stats, k):
    g = golden_stats
    c, s = T(g[f"adain_{k}_content"]), T(g[f"adain_{k}_style"])
    m, sd = R.channel_stats(c)
    assert same(m, g[f"adain_{k}_cmean"])
    assert same(sd, g[f"adain_{k}_cstd"])
    assert same(R.adain(c, s), g[f"adain_{k}_out"])
    # fp64 numpy restatement agrees with the fp32 reference to fp32 rounding
    m64, sd64 = R.channel_stats_np(c.numpy())
    np.testing.assert_allclose(m64, g[f"adain_{k}_cmean"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(sd64, g[f"adain_{k}_cstd"], rtol=2e-6, atol=1e-6)
    # K = 1 multi-style form == the reference AdaIN (models.py:43-51), up to fp32 re-association
    np.testing.assert_allclose(R.adain_multi(c, [s], [1.0]).numpy(), g[f"adain_{k}_out"],
                               rtol=1e-5, atol=1e-5, equal_nan=True)
    np.testing.assert_allclose(R.adain_multi(c, [s], [1.0], alpha=0.6).numpy(),
                               g[f"adain_{k}_blend06"], rtol=1e-5, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize("k", CASES)
def test_calc_mean_std_mvn(golden_stats, k):
    g = golden_stats
    c = T(g[f"adain_{k}_content"])
    m, sd = R.calc_mean_std(c)
    assert same(m, g[f"adain_{k}_cms_mean"])
    assert same(sd, g[f"adain_{k}_cms_std"])
    assert same(R.mean_variance_norm(c), g[f"adain_{k}_mvn"])


def test_dead_channel_is_nan_like_reference(golden_stats):
    g = golden_stats
    out = R.adain(T(g["adain_dead_content"]), T(g["adain_dead_style"])).numpy()
    ref = g["adain_dead_out"]
    assert np.isnan(ref[:, 1]).all()            # the reference has no epsilon: 0/0
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    np.testing.assert_array_equal(out[~np.isnan(out)], ref[~np.isnan(ref)])


def test_swapped_statistics(golden_stats):
    """models.py:44 binds (style_std, style_mean) = channel_stats(style) = (mean, std)."""
    g = golden_stats
    c, s = T(g["adain_b_content"]), T(g["adain_b_style"])
    out = T(g["adain_b_out"])
    m, sd = R.channel_stats(out)
    sm, ssd = R.channel_stats(s)
    # per-channel std of the output = |mean(style)|, per-channel mean = std(style)
    torch.testing.assert_close(sd, sm.abs(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(m, ssd, rtol=1e-4, atol=1e-4)
    assert not torch.allclose(out, R.adain_canonical(c, s), atol=1e-2)


def test_backward_goldens(golden_stats):
    g = golden_stats
    x = T(g["mvn_bwd_x"]).clone().requires_grad_(True)
    R.mean_variance_norm(x).backward(T(g["mvn_bwd_gy"]))
    torch.testing.assert_close(x.grad, T(g["mvn_bwd_gx"]), rtol=1e-6, atol=1e-6)
    x2 = T(g["mvn_bwd_x"]).clone().requires_grad_(True)
    m, sd = R.channel_stats(x2)
    ((m * T(g["cs_bwd_gm"])).sum() + (sd * T(g["cs_bwd_gs"])).sum()).backward()
    torch.testing.assert_close(x2.grad, T(g["cs_bwd_gx"]), rtol=1e-6, atol=1e-6)


def test_losses(golden_losses):
    g = golden_losses
    a, b = T(g["loss_a"]).clone().requires_grad_(True), T(g["loss_b"])
    l = R.compute_content_loss(a, b)
    assert l.item() == pytest.approx(float(g["content_loss"]), rel=1e-6)
    assert l.item() == pytest.approx(R.huber_np(a.detach().numpy(), b.numpy()), rel=1e-5)
    l.backward()
    torch.testing.assert_close(a.grad, T(g["content_loss_ga"]))
    a.grad = None
    gm = R.gram_matrix(a)
    torch.testing.assert_close(gm.detach(), T(g["gram_a"]), rtol=1e-6, atol=1e-7)
    (gm * T(g["gram_gg"])).sum().backward()
    torch.testing.assert_close(a.grad, T(g["gram_ga"]), rtol=1e-5, atol=1e-7)
    a.grad = None
    l = R.compute_style_loss(a, b)
    assert l.item() == pytest.approx(float(g["style_loss"]), rel=1e-6)
    l.backward()
    torch.testing.assert_close(a.grad, T(g["style_loss_ga"]), rtol=1e-5, atol=1e-8)
    a.grad = None
    l = R.tv_loss(a)
    assert l.item() == pytest.approx(float(g["tv_loss"]), rel=1e-6)
    (l * 0.37).backward()
    torch.testing.assert_close(a.grad, T(g["tv_loss_ga"]), rtol=1e-5, atol=1e-6)


def test_vgg_taps_and_decoder(golden_networks):
    g = golden_networks
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)
    with torch.no_grad():
        taps = R.vgg_forward(T(g["vgg_x32"]), vw, vb)
        assert len(taps) == 6
        for i, t in enumerate(taps):
            torch.testing.assert_close(t, T(g[f"vgg_x32_tap{i}"]), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(R.decoder_forward(T(g["dec_in"]), dw, db), T(g["dec_out"]),
                                   rtol=1e-5, atol=1e-5)


def test_stylize_64(golden_networks):
    g = golden_networks
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)
    c, s = R.rand_image(1, 64, 101), R.rand_image(1, 64, 102)
    with torch.no_grad():
        fc = R.vgg_relu4_1(c, vw, vb)
        torch.testing.assert_close(fc, T(g["s64_fc"]), rtol=1e-5, atol=1e-5)
        img = R.stylize(c, s, vw, vb, dw, db)
    assert torch.isfinite(img).all()
    assert R.psnr(img, T(g["s64_img"])) > 80.0


def test_hist_loss_restatement_matches_reference_golden():
    """compute_hist_loss (losses.py:8-87): value and both input gradients, bit for bit in one chunk and to fp32
    round-off when the element axis is walked in chunks."""
    from tests.conftest import load_golden
    g = load_golden("hist")
    for tag in "ab":
        x = T(g[f"hist_{tag}_x"]).clone().requires_grad_(True)
        y = T(g[f"hist_{tag}_y"]).clone().requires_grad_(True)
        l = R.compute_hist_loss(x, y)
        assert l.item() == float(g[f"hist_{tag}_loss"])
        (l * 0.5).backward()
        torch.testing.assert_close(x.grad, T(g[f"hist_{tag}_gx"]), rtol=0, atol=0)
        torch.testing.assert_close(y.grad, T(g[f"hist_{tag}_gy"]), rtol=0, atol=0)
        h = R.soft_histogram(x.detach(), chunk=100)
        torch.testing.assert_close(h, R.soft_histogram(x.detach()), rtol=1e-5, atol=1e-6)
        # N = C*H (losses.py:54), not C*H*W: the bins of one image sum to W
        assert h.sum(1).mean().item() == pytest.approx(x.shape[3] * ((x.detach() > 0.02) & (x.detach() < 0.98)).float().mean().item(), rel=0.1)


# ---- the oracle at the benchmarked sizes (fixtures made by the genuine reference: make_golden.golden_bigsizes) ----
@pytest.fixture(scope="module")
def classic_weights():
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)
    return vw, vb, dw, db


def _views(img):
    S = img.shape[2]
    a = S // 2 - 32
    return {"crop": img[:, :, a:a + 64, a:a + 64], "corner": img[:, :, :48, S - 48:], "sub8": img[:, :, ::8, ::8]}


def test_oracle_config4_512_vs_reference(golden_big, classic_weights):
    vw, vb, dw, db = classic_weights
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        img = R.stylize(R.rand_image(1, 512, 401), R.rand_image(1, 512, 402), vw, vb, dw, db)
    for k, v in _views(img).items():
        np.testing.assert_allclose(v.numpy(), golden_big[f"cfg4_img_{k}"], rtol=1e-4, atol=1e-5)


def test_oracle_config5_2048_vs_reference(golden_big, classic_weights):
    """The K-style mix of the oracle (adain_multi: one affine with mixed statistics) against the fixture built from
    the reference's own AdaIN applied per style and summed with the weights; 2048x2048, about 40 s of CPU."""
    vw, vb, dw, db = classic_weights
    c = R.rand_image(1, 2048, 501)
    styles = R.rand_image(4, 2048, 502)
    w = [0.4, 0.3, 0.2, 0.1]
    with torch.no_grad():
        fc = R.vgg_relu4_1(c, vw, vb)
        fs = [R.vgg_relu4_1(styles[k:k + 1], vw, vb) for k in range(4)]
        np.testing.assert_allclose(fc[:, ::16, ::4, ::4].numpy(), golden_big["cfg5_a10_fc_sub"], rtol=1e-4, atol=1e-5)
        t = R.adain_multi(fc, fs, w, 0.6)
        img = R.decoder_forward(t, dw, db)
    for k, v in _views(img).items():
        np.testing.assert_allclose(v.numpy(), golden_big[f"cfg5_a06_img_{k}"], rtol=2e-4, atol=2e-5)
