"""oracle/restate.py against the GENUINE reference classes executed live from /root/reference
(oracle/ref_loader.py).  Only runs where the reference tree exists (the build container)."""
import pytest
import torch

from oracle import ref_loader, restate as R

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return (ref_loader.load_reference_models(), ref_loader.load_reference_module("model_util"),
            ref_loader.load_reference_module("losses"))


def _rand(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("shape", [(2, 5, 3, 7), (1, 8, 16, 16), (1, 3, 40, 24)])
def test_stats_family(ref, shape):
    M, mu, _ = ref
    c, s = _rand(1, *shape) * 2 + 1, _rand(2, *shape) + 3
    assert torch.equal(M.AdaIN()(c, s), R.adain(c, s))
    for a, b in zip(mu.channel_stats(c), R.channel_stats(c)):
        assert torch.equal(a, b)
    for a, b in zip(M.calc_mean_std(c), R.calc_mean_std(c)):
        assert torch.equal(a, b)
    assert torch.equal(M.mean_variance_norm(c), R.mean_variance_norm(c))


def test_losses(ref):
    _, _, Ls = ref
    a, b = _rand(3, 2, 4, 6, 5), _rand(4, 2, 4, 6, 5)
    assert torch.equal(Ls.compute_content_loss(a, b), R.compute_content_loss(a, b))
    assert torch.equal(Ls.gram_matrix(a), R.gram_matrix(a))
    assert torch.equal(Ls.compute_style_loss(a, b), R.compute_style_loss(a, b))


def test_decoder_spec_matches_commented_sequential():
    dec = ref_loader.build_reference_classic_decoder()
    convs = [m for m in dec if isinstance(m, torch.nn.Conv2d)]
    assert [(c.in_channels, c.out_channels) for c in convs] == [(a, b) for a, b, _, _ in R.DECODER_SPEC]
    assert len(dec) == 29
    # ReLU / Upsample placement
    relu, up = [], []
    mods = list(dec)
    for i, m in enumerate(mods):
        if isinstance(m, torch.nn.Conv2d):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            nxt2 = mods[i + 2] if i + 2 < len(mods) else None
            relu.append(isinstance(nxt, torch.nn.ReLU))
            up.append(isinstance(nxt2, torch.nn.Upsample))
    assert relu == [r for _, _, r, _ in R.DECODER_SPEC]
    assert up == [u for _, _, _, u in R.DECODER_SPEC]


def test_vgg_names_and_early_return(ref):
    M = ref[0]
    enc = M.PretrainedEncoder(['relu_9'])
    names = [l.name for l in enc._vgg_layers]
    assert names[0] == "norm" and names[1] == "conv_1" and "relu_9" in names
    convs = [l for l in enc._vgg_layers if isinstance(l, torch.nn.Conv2d)]
    assert [(c.in_channels, c.out_channels) for c in convs] == R.vgg_conv_shapes()
    assert all(c.padding == (1, 1) and c.padding_mode == "zeros" for c in convs)
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    with torch.no_grad():
        for c, w, b in zip(convs, vw, vb):
            c.weight.copy_(w)
            c.bias.copy_(b)
        x = R.rand_image(1, 32, 7)
        out = enc(x)
        assert len(out) == 1
        torch.testing.assert_close(out[0], R.vgg_relu4_1(x, vw, vb), rtol=1e-5, atol=1e-5)


def test_autoencoder_restatement_live(ref):
    """oracle/restate_ae.py vs the genuine AutoEncoder at a ragged size (40 x 24), eval and train mode."""
    from oracle import restate_ae as A
    M = ref[0]
    torch.manual_seed(5)
    ae = M.AutoEncoder()
    sd = A.make_ae_state(5)
    ref_sd = ae.state_dict()
    assert sorted(sd) == sorted(ref_sd)
    assert all(torch.equal(sd[k], ref_sd[k]) for k in sd)
    x = torch.rand(3, 3, 40, 24, generator=torch.Generator().manual_seed(9))
    ae.train()
    P = A.clone_state(sd)
    with torch.no_grad():
        assert torch.equal(ae(x), A.autoencoder_forward(P, x, training=True))
        ae.eval()
        assert torch.equal(ae(x), A.autoencoder_forward(P, x, training=False))
        dec = M.Decoder(exporting=True)
        dsd = {"decoder." + k: v for k, v in dec.state_dict().items()}
        z = torch.randn(1, 128, 3, 5, generator=torch.Generator().manual_seed(3))
        assert torch.equal(dec(z), A.decoder_forward(dsd, z, exporting=True))


def test_adaattn_ast_and_hist_restatements_live(ref):
    """oracle/restate_attn.py vs the genuine AdaAttN / AST (ada_att_2 / ada_out restored, models.py:407, 410) at a
    ragged size, and oracle/restate.py::compute_hist_loss vs the genuine losses.compute_hist_loss."""
    from oracle import restate as R, restate_ae as A, restate_attn as T
    from oracle.make_golden import restore_ast
    M, _, Ls = ref[0], ref[1], ref[2]
    torch.manual_seed(13)
    layer = M.AdaAttN(24)
    P = {f"a.{n}.weight": getattr(layer, n).weight.detach().clone() for n in ("W_q", "W_k", "W_v")}
    g = torch.Generator().manual_seed(4)
    c, s = torch.randn(2, 24, 5, 9, generator=g) * 2, torch.randn(2, 24, 7, 3, generator=g) + 1
    with torch.no_grad():
        torch.testing.assert_close(T.adaattn(P, "a", c, s), layer(c, s), rtol=1e-5, atol=1e-5)
    sd = A.activate_gates(T.make_ast_state(7))
    torch.manual_seed(1)
    ast = restore_ast(M, M.AST())
    ast.load_state_dict(sd, strict=True)
    ast.train()
    ci, si = R.rand_image(2, 48, 21), R.rand_image(2, 48, 22)
    with torch.no_grad():
        want = ast(ci, si, alpha=0.5)
        got = T.ast_forward(A.clone_state(sd), ci, si, alpha=0.5)
    for a, b in zip(got, want):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    x, y = torch.rand(2, 3, 12, 10, generator=g) * 1.2 - 0.1, torch.rand(2, 3, 12, 10, generator=g)
    assert R.compute_hist_loss(x, y).item() == Ls.compute_hist_loss(x, y).item()
