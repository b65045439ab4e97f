"""K1 parity: fused AdaIN / channel statistics / MVN kernels (through the C ABI) against the CPU
oracle and the golden vectors of the genuine reference.  north_star tolerance: AdaIN statistics and
output within 1e-5 relative in fp32 (written below as rtol=1e-5 plus an absolute floor of 1e-5 x the
tensor's scale for values that cancel to ~0)."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, ref, rtol=RTOL, scale=None):
    ref = ref if isinstance(ref, torch.Tensor) else T(ref)
    scale = float(ref[torch.isfinite(ref)].abs().max()) if scale is None else scale
    torch.testing.assert_close(got.cpu(), ref, rtol=rtol, atol=rtol * max(scale, 1e-30), equal_nan=True)


@pytest.mark.parametrize("k", ["a", "b", "c", "d"])
def test_golden_adain_family(golden_stats, k):
    from arbitrarystyletransfer_b200 import models as M
    g = golden_stats
    c, s = T(g[f"adain_{k}_content"]).cuda(), T(g[f"adain_{k}_style"]).cuda()
    close(M.AdaIN()(c, s), g[f"adain_{k}_out"])
    close(M.AdaIN()(c, s, alpha=0.6), g[f"adain_{k}_blend06"])
    m, sd = M.channel_stats(c)
    close(m, g[f"adain_{k}_cmean"]); close(sd, g[f"adain_{k}_cstd"])
    m, sd = M.calc_mean_std(c)
    close(m, g[f"adain_{k}_cms_mean"]); close(sd, g[f"adain_{k}_cms_std"])
    close(M.mean_variance_norm(c), g[f"adain_{k}_mvn"])


def test_dead_channel_nan_parity(golden_stats):
    from arbitrarystyletransfer_b200 import models as M
    g = golden_stats
    out = M.AdaIN()(T(g["adain_dead_content"]).cuda(), T(g["adain_dead_style"]).cuda()).cpu().numpy()
    assert np.array_equal(np.isnan(out), np.isnan(g["adain_dead_out"]))
    np.testing.assert_allclose(out[~np.isnan(out)], g["adain_dead_out"][~np.isnan(out)], rtol=1e-5, atol=1e-5)


SHAPES = [
    (1, 512, 32, 32),     # config 1
    (2, 512, 64, 64),     # config 4 (per-image shape)
    (1, 16, 256, 256),    # config 5 row length (65 536 elements): cluster-split path
    (3, 7, 5, 3),         # odd everything: scalar path
    (2, 3, 33, 31),       # HW % 4 != 0
    (1, 4, 1, 2),         # HW = 2
    (1, 2, 300, 300),     # longer than the register-cached limit? (90 000) generic two-pass path
    (1, 1, 640, 512),     # 327 680 elements: beyond 8 CTAs x 8 x 256 vectors
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("alpha", [1.0, 0.6])
def test_adain_vs_oracle(shape, alpha):
    from arbitrarystyletransfer_b200 import functional as Fn
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    c = torch.relu(torch.randn(*shape, generator=g) * 3 + 1)
    s = torch.randn(*shape, generator=g) * 2 + 3
    ref = R.adain_multi(c, [s], [1.0], alpha=alpha)
    out, stats = Fn.adain_forward(c.cuda(), [s.cuda()], [1.0], alpha, return_stats=True)
    close(out, ref)
    m64, sd64 = R.channel_stats_np(c.numpy())
    st = stats.cpu().double().numpy().reshape(shape[0], shape[1], 4)
    np.testing.assert_allclose(st[..., 0], m64[..., 0, 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st[..., 1], sd64[..., 0, 0], rtol=1e-5, atol=1e-6)
    sm64, ssd64 = R.channel_stats_np(s.numpy())
    np.testing.assert_allclose(st[..., 2], sm64[..., 0, 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st[..., 3], ssd64[..., 0, 0], rtol=1e-5, atol=1e-6)


def test_large_mean_small_std_is_stable():
    """Welford, not sum-of-squares: with mean 1000 and std 0.01 the fp32 INPUT is already quantised
    to 6e-5 (ulp of 1000), so 1e-5 on the std is not attainable in fp32 by any method; a naive
    sum-of-squares would lose every digit here.  Bar: 1e-3 relative vs the fp64 restatement."""
    from arbitrarystyletransfer_b200 import functional as Fn
    g = torch.Generator().manual_seed(5)
    c = (torch.randn(2, 8, 64, 64, generator=g) * 0.01 + 1000.0)
    m, sd = Fn.channel_stats_flat(c.cuda())
    m64, sd64 = R.channel_stats_np(c.numpy())
    np.testing.assert_allclose(sd.cpu().numpy(), sd64[..., 0, 0], rtol=1e-3)
    np.testing.assert_allclose(m.cpu().numpy(), m64[..., 0, 0], rtol=1e-6)


def test_multi_style_interpolation_cfg5_shape():
    """BASELINE config 5: 4-style interpolation weights on a long-row map."""
    from arbitrarystyletransfer_b200 import models as M
    g = torch.Generator().manual_seed(9)
    c = torch.relu(torch.randn(1, 8, 256, 256, generator=g) * 3 + 1)
    styles = [torch.relu(torch.randn(1, 8, 256, 256, generator=g) * (k + 1) + k) for k in range(4)]
    w = [0.4, 0.3, 0.2, 0.1]
    for alpha in (1.0, 0.6):
        for canonical in (False, True):
            ref = R.adain_multi(c, styles, w, alpha=alpha, canonical=canonical)
            out = M.AdaIN(canonical)(c.cuda(), [s.cuda() for s in styles], alpha=alpha, style_weights=w)
            close(out, ref)


def test_properties():
    from arbitrarystyletransfer_b200 import models as M, functional as Fn
    g = torch.Generator().manual_seed(3)
    c = (torch.randn(2, 6, 20, 24, generator=g) * 2 + 1).cuda()
    s = (torch.randn(2, 6, 20, 24, generator=g).abs() + 2).cuda()
    # alpha = 0 is the identity on the content features (models.py:471)
    close(M.AdaIN()(c, s, alpha=0.0), c.cpu())
    # reference mode: output mean = std(style), output std = |mean(style)|  (models.py:44 swap)
    out = M.AdaIN()(c, s)
    om, osd = Fn.channel_stats_flat(out)
    sm, ssd = Fn.channel_stats_flat(s)
    close(om, ssd.cpu(), rtol=1e-4); close(osd, sm.abs().cpu(), rtol=1e-4)
    # canonical mode: output stats = style stats
    out = M.AdaIN(canonical=True)(c, s)
    om, osd = Fn.channel_stats_flat(out)
    close(om, sm.cpu(), rtol=1e-4); close(osd, ssd.cpu(), rtol=1e-4)
    # in-place (out aliases content) gives the same result
    c2 = c.clone()
    Fn.adain_forward(c2, [s], out=c2)
    close(c2, M.AdaIN()(c, s).cpu())


def test_bf16_io():
    from arbitrarystyletransfer_b200 import functional as Fn
    g = torch.Generator().manual_seed(4)
    c = torch.relu(torch.randn(2, 16, 32, 32, generator=g) * 3 + 1).bfloat16()
    s = (torch.randn(2, 16, 32, 32, generator=g) * 2 + 3).bfloat16()
    ref = R.adain(c.float(), s.float())
    out = Fn.adain_forward(c.cuda(), [s.cuda()])
    assert out.dtype == torch.bfloat16
    torch.testing.assert_close(out.float().cpu(), ref, rtol=1e-2, atol=2e-2)


def test_backward_goldens(golden_stats):
    from arbitrarystyletransfer_b200 import models as M
    g = golden_stats
    x = T(g["mvn_bwd_x"]).cuda().requires_grad_(True)
    M.mean_variance_norm(x).backward(T(g["mvn_bwd_gy"]).cuda())
    close(x.grad, g["mvn_bwd_gx"], rtol=1e-4)
    x2 = T(g["mvn_bwd_x"]).cuda().requires_grad_(True)
    m, sd = M.channel_stats(x2)
    ((m * T(g["cs_bwd_gm"]).cuda()).sum() + (sd * T(g["cs_bwd_gs"]).cuda()).sum()).backward()
    close(x2.grad, g["cs_bwd_gx"], rtol=1e-4)


def test_adain_autograd_matches_oracle():
    from arbitrarystyletransfer_b200 import models as M
    g = torch.Generator().manual_seed(8)
    c = (torch.randn(2, 5, 9, 11, generator=g) * 2 + 1)
    s = (torch.randn(2, 5, 7, 6, generator=g) + 3)
    go = torch.randn(2, 5, 9, 11, generator=g)
    cr, sr = c.clone().requires_grad_(True), s.clone().requires_grad_(True)
    R.alpha_blend(R.adain(cr, sr), cr, 0.7).backward(go)
    cg, sg = c.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    out = M.AdaIN()(cg, sg, alpha=0.7)
    out.backward(go.cuda())
    close(out.detach(), R.alpha_blend(R.adain(c, s), c, 0.7))
    close(cg.grad, cr.grad, rtol=1e-4); close(sg.grad, sr.grad, rtol=1e-4)


@pytest.mark.parametrize("canonical", [False, True])
def test_adain_autograd_two_styles_weights_and_partial_grads(canonical):
    """The kernel-only backward (ast_adain_bwd + ast_channel_stats_bwd) against torch autograd through the oracle's
    formula: K = 2 styles of different sizes, mix weights, alpha blend, both bindings of (A, B); then with only the
    styles / only the content requiring grad."""
    from arbitrarystyletransfer_b200 import models as M
    g = torch.Generator().manual_seed(18)
    c = torch.randn(2, 6, 8, 12, generator=g) * 1.5 + 0.5
    s1 = torch.randn(2, 6, 5, 7, generator=g) * 0.7 + 2
    s2 = torch.randn(2, 6, 9, 4, generator=g) * 1.2 - 1
    go = torch.randn(2, 6, 8, 12, generator=g)
    w, alpha = [0.3, 0.7], 0.6

    def ref(c_, a_, b_):
        mu = c_.mean(dim=(2, 3), keepdim=True)
        sd = c_.std(dim=(2, 3), keepdim=True)
        A = B = 0
        for wk, sk in zip(w, (a_, b_)):
            m, d = sk.mean(dim=(2, 3), keepdim=True), sk.std(dim=(2, 3), keepdim=True)
            A = A + wk * (d if canonical else m)      # models.py:44 binds them swapped; canonical = the paper's form
            B = B + wk * (m if canonical else d)
        t = (c_ - mu) / sd * A + B
        return alpha * t + (1 - alpha) * c_

    cr, ar, br = (t.clone().double().requires_grad_(True) for t in (c, s1, s2))
    ref(cr, ar, br).backward(go.double())
    cg, ag, bg = (t.cuda().requires_grad_(True) for t in (c, s1, s2))
    layer = M.AdaIN()
    layer.canonical = canonical
    out = layer(cg, [ag, bg], alpha=alpha, style_weights=w)
    out.backward(go.cuda())
    close(out.detach(), ref(c.double(), s1.double(), s2.double()).float(), rtol=1e-4)
    for got, want in ((cg.grad, cr.grad), (ag.grad, ar.grad), (bg.grad, br.grad)):
        close(got, want.float(), rtol=2e-4)
    # only the styles require grad / only the content requires grad
    a2, b2 = (t.cuda().requires_grad_(True) for t in (s1, s2))
    layer(c.cuda(), [a2, b2], alpha=alpha, style_weights=w).backward(go.cuda())
    close(a2.grad, ar.grad.float(), rtol=2e-4); close(b2.grad, br.grad.float(), rtol=2e-4)
    c2 = c.cuda().requires_grad_(True)
    layer(c2, [s1.cuda(), s2.cuda()], alpha=alpha, style_weights=w).backward(go.cuda())
    close(c2.grad, cr.grad.float(), rtol=2e-4)


def test_errors():
    from arbitrarystyletransfer_b200 import _lib as L, functional as Fn
    c = torch.zeros(1, 4, 8, 8, device="cuda")
    with pytest.raises(L.AstError):
        Fn.adain_forward(c, [c] * 9)                      # more than AST_MAX_STYLES
    with pytest.raises(L.AstError):
        Fn.adain_forward(c, [torch.zeros(1, 5, 8, 8, device="cuda")])   # channel mismatch
    with pytest.raises(L.AstError):
        Fn.adain_forward(c.cpu(), [c.cpu()])              # no CPU fallback
