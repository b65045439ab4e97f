"""K3 parity: Huber / Gram / style-loss kernels and their backward passes vs the golden vectors of
the genuine reference (losses.py:105-139) and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import restate as R
from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(params=["fp32", "tf32"])
def gram_mode(request):
    """fp32 = CUDA-core Gram (1e-5 bars); tf32 = tensor-core Gram forward (inputs rounded to 10
    mantissa bits: Gram entries agree to ~1e-3, so loss / gradient bars are 2e-3 / 1e-2)."""
    from arbitrarystyletransfer_b200 import functional as Fn
    old = Fn.GRAM_PRECISION
    Fn.GRAM_PRECISION = request.param
    yield request.param
    Fn.GRAM_PRECISION = old


def test_golden_losses(golden_losses, gram_mode):
    tol = 1.0 if gram_mode == "fp32" else 200.0
    from arbitrarystyletransfer_b200 import losses as Ls
    g = golden_losses
    a, b = T(g["loss_a"]).cuda().requires_grad_(True), T(g["loss_b"]).cuda()
    l = Ls.compute_content_loss(a, b)
    assert l.dim() == 0
    assert l.item() == pytest.approx(float(g["content_loss"]), rel=1e-5)
    l.backward()
    torch.testing.assert_close(a.grad.cpu(), T(g["content_loss_ga"]), rtol=1e-5, atol=1e-9)
    a.grad = None
    gm = Ls.gram_matrix(a)
    torch.testing.assert_close(gm.detach().cpu(), T(g["gram_a"]), rtol=1e-5 * tol, atol=1e-6 * tol)
    (gm * T(g["gram_gg"]).cuda()).sum().backward()
    torch.testing.assert_close(a.grad.cpu(), T(g["gram_ga"]), rtol=1e-4, atol=1e-6)
    a.grad = None
    l = Ls.compute_style_loss(a, b)
    assert l.item() == pytest.approx(float(g["style_loss"]), rel=1e-5 * tol)
    l.backward()
    torch.testing.assert_close(a.grad.cpu(), T(g["style_loss_ga"]), rtol=1e-4 * tol, atol=1e-7 * tol)


@pytest.mark.parametrize("shape", [(2, 64, 32, 32), (1, 128, 17, 19), (1, 512, 16, 16), (3, 3, 40, 40)])
def test_style_loss_vs_oracle(shape, gram_mode):
    tol = 1.0 if gram_mode == "fp32" else 100.0
    from arbitrarystyletransfer_b200 import losses as Ls
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.randn(*shape, generator=g) * 1.5
    b = torch.randn(*shape, generator=g) * 1.2 + 0.2
    ar = a.clone().requires_grad_(True)
    lr = R.compute_style_loss(ar, b)
    lr.backward()
    ag = a.cuda().requires_grad_(True)
    lg = Ls.compute_style_loss(ag, b.cuda())
    lg.backward()
    assert lg.item() == pytest.approx(lr.item(), rel=2e-5 * tol)
    torch.testing.assert_close(ag.grad.cpu(), ar.grad, rtol=1e-3 * min(tol, 10), atol=(1e-7 + 1e-4 * ar.grad.abs().max().item()) * min(tol, 10))


def test_huber_large_and_both_branches():
    from arbitrarystyletransfer_b200 import functional as Fn
    g = torch.Generator().manual_seed(1)
    a = torch.randn(3, 7, 129, 65, generator=g) * 2      # |d| on both sides of delta = 1
    b = torch.randn(3, 7, 129, 65, generator=g)
    ref = R.huber_np(a.numpy(), b.numpy())
    got = Fn.huber_loss(a.cuda(), b.cuda()).item()
    assert got == pytest.approx(ref, rel=1e-5)
    assert Fn.huber_loss(a.cuda(), b.cuda(), 10.0).item() == pytest.approx(10 * ref, rel=1e-5)


def test_tv_loss_golden_and_shapes(golden_losses):
    """tv_loss (losses.py:90-103): the reference's golden value / gradient, then ragged shapes vs the oracle
    (1-pixel-wide and 1-pixel-high images exercise the missing-neighbour cases)."""
    from arbitrarystyletransfer_b200 import losses as Ls
    g = golden_losses
    a = T(g["loss_a"]).cuda().requires_grad_(True)
    l = Ls.tv_loss(a)
    assert l.dim() == 0 and l.item() == pytest.approx(float(g["tv_loss"]), rel=1e-5)
    (l * 0.37).backward()
    torch.testing.assert_close(a.grad.cpu(), T(g["tv_loss_ga"]), rtol=1e-5, atol=1e-6)
    for shape in [(1, 3, 64, 48), (2, 3, 1, 17), (2, 5, 9, 1), (1, 1, 1, 1), (4, 3, 256, 256)]:
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(sum(shape)))
        xr = x.clone().requires_grad_(True)
        lr = R.tv_loss(xr)
        lr.backward()
        xg = x.cuda().requires_grad_(True)
        lg = Ls.tv_loss(xg)
        lg.backward()
        assert lg.item() == pytest.approx(lr.item(), rel=2e-5, abs=1e-6), shape
        torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape", [(2, 64, 32, 32), (1, 128, 16, 24), (3, 256, 16, 16), (1, 512, 8, 32), (2, 64, 64, 64)])
def test_gram_backward_tensor_core(shape):
    """Gram backward as a bf16 MN-major tcgen05 GEMM vs fp64: (gg + gg^T) X / (C*HW)."""
    from arbitrarystyletransfer_b200 import functional as Fn
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, C, H, W, generator=g)
    gg = torch.randn(B, C, C, generator=g)
    X = x.double().view(B, C, H * W)
    want = torch.bmm(gg.double() + gg.double().transpose(1, 2), X).view(B, C, H, W) / (C * H * W)
    old = Fn.GRAM_PRECISION
    try:
        Fn.GRAM_PRECISION = "tf32"
        xg = x.cuda().requires_grad_(True)
        (Fn.gram_matrix(xg) * gg.cuda()).sum().backward()
        got = xg.grad.cpu().double()
    finally:
        Fn.GRAM_PRECISION = old
    err = ((got - want).norm() / want.norm()).item()
    assert err < 6e-3, err      # bf16 operands: 2^-9 relative rounding of X and S, fp32 accumulation


def test_hist_loss_golden_oracle_and_determinism():
    """compute_hist_loss (losses.py:8-87): the reference's golden values / gradients (out-of-range values included),
    then train.py's size (8 x 3 x 256 x 256, where the reference would materialise 2 x 1.6 GB) against the chunked
    fp64 oracle; integer accumulation makes the result bit-deterministic."""
    from arbitrarystyletransfer_b200 import losses as Ls
    from tests.conftest import load_golden
    g = load_golden("hist")
    for tag in "ab":
        x = T(g[f"hist_{tag}_x"]).cuda().requires_grad_(True)
        y = T(g[f"hist_{tag}_y"]).cuda().requires_grad_(True)
        l = Ls.compute_hist_loss(x, y)
        assert l.dim() == 0 and l.item() == pytest.approx(float(g[f"hist_{tag}_loss"]), rel=2e-5)
        (l * 0.5).backward()
        for got, ref in ((x.grad, T(g[f"hist_{tag}_gx"])), (y.grad, T(g[f"hist_{tag}_gy"]))):
            assert rel_err(got.cpu(), ref) < 1e-4
            torch.testing.assert_close(got.cpu(), ref, rtol=1e-3, atol=1e-4 * float(ref.abs().max()))
    gen = torch.Generator().manual_seed(77)
    x = torch.rand(8, 3, 256, 256, generator=gen) * 1.1 - 0.05
    y = torch.rand(8, 3, 256, 256, generator=gen) ** 1.5
    xr = x[:2].double().requires_grad_(True)
    lr = R.compute_hist_loss(xr, y[:2].double())
    lr.backward()
    xg = x[:2].cuda().requires_grad_(True)
    lg = Ls.compute_hist_loss(xg, y[:2].cuda())
    lg.backward()
    assert lg.item() == pytest.approx(lr.item(), rel=2e-5)
    assert rel_err(xg.grad.cpu(), xr.grad.float()) < 1e-4
    a = Ls.compute_hist_loss(x.cuda(), y.cuda())
    b = Ls.compute_hist_loss(x.cuda(), y.cuda())
    assert torch.equal(a, b) and torch.isfinite(a)
    assert Ls.compute_hist_loss(y.cuda(), y.cuda()).item() == 0.0
    bad = x.clone()
    bad[3, 1, 5, 5] = float("nan")
    assert torch.isnan(Ls.compute_hist_loss(bad.cuda(), y.cuda()))
