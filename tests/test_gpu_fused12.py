"""conv1_1 + conv1_2 (+ pool) fused kernel (ast_conv12_fused, csrc/conv12_fused.cuh) against the two separate launches
and against torch fp32 on the same bf16-rounded operands: ragged sizes (partial tiles, odd tile counts -> phantom tile of
the last CTA pair), borders, batch > 1."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _weights(seed):
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(64, 3, 3, 3, generator=g) * 0.4
    b1 = torch.randn(64, generator=g) * 0.2
    w2 = torch.randn(64, 64, 3, 3, generator=g) * (2.0 / (9 * 64)) ** 0.5
    b2 = torch.randn(64, generator=g) * 0.1
    return w1, b1, w2, b2


@pytest.mark.parametrize("N,H,W", [(1, 16, 8), (1, 32, 16), (2, 64, 64), (1, 18, 20), (3, 34, 44), (1, 256, 256),
                                   (1, 2, 4), (5, 16, 8)])
def test_fused12_matches_separate_launches_and_torch(N, H, W):
    from arbitrarystyletransfer_b200 import engine as E, _lib as L
    dev = torch.device("cuda")
    w1, b1, w2, b2 = (t.to(dev) for t in _weights(7))
    img = torch.rand(N, 3, H, W, generator=torch.Generator().manual_seed(H * 131 + W)).to(dev)
    wpk2 = E.pack_conv_weight(w2)
    fused = E.native_empty(N, H // 2, W // 2, 64, dev, True)
    E.conv12_fused(img, w1, b1, wpk2, b2, fused)
    # the two separate launches
    x = E.native_empty(N, H, W, 64, dev, True)
    E.conv3x3_first(img, w1, b1, x)
    sep = E.native_empty(N, H // 2, W // 2, 64, dev, True)
    E.conv3x3(x, wpk2, b2, sep, N=N, H=H, W=W, cin=64, cout=64, relu=True, epilogue=L.EPI_POOL2, halo=L.HALO_KEEP)
    torch.cuda.synchronize()
    f, s = fused.float(), sep.float()
    assert torch.isfinite(f).all()
    assert float(f[:, 0].abs().max()) == 0.0 and float(f[:, :, 0].abs().max()) == 0.0      # halo untouched (zeros)
    # same arithmetic (bf16 operands, fp32 accumulation; only the accumulation order inside conv1_1 can differ): 1 bf16 ulp
    torch.testing.assert_close(f, s, rtol=2 ** -7, atol=2e-2)
    assert float((f - s).abs().mean()) < 2e-3
    # torch fp32 on bf16-rounded operands
    mean = torch.tensor(E.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(E.IMAGENET_STD, device=dev).view(1, 3, 1, 1)
    xn = ((img - mean) / std).bfloat16().float()
    a1 = torch.relu(torch.nn.functional.conv2d(xn, w1.bfloat16().float(), b1, padding=1)).bfloat16().float()
    a2 = torch.relu(torch.nn.functional.conv2d(a1, w2.bfloat16().float(), b2, padding=1))
    ref = torch.nn.functional.max_pool2d(a2, 2, 2).permute(0, 2, 3, 1)
    got = f[:, 1:-1, 1:-1, :]
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 1e-2, rel


def test_engine_uses_the_fused_kernel_and_matches_the_unfused_path():
    from arbitrarystyletransfer_b200.engine import StyleTransferEngine
    from oracle import restate as R
    vw, _ = R.make_vgg_weights(0)
    vb = R.calibrate_vgg_bias(vw)
    dw, db = R.make_decoder_weights(1)
    eng = StyleTransferEngine(vw[:9], vb[:9], dw, db, device="cuda:0")
    c, s = R.rand_image(2, 96, 11).cuda(), R.rand_image(2, 96, 12).cuda()
    assert eng.fused12_ok(96, 96) and eng.launches_per_stylize(1, 96, 96) == 28
    a = eng.stylize(c, s).clone()
    eng.fuse12 = False
    assert eng.launches_per_stylize(1, 96, 96) == 30
    b = eng.stylize(c, s).clone()
    assert R.psnr(a.cpu(), b.cpu()) > 45.0
