"""AdaAttN (models.py:70-115) and AST (models.py:393-575) -- SURVEY.md section 8 row f1 -- on the GPU, through the
C ABI, against the CPU oracle (oracle/restate_attn.py) and the golden vectors made by the genuine reference classes
(tests/golden/adaattn.npz).

Tolerances.  The logits Q K^T (and W_q, W_k in front of them) run as two-term bf16 splits (~2^-17); the attention
weights, v and the gradients are bf16 (the second moment is exact: v^2 = hi + lo), accumulation and softmax are fp32,
the layer's output is stored in bf16: a single layer is held to
relative L2 <= 1e-2 on its output and <= 3e-2 / cosine >= 0.999 on its gradients, the AST network (15 encoder
blocks, two attention layers, 18 decoder blocks, train-mode BatchNorm over 2 samples) to the bar the autoencoder
tests use: image PSNR >= 40 dB, gradient cosine >= 0.97, norm ratio within 15 %."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate as R
from oracle import restate_ae as A
from oracle import restate_attn as T
from tests.conftest import load_golden
from tests.gpu_util import bf16r

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()


def cos(a, b):
    return F.cosine_similarity(a.double().cpu().flatten(), b.double().cpu().flatten(), dim=0).item()


def t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def g():
    return load_golden("adaattn")


# ------------------------------------------------------------------------------------------------
# the batched tcgen05 GEMM, all four operand layouts, ragged sizes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("mnk", [(96, 60, 32), (128, 256, 64), (300, 520, 200), (1024, 384, 1024), (60, 64, 96),
                                 (17, 24, 8)])
def test_bgemm_layouts(a_mn, b_mn, mnk):
    from arbitrarystyletransfer_b200 import attention as AT
    M, N, K = mnk
    B = 3
    gen = torch.Generator().manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    a = bf16r(torch.randn(B, M, K, generator=gen))
    b = bf16r(torch.randn(B, N, K, generator=gen))
    ref = torch.einsum("bik,bjk->bij", a.double(), b.double())

    def dev(x, mn):
        # stored [rows][inner] with the inner extent padded to a multiple of 8 (the ABI's alignment rule)
        x = x.transpose(1, 2).contiguous() if mn else x.contiguous()
        inner = x.shape[2]
        buf = torch.zeros(B, x.shape[1], (inner + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")
        buf[:, :, :inner] = x.to(torch.bfloat16).cuda()
        return buf

    da, db = dev(a, a_mn), dev(b, b_mn)
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 4e-3)):
        d = AT.bgemm(da, a_mn, db, b_mn, M, N, K, out_dtype=dt)
        assert d.shape[1] == M
        assert rel(d[:, :, :N].float(), ref) < tol, (a_mn, b_mn, mnk, dt)


# ------------------------------------------------------------------------------------------------
# streaming passes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(64, 60), (200, 1024), (7, 4100)])
def test_softmax_rows_and_backward(rows, cols):
    from arbitrarystyletransfer_b200 import _lib as L
    lib = L.load()
    gen = torch.Generator().manual_seed(rows + cols)
    s = (torch.randn(rows, cols, generator=gen) * 4).cuda()
    ld = (cols + 7) // 8 * 8
    p = torch.zeros(rows, ld, dtype=torch.bfloat16, device="cuda")
    lsum = torch.empty(rows, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.ast_attn_softmax(s.data_ptr(), cols, p.data_ptr(), ld, lsum.data_ptr(), rows, cols, st))
    ref = torch.softmax(s.double(), -1)
    assert rel(p[:, :cols].float(), ref) < 4e-3
    torch.testing.assert_close(lsum.double(), p[:, :cols].double().sum(-1), rtol=1e-5, atol=1e-6)
    da = torch.randn(rows, cols, generator=gen).cuda()
    ds = torch.zeros(rows, ld, dtype=torch.bfloat16, device="cuda")
    L.check(lib.ast_attn_softmax_bwd(p.data_ptr(), ld, lsum.data_ptr(), da.data_ptr(), cols, ds.data_ptr(), ld, rows,
                                     cols, st))
    pr = p[:, :cols].double()
    ref_ds = pr * (da.double() - (da.double() * pr).sum(-1, keepdim=True) / pr.sum(-1, keepdim=True))
    # a per-row constant added to dA must not change dS (softmax is shift-invariant): this is what the /lsum buys
    da2 = da + 1000.0
    ds2 = torch.zeros_like(ds)
    L.check(lib.ast_attn_softmax_bwd(p.data_ptr(), ld, lsum.data_ptr(), da2.data_ptr(), cols, ds2.data_ptr(), ld, rows,
                                     cols, st))
    assert rel(ds2[:, :cols].float(), ref_ds) < 2e-2
    assert rel(ds[:, :cols].float(), ref_ds) < 5e-3


def test_second_moment_is_exact_under_peaked_attention():
    """One key carries all the weight: the attention-weighted std must vanish (models.py:103), which it only does
    if sum p v^2 and (sum p v)^2 are computed from the same v without rounding v^2."""
    from arbitrarystyletransfer_b200 import attention as AT
    C, HW = 32, 64
    gen = torch.Generator().manual_seed(3)
    key = F.normalize(torch.randn(1, HW, C, generator=gen), dim=-1)
    nchw = lambda x: x.transpose(1, 2).reshape(1, C, 8, 8).contiguous().cuda()
    v = bf16r(torch.randn(1, HW, C, generator=gen) * 3 + 5)
    eye = torch.eye(C).view(C, C, 1, 1).cuda()
    # q = k = 12 * key: each query matches exactly one key (logit 144 vs <~ 90 for the others); v = style itself
    out, _ = AT._layer_forward(nchw(key), nchw(key), nchw(v), eye * 12, eye * 12, eye)
    assert rel(out.float().view(1, HW, C), v) < 5e-3              # std = 0, mean = the matched key's v


def test_split_logits_are_fp32_accurate():
    """Q K^T and the 1x1 convolutions in front of it run as two-term bf16 splits: logits of O(50) must come out with
    an absolute error << 1e-2 (plain bf16 operands: ~1e-1, i.e. 10 % on the attention weights)."""
    from arbitrarystyletransfer_b200 import attention as AT
    N, C, HW = 2, 64, 96
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(N * HW, C, generator=gen)
    w = torch.randn(C, C, generator=gen) * 0.9
    y = torch.randn(N * HW, C, generator=gen)
    x3 = AT._split3_rows(x.cuda(), C, 0)
    w3 = AT._split3_rows(w.cuda(), C, 1)
    q = AT.bgemm(x3.view(1, N * HW, 3 * C), 0, w3.view(1, C, 3 * C), 0, N * HW, C, 3 * C)[0]
    ref_q = x.double() @ w.double().t()
    assert (q.double().cpu() - ref_q).abs().max() < 2e-4 * ref_q.abs().max()
    q3 = AT._split3_rows(q, C, 0).view(N, HW, 3 * C)
    k3 = AT._split3_rows(y.cuda(), C, 1).view(N, HW, 3 * C)
    S = AT.bgemm(q3, 0, k3, 0, HW, HW, 3 * C)
    ref = torch.einsum("nic,njc->nij", ref_q.view(N, HW, C), y.double().view(N, HW, C))
    assert ref.abs().max() > 100
    assert (S[:, :, :HW].double().cpu() - ref).abs().max() < 5e-3


# ------------------------------------------------------------------------------------------------
# the layer against the genuine reference (golden) and the oracle
# ------------------------------------------------------------------------------------------------
def _layer(g, tag, C):
    from arbitrarystyletransfer_b200 import attention as AT
    layer = AT.AdaAttN(C).cuda()
    with torch.no_grad():
        for n in ("W_q", "W_k", "W_v"):
            getattr(layer, n).weight.copy_(t(g[f"{tag}_{n}"]))
    return layer


@pytest.mark.parametrize("tag", ["flat", "sharp"])
def test_adaattn_forward_backward_vs_reference_golden(g, tag):
    layer = _layer(g, tag, 32)
    c = t(g[f"{tag}_content"]).cuda().requires_grad_(True)
    s = t(g[f"{tag}_style"]).cuda().requires_grad_(True)
    y = layer(c, s)
    assert y.shape == tuple(g[f"{tag}_out"].shape) and y.dtype == torch.float32
    assert rel(y, t(g[f"{tag}_out"])) < 1e-2
    y.backward(t(g[f"{tag}_gy"]).cuda())
    for n in ("W_q", "W_k", "W_v"):
        got, ref = getattr(layer, n).weight.grad, t(g[f"{tag}_g{n}"])
        assert cos(got, ref) > 0.999 and rel(got, ref) < 3e-2, (n, cos(got, ref), rel(got, ref))
    for got, ref, name in ((c.grad, t(g[f"{tag}_gcontent"]), "content"), (s.grad, t(g[f"{tag}_gstyle"]), "style")):
        assert cos(got, ref) > 0.999 and rel(got, ref) < 3e-2, (name, cos(got, ref), rel(got, ref))


def test_adaattn_no_grad_matches_grad_path_and_is_deterministic(g):
    layer = _layer(g, "sharp", 32)
    c, s = t(g["sharp_content"]).cuda(), t(g["sharp_style"]).cuda()
    with torch.no_grad():
        y0, y1 = layer(c, s), layer(c, s)
    y2 = layer(c.clone().requires_grad_(True), s)
    assert torch.equal(y0, y1) and torch.equal(y0, y2.detach())


@pytest.mark.parametrize("shape", [(2, 128, 32, 32, 32, 32), (1, 128, 20, 12, 16, 24), (3, 64, 12, 12, 40, 40)])
def test_adaattn_vs_oracle_at_network_sizes(shape):
    """C = 128, 32 x 32 = 1024 positions is the layer inside AST at 256^2 input."""
    from arbitrarystyletransfer_b200 import attention as AT
    N, C, h, w, hs, ws = shape
    torch.manual_seed(sum(shape))
    layer = AT.AdaAttN(C)
    with torch.no_grad():
        layer.W_q.weight.mul_(3.0)
        layer.W_k.weight.mul_(3.0)
    P = {"a." + n + ".weight": getattr(layer, n).weight.detach().clone().requires_grad_(True) for n in ("W_q", "W_k", "W_v")}
    layer = layer.cuda()
    gen = torch.Generator().manual_seed(9)
    c = (torch.randn(N, C, h, w, generator=gen) * 1.5 + 0.5)
    s = (torch.randn(N, C, hs, ws, generator=gen) * 2 + 1)
    gy = torch.randn(N, C, h, w, generator=gen)
    cr, sr = c.clone().requires_grad_(True), s.clone().requires_grad_(True)
    ref = T.adaattn(P, "a", cr, sr)
    ref.backward(gy)
    cd, sd = c.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    y = layer(cd, sd)
    y.backward(gy.cuda())
    assert rel(y, ref.detach()) < 1e-2
    for n in ("W_q", "W_k", "W_v"):
        got, want = getattr(layer, n).weight.grad, P[f"a.{n}.weight"].grad
        assert cos(got, want) > 0.999 and rel(got, want) < 3e-2, (n, cos(got, want), rel(got, want))
    assert cos(cd.grad, cr.grad) > 0.999 and rel(cd.grad, cr.grad) < 3e-2
    assert cos(sd.grad, sr.grad) > 0.999 and rel(sd.grad, sr.grad) < 3e-2


def test_adaattn_argument_errors():
    from arbitrarystyletransfer_b200 import _lib as L
    from arbitrarystyletransfer_b200 import attention as AT
    layer = AT.AdaAttN(16).cuda()
    with pytest.raises(L.AstError):
        layer(torch.zeros(1, 16, 4, 4), torch.zeros(1, 16, 4, 4))                     # CPU tensors: no fallback
    with pytest.raises(L.AstError):
        layer(torch.zeros(1, 8, 4, 4).cuda(), torch.zeros(1, 16, 4, 4).cuda())        # channel mismatch
    with pytest.raises(L.AstError):
        layer(torch.zeros(2, 16, 4, 4).cuda(), torch.zeros(1, 16, 4, 4).cuda())       # batch mismatch


# ------------------------------------------------------------------------------------------------
# the AST network against the genuine reference (golden)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ast_state(g):
    """activate_gates(seeded state) with the encoder's running statistics calibrated on the fixture's two images
    (oracle/restate_attn.py::calibrate_encoder says why)."""
    P = A.clone_state(A.activate_gates(T.make_ast_state(3)))
    x = torch.cat((t(g["ast_content"]), t(g["ast_style"])))
    return T.calibrate_encoder(P, x)


def test_ast_forward_and_gradients_vs_reference_golden(g, ast_state):
    from arbitrarystyletransfer_b200 import attention as AT
    net = AT.AST().cuda()
    assert sorted(net.state_dict().keys()) == list(g["ast_state_keys"])
    net.load_state_dict(ast_state, strict=True)
    net.train()
    c, s = t(g["ast_content"]).cuda(), t(g["ast_style"]).cuda()
    t_cs, t_ret, org = net(c, s, alpha=0.75)
    for got, key in ((t_cs, "ast_t_cs"), (org, "ast_org_out")):
        ref = t(g[key])
        assert got.shape == ref.shape and got.dtype == torch.float32
        assert R.psnr(got.detach().cpu(), ref) >= 40.0, (key, R.psnr(got.detach().cpu(), ref))
    # t_return = ada_att_1 of the eval-mode encoder taps: the taps themselves carry the encoder's bf16-storage drift
    # (8-10 % relative L2 at the deepest taps on this state, profiles/r1_ae_error_growth.txt); the layer alone is
    # isolated in test_ast_encode_public_surface
    assert rel(t_ret, t(g["ast_t_return"])) < 0.2 and cos(t_ret, t(g["ast_t_return"])) > 0.98
    assert net._enc.training                                     # models.py:547 leaves the encoder in train mode
    loss = F.huber_loss(t_cs, s) + 0.5 * F.huber_loss(org, c) + 0.1 * t_ret.mean()
    np.testing.assert_allclose(loss.item(), float(g["ast_loss"]), rtol=2e-2)
    loss.backward()
    named = dict(net.named_parameters())
    gk = sorted(k for k, p in named.items() if p.grad is not None)
    assert gk == list(g["ast_grad_keys"])
    ratios = np.array([named[k].grad.double().norm().item() for k in gk]) / np.maximum(g["ast_grad_norm"], 1e-30)
    big = g["ast_grad_norm"] > 1e-3 * g["ast_grad_norm"].max()
    assert np.all(np.abs(ratios[big] - 1) < 0.15), (ratios[big].min(), ratios[big].max())
    for k in T.GOLDEN_GRAD_KEYS:
        ref = t(g["ast_grad::" + k])
        # the attention layers read the deepest encoder taps (8-10 % bf16-storage drift on this state): measured
        # cosines 0.969-0.999 there, >= 0.97 elsewhere (the autoencoder tests' bar)
        bar = 0.95 if k.startswith("ada_att_") else 0.97
        assert cos(named[k].grad, ref) > bar, (k, cos(named[k].grad, ref))
    rm = net.state_dict()["_enc.mob_net.1._layers.1.running_mean"]
    assert rel(rm, t(g["ast_buf::_enc.mob_net.1._layers.1.running_mean"])) < 2e-2


def test_ast_exporting_forward_vs_reference_golden(g, ast_state):
    from arbitrarystyletransfer_b200 import attention as AT
    net = AT.AST(exporting=True).cuda()
    net.load_state_dict(ast_state, strict=True)
    net.eval()
    c, s = t(g["ast_content"]).cuda(), t(g["ast_style"]).cuda()
    with torch.no_grad():
        y = net(c, s)
    ref = t(g["ast_export_t_cs"])
    assert y.shape == ref.shape and float(y.min()) >= 0.0 and float(y.max()) <= 1.0
    assert R.psnr(y.cpu(), ref) >= 40.0, R.psnr(y.cpu(), ref)


def test_ast_encode_public_surface(g, ast_state):
    from arbitrarystyletransfer_b200 import attention as AT
    net = AT.AST().cuda()
    net.load_state_dict(ast_state, strict=True)
    c, s = t(g["ast_content"]).cuda(), t(g["ast_style"]).cuda()
    s1, s2, z = net.encode(c, s, detach=True, return_maps=True)
    assert s1.shape == (2, 128, 8, 8) and s2.shape == (2, 128, 8, 8) and z.shape == (2, 128, 8, 8)
    assert z.dtype == torch.float32 and z.requires_grad           # W_q/k/v and ada_out are trainable through it
    assert net._enc.training                                      # models.py:547
    # the layers and ada_out in isolation: the oracle fed the CUDA encoder's own taps
    net._enc.eval()
    with torch.no_grad():
        cm = [x.cpu() for x in net._enc(c, out_layers=AT.enc_out_layers)]
        sm = [x.cpu() for x in net._enc(s, out_layers=AT.enc_out_layers)]
    P = A.clone_state(ast_state)
    with torch.no_grad():
        r1 = T.adaattn(P, "ada_att_1", cm[0], sm[0])
        r2 = T.adaattn(P, "ada_att_2", cm[1], sm[1])
        rz = T.ada_out(P, torch.cat((s1.detach().cpu(), s2.detach().cpu()), dim=1))
    assert rel(s1, r1) < 1e-2 and rel(s2, r2) < 1e-2, (rel(s1, r1), rel(s2, r2))
    assert rel(z, rz) < 3e-2, rel(z, rz)


def test_ast_alpha_zero_is_independent_of_style(g, ast_state):
    """alpha = 0 (train.py:380 preview): t = content_map, so t_cs == org_out whatever the style image is."""
    from arbitrarystyletransfer_b200 import attention as AT
    net = AT.AST().cuda()
    net.load_state_dict(ast_state, strict=True)
    c, s = t(g["ast_content"]).cuda(), t(g["ast_style"]).cuda()
    with torch.no_grad():
        a, _, org_a = net(c, s, alpha=0.0)
    assert rel(a, org_a) < 1e-3


@pytest.mark.parametrize("gain", [1.0, 6.0])
def test_adaattn_forward_vs_its_precision_contract(gain):
    """The CUDA forward path against the CPU restatement of its own storage / precision contract
    (oracle/restate_attn.py::adaattn_contract): what remains is accumulation order, the fast exponential and bf16
    rounding flips of the output -- several times smaller than the contract's own distance to fp32 (3e-3)."""
    from arbitrarystyletransfer_b200 import attention as AT
    torch.manual_seed(0)
    C = 64
    layer = AT.AdaAttN(C)
    with torch.no_grad():
        layer.W_q.weight.mul_(gain)
        layer.W_k.weight.mul_(gain)
    P = {f"a.{n}.weight": getattr(layer, n).weight.detach().clone() for n in ("W_q", "W_k", "W_v")}
    layer = layer.cuda()
    c, s = torch.randn(2, C, 12, 12) * 1.5 + 0.5, torch.randn(2, C, 10, 14) * 2 + 1
    with torch.no_grad():
        y = layer(c.cuda(), s.cuda())
        con = T.adaattn_contract(P, "a", c, s)
    assert rel(y, con) < 2e-3, rel(y, con)
