"""The AdaAttN / AST oracle (oracle/restate_attn.py) against golden vectors produced by the GENUINE reference
``AdaAttN`` (models.py:70-115) and ``AST`` (models.py:393-575, with the two commented-out attributes restored and the
:459 token fix; oracle/make_golden.py::golden_adaattn, executed in the build container).  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate_ae as A
from oracle import restate_attn as T
from tests.conftest import load_golden


@pytest.fixture(scope="module")
def g():
    return load_golden("adaattn")


def t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag", ["flat", "sharp"])
def test_adaattn_layer_forward_and_gradients(g, tag):
    P = {f"att.{n}.weight": t(g[f"{tag}_{n}"]).clone().requires_grad_(True) for n in ("W_q", "W_k", "W_v")}
    c = t(g[f"{tag}_content"]).clone().requires_grad_(True)
    s = t(g[f"{tag}_style"]).clone().requires_grad_(True)
    y = T.adaattn(P, "att", c, s)
    torch.testing.assert_close(y.detach(), t(g[f"{tag}_out"]), rtol=1e-5, atol=1e-5)
    y.backward(t(g[f"{tag}_gy"]))
    for n in ("W_q", "W_k", "W_v"):
        torch.testing.assert_close(P[f"att.{n}.weight"].grad, t(g[f"{tag}_g{n}"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(c.grad, t(g[f"{tag}_gcontent"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(s.grad, t(g[f"{tag}_gstyle"]), rtol=1e-4, atol=1e-5)


def test_sharp_fixture_really_is_peaked(g):
    """The 'sharp' fixture must exercise softmax: its largest attention weight per query is far from uniform."""
    c, s = t(g["sharp_content"]), t(g["sharp_style"])
    q = F.conv2d(T.instance_norm(c), t(g["sharp_W_q"])).flatten(2).transpose(1, 2)
    k = F.conv2d(T.instance_norm(s), t(g["sharp_W_k"])).flatten(2)
    att = torch.softmax(q @ k, -1)
    assert att.max(-1).values.median() > 5.0 / att.shape[-1]


@pytest.fixture(scope="module")
def ast_state(g):
    """activate_gates(seeded state) with the encoder's running statistics calibrated on the fixture's two images
    (oracle/restate_attn.py::calibrate_encoder says why)."""
    P = A.clone_state(A.activate_gates(T.make_ast_state(3)))
    x = torch.cat((t(g["ast_content"]), t(g["ast_style"])))
    return T.calibrate_encoder(P, x)


def test_ast_state_keys_match_reference(g, ast_state):
    assert sorted(ast_state.keys()) == list(g["ast_state_keys"])
    k = "_enc.mob_net.14._layers.8.running_var"
    torch.testing.assert_close(ast_state[k], t(g["ast_buf_cal::" + k]), rtol=1e-5, atol=1e-7)


def test_ast_fixture_attention_is_peaked(g, ast_state):
    """The network fixture must exercise softmax: median of the largest weight per query >> 1/64."""
    P = A.clone_state(ast_state)
    with torch.no_grad():
        cm = A.encoder_forward(P, t(g["ast_content"]), A.ENC_OUT_LAYERS, prefix="_enc")
        sm = A.encoder_forward(P, t(g["ast_style"]), A.ENC_OUT_LAYERS, prefix="_enc")
        q = F.conv2d(T.instance_norm(cm[0]), P["ada_att_1.W_q.weight"]).flatten(2).transpose(1, 2)
        k = F.conv2d(T.instance_norm(sm[0]), P["ada_att_1.W_k.weight"]).flatten(2)
        att = torch.softmax(q @ k, -1)
    assert att.max(-1).values.median() > 0.2


def test_ast_forward_losses_and_gradients(g, ast_state):
    P = A.clone_state(ast_state, requires_grad=True)
    c, s = t(g["ast_content"]), t(g["ast_style"])
    t_cs, t_ret, org = T.ast_forward(P, c, s, alpha=0.75)
    torch.testing.assert_close(t_cs.detach(), t(g["ast_t_cs"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(t_ret.detach(), t(g["ast_t_return"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(org.detach(), t(g["ast_org_out"]), rtol=1e-4, atol=1e-5)
    loss = F.huber_loss(t_cs, s) + 0.5 * F.huber_loss(org, c) + 0.1 * t_ret.mean()
    np.testing.assert_allclose(loss.item(), float(g["ast_loss"]), rtol=1e-5)
    loss.backward()
    gk = sorted(k for k, v in P.items() if v.requires_grad and v.grad is not None)
    assert gk == list(g["ast_grad_keys"])
    norms = np.array([P[k].grad.double().norm().item() for k in gk])
    np.testing.assert_allclose(norms, g["ast_grad_norm"], rtol=2e-3, atol=1e-7)
    for k in T.GOLDEN_GRAD_KEYS:
        torch.testing.assert_close(P[k].grad, t(g["ast_grad::" + k]), rtol=2e-3, atol=1e-6)
    torch.testing.assert_close(P["_enc.mob_net.1._layers.1.running_mean"],
                               t(g["ast_buf::_enc.mob_net.1._layers.1.running_mean"]), rtol=1e-5, atol=1e-7)


def test_ast_exporting_forward(g, ast_state):
    P = A.clone_state(ast_state)
    with torch.no_grad():
        y = T.ast_forward(P, t(g["ast_content"]), t(g["ast_style"]), exporting=True, training=False)
    torch.testing.assert_close(y, t(g["ast_export_t_cs"]), rtol=1e-4, atol=1e-5)
    assert y.min() >= 0 and y.max() <= 1


@pytest.mark.parametrize("gain", [1.0, 3.0, 6.0])
def test_precision_contract_of_the_cuda_path_costs_under_half_a_percent(gain):
    """oracle/restate_attn.py::adaattn_contract (split-precision logits, bf16 attention weights normalised by their
    rounded sum, exact second moment, bf16 output) vs the fp32 restatement, flat to peaked attention: the error
    does not grow with the sharpness of the attention (which it does -- 2.5 % -- with plain bf16 logits)."""
    torch.manual_seed(0)
    C = 64
    P = {f"a.{n}.weight": torch.nn.Conv2d(C, C, 1, bias=False).weight.detach() * (gain if n != "W_v" else 1.0)
         for n in ("W_q", "W_k", "W_v")}
    c, s = torch.randn(2, C, 12, 12) * 1.5 + 0.5, torch.randn(2, C, 10, 14) * 2 + 1
    with torch.no_grad():
        ref, con = T.adaattn(P, "a", c, s), T.adaattn_contract(P, "a", c, s)
    assert ((con - ref).norm() / ref.norm()).item() < 5e-3
