"""The drop-in boundary: the reference's trainers bind to this package through the shim modules
under dropin/ (``models``, ``losses``, ``model_util``, ``mobilenetv2``, ``conf``).  CPU only.

* every name train.py / train_autoencoder.py take from ``from models import *`` / ``from conf import *`` /
  ``from losses import ...`` resolves in the shims (found by parsing the reference sources -- needs /root/reference);
* the GENUINE ``AutoencoderTrainer`` (train_autoencoder.py:17-72, exec'd from the reference tree with matplotlib and
  data_loader stubbed) constructs on the shim modules and round-trips ``ae.pth`` through its own save() / load();
* ``PretrainedEncoder(weights_path=...)`` loads a torchvision VGG-19 state dict offline (models.py:192)."""
import builtins
import importlib
import os
import symtable
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dropin")
REF = "/root/reference"
SHIMS = ("models", "losses", "model_util", "mobilenetv2", "conf")

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")),
                               reason="reference tree not present (build container only)")


@pytest.fixture()
def shim_path():
    """dropin/ in front of sys.path, shim modules freshly imported, everything restored afterwards."""
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k in SHIMS or k in ("data_loader",)}
    sys.path.insert(0, DROPIN)
    try:
        yield {name: importlib.import_module(name) for name in SHIMS}
    finally:
        for k in SHIMS + ("data_loader",):
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


def _free_globals(path):
    """Names a source file reads from its global scope without binding them itself (i.e. what its star imports must
    supply), and the names it binds at module level."""
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    bound = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    need = set()

    def walk(tab, is_top):
        for s in tab.get_symbols():
            if not s.is_referenced():
                continue
            if is_top:
                if not (s.is_assigned() or s.is_imported() or s.is_namespace()):
                    need.add(s.get_name())
            elif s.is_global():
                need.add(s.get_name())
        for ch in tab.get_children():
            walk(ch, False)
    walk(top, True)
    return {n for n in need if n not in bound and not hasattr(builtins, n)}, bound


def test_star_import_yields_the_reference_names(shim_path):
    m = shim_path["models"]
    for name in ("device", "enc_out_layers", "enc_out_channels", "compute_content_loss", "compute_style_loss",
                 "compute_hist_loss", "tv_loss", "gram_matrix", "channel_stats", "mean_variance_norm", "calc_mean_std",
                 "nn", "F", "torch", "AdaIN", "AdaAttN", "AST", "Encoder", "Decoder", "DecoderBlock", "AutoEncoder",
                 "PretrainedEncoder", "DepthWiseConv", "conv_3x3_bn", "EXPAND_RATIO", "img_sizes", "imsize",
                 "enc_conv_shapes", "decoder_conv_shapes", "discriminator_loss", "rgb2lab", "lab2rgb"):
        assert hasattr(m, name), name
    ns = {}
    exec("from models import *", ns)       # what train.py:15 executes
    assert ns["device"] in ("cuda", "cpu") and ns["enc_out_layers"] == [12, 14] and ns["enc_out_channels"] == 128
    assert ns["compute_content_loss"] is shim_path["losses"].compute_content_loss
    assert shim_path["conf"].img_sizes == [96, 128, 160]


@needs_ref
@pytest.mark.parametrize("trainer", ["train.py", "train_autoencoder.py"])
def test_every_name_the_reference_trainers_use_resolves(shim_path, trainer):
    need, _ = _free_globals(os.path.join(REF, trainer))
    # names the reference's own data_loader star import supplies (host-side data path, kept as is by a user)
    _, dl_bound = _free_globals(os.path.join(REF, "data_loader.py"))
    ns = {}
    exec("from models import *\nfrom conf import *\nfrom losses import compute_content_loss", ns)
    missing = sorted(n for n in need if n not in ns and n not in dl_bound)
    assert not missing, f"{trainer} reads names the drop-in modules do not supply: {missing}"


@needs_ref
def test_reference_autoencoder_trainer_constructs_and_round_trips_checkpoint(shim_path, tmp_path):
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    dl = types.ModuleType("data_loader")            # the loader is host-side code a user keeps; not needed here
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot", "data_loader")}
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "data_loader": dl})
    try:
        mod = types.ModuleType("ref_train_autoencoder")
        path = os.path.join(REF, "train_autoencoder.py")
        exec(compile(open(path).read(), path, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    from arbitrarystyletransfer_b200 import mobilenet as MB, models as M
    assert mod.AutoEncoder is MB.AutoEncoder and mod.PretrainedEncoder is M.PretrainedEncoder
    args = types.SimpleNamespace(lr=2e-4, save_dir=str(tmp_path), load=False, batch_size=2, train_iter=1,
                                 recon_lam=100.0, perp_lam=0.01)
    mod.device = "cpu"                               # no GPU in the build container: construction / checkpoint I/O only
    torch.manual_seed(2)
    tr = mod.AutoencoderTrainer(args, None, None)    # train_autoencoder.py:17-44
    assert sum(p.numel() for p in tr.model.parameters()) == 2925931
    g = tr.ae_optim.param_groups[0]
    assert tuple(g["betas"]) == (0.9, 0.99) and g["eps"] == 1e-7 and g["lr"] == 2e-4
    # attribute paths the reference's loop touches (train_autoencoder.py:145-146)
    assert tr.model.encoder.mob_net[0][0].weight.shape == (16, 3, 3, 3)
    assert tr.model.decoder._img_out.weight.shape == (3, 16, 3, 3)
    tr.train_dict["train_loss"].append(0.25)
    tr.save()                                        # :46-61 -> ae.pth {"AE", "optim"} + train_dict.json
    ck = torch.load(os.path.join(str(tmp_path), "ae.pth"))
    assert set(ck) == {"AE", "optim"} and len(ck["AE"]) == 434
    args2 = types.SimpleNamespace(**{**vars(args), "load": True})
    torch.manual_seed(99)
    tr2 = mod.AutoencoderTrainer(args2, None, None)  # -> load() :65-72
    for k, v in tr.model.state_dict().items():
        assert torch.equal(v, tr2.model.state_dict()[k]), k
    assert tr2.train_dict["train_loss"] == [0.25]
    # no CPU fallback: running the model on the CPU must fail loudly, not silently compute
    with pytest.raises(Exception):
        tr.model(torch.rand(1, 3, 32, 32))


def test_pretrained_encoder_loads_torchvision_vgg19_state_dict(tmp_path):
    tv = pytest.importorskip("torchvision")
    from arbitrarystyletransfer_b200 import models as M
    torch.manual_seed(5)
    vgg = tv.models.vgg19(weights=None)
    p = os.path.join(str(tmp_path), "vgg19.pth")
    torch.save(vgg.state_dict(), p)
    enc = M.PretrainedEncoder(['relu_9'], weights_path=p)        # offline stand-in for pretrained=True (models.py:192)
    convs = [m for m in vgg.features if isinstance(m, torch.nn.Conv2d)]
    assert len(convs) == 16
    for a, b in zip(enc._convs(), convs):
        assert torch.equal(a.weight, b.weight) and torch.equal(a.bias, b.bias)
    enc2 = M.PretrainedEncoder().load_vgg19_weights(vgg.features.state_dict())   # the features sub-module's own keys
    enc3 = M.PretrainedEncoder().load_vgg19_weights(enc.state_dict())            # the reference's keys
    for a, b, c in zip(enc._convs(), enc2._convs(), enc3._convs()):
        assert torch.equal(a.weight, b.weight) and torch.equal(a.weight, c.weight)
    bad = {k: v for k, v in vgg.state_dict().items() if not k.startswith("features.0.")}
    with pytest.raises(Exception):
        M.PretrainedEncoder().load_vgg19_weights(bad)


@needs_ref
def test_lab_colour_helpers_match_reference(shim_path):
    sys.path.insert(0, REF)
    try:
        sys.modules.pop("model_util", None)
        spec = importlib.util.spec_from_file_location("ref_model_util", os.path.join(REF, "model_util.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.path.remove(REF)
    mu = shim_path["model_util"]
    x = torch.rand(2, 3, 9, 11, generator=torch.Generator().manual_seed(4))
    x[0, :, 0, 0] = 0.0
    x[0, :, 0, 1] = 1.0
    lab = mu.rgb2lab(x)
    torch.testing.assert_close(lab, ref.rgb2lab(x), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mu.lab2rgb(lab), ref.lab2rgb(lab), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(mu.lab2rgb(lab), x, rtol=1e-3, atol=2e-3)
