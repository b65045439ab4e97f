"""Load the UNMODIFIED reference modules from /root/reference for oracle pinning.

TEST INFRASTRUCTURE ONLY.  Usable only inside the build container (the GPU box has no
/root/reference).  Nothing under ``arbitrarystyletransfer_b200/`` may import this file; it is
used by ``oracle/make_golden.py`` (fixture generator) and by ``tests/test_oracle_vs_reference.py``
(skipped when the reference tree is absent).

What it works around (all facts about the reference as shipped):
  * models.py:459 is a syntax error (``stylized_map_1 t = ...``); we patch that one token
    sequence in memory to what ``encode(..., return_maps=True)`` returns at models.py:568-569.
  * models.py:192 hard-codes ``models.vgg19(pretrained=True)`` which needs a network; we shim
    ``torchvision.models.vgg19`` to ``weights=None`` for the duration of the exec.
  * the classic mirrored decoder only exists as a commented ``nn.Sequential`` at
    models.py:598-628; we strip the leading ``# `` and exec that text.
No reference source is copied into this repo: it is read from /root/reference at run time.
"""
from __future__ import annotations

import os
import sys
import types

REF_DIR = os.environ.get("AST_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "models.py"))


_cache: dict = {}


def load_reference_models() -> types.ModuleType:
    """exec /root/reference/models.py with the line-459 fix and the vgg19 shim."""
    if "models" in _cache:
        return _cache["models"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_DIR}")
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torchvision.models as tvm

    orig_vgg19 = tvm.vgg19

    def vgg19_no_download(pretrained=False, **kw):  # models.py:192 call site
        return orig_vgg19(weights=None)

    path = os.path.join(REF_DIR, "models.py")
    with open(path, "r") as f:
        src = f.read()
    broken = "stylized_map_1 t = "
    assert src.count(broken) == 1, "reference models.py:459 no longer matches the survey"
    src = src.replace(broken, "stylized_map_1, stylized_map_2, t = ")
    mod = types.ModuleType("ref_models")
    mod.__file__ = path
    tvm.vgg19 = vgg19_no_download
    try:
        exec(compile(src, path, "exec"), mod.__dict__)
        # PretrainedEncoder.__init__ looks vgg19 up through the module-level ``models`` alias
        # at call time, so keep a shimmed alias inside the exec'd module.
        shim = types.SimpleNamespace(vgg19=vgg19_no_download)
        mod.__dict__["models"] = shim
    finally:
        tvm.vgg19 = orig_vgg19
    _cache["models"] = mod
    return mod


def load_reference_module(name: str) -> types.ModuleType:
    """Import one of the reference modules that import cleanly as shipped
    (conf, model_util, losses, mobilenetv2)."""
    assert name in ("conf", "model_util", "losses", "mobilenetv2")
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_DIR}")
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib

    return importlib.import_module(name)


def build_reference_classic_decoder():
    """exec the commented decoder spec at models.py:598-628 -> nn.Sequential (29 layers)."""
    import torch.nn as nn

    path = os.path.join(REF_DIR, "models.py")
    with open(path, "r") as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("# decoder = nn.Sequential("))
    end = next(i for i in range(start, len(lines)) if lines[i].startswith("# ).to(device)"))
    body = [l[2:] for l in lines[start:end]] + [")"]
    ns = {"nn": nn}
    exec("\n".join(body), ns)
    return ns["decoder"]
