"""pytest configuration: the ``gpu`` marker, repo-root import path and golden-fixture loading."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


@pytest.fixture(scope="session")
def golden_stats():
    return load_golden("stats_adain")


@pytest.fixture(scope="session")
def golden_losses():
    return load_golden("losses")


@pytest.fixture(scope="session")
def golden_networks():
    return load_golden("networks")


@pytest.fixture(scope="session")
def golden_big():
    """The classic path at the benchmarked sizes (512x512 config 4, 2048x2048 config 5), made by the genuine reference."""
    return load_golden("bigsizes")


@pytest.fixture(scope="session")
def golden_ae256():
    """The genuine AutoEncoder at config 3's resolution (256x256, batch 2)."""
    return load_golden("autoencoder256")
