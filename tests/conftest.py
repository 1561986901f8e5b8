import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` need a CUDA device: without one they are skipped (not errored), whatever -m selects."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory nerf-and-dietnerf_b200/), with the C-ABI library built and loaded."""
    import importlib
    import __graft_entry__ as entry
    if not os.path.exists(entry.LIB):
        entry.build()
    p = importlib.import_module("nerf-and-dietnerf_b200")
    p.load()
    return p
