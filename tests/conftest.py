import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        # a hung kernel must fail its test, not eat the whole GPU session (pytest-timeout is in the image)
        if config.pluginmanager.hasplugin("timeout"):
            for it in items:
                if "gpu" in it.keywords and it.get_closest_marker("timeout") is None:
                    it.add_marker(pytest.mark.timeout(600))
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def kat():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "ops_kat.npz"))
