import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def cloud(rng, b, n, scale=1.0):
    """Synthetic cloud as in SURVEY.md 8(d): points uniform in [-0.5, 0.5]^3, float32."""
    return ((rng.random((b, n, 3), dtype=np.float32) - 0.5) * scale).astype(np.float32)


@pytest.fixture
def rng():
    return np.random.default_rng(1234)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rfnet_b200 import _lib
    _lib.load()  # fail loudly if the extension is missing on a GPU box
    return torch.device("cuda:0")
