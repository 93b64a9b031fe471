import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def ctx():
    """Native context on cuda:0 -- fails loudly (no fallback) when the extension or the GPU is missing."""
    import torch
    from vae_tagger_b200 import _native

    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    return _native.get_context(0)


@pytest.fixture(scope="session")
def golden():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "head_golden.pt"), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def train_golden():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "head_train_golden.pt"), map_location="cpu",
                      weights_only=False)
