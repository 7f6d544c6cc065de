import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kmer_profile_golden.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def engine():
    """One GPU context for the whole session; fails loudly if the CUDA library is absent."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a GPU"
    from karma_b200.engine import Engine
    return Engine(0)
