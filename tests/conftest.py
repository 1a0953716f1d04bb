import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_DATA = "/root/reference/bioimitation/imitation_envs/data"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: reads /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REF_DATA)
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this machine")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def models():
    """Compiled models from the committed table fixtures (no /root/reference)."""
    from bioimitation_gym_b200 import assets
    return {k: assets.load_model(k) for k in assets.MODEL_KEYS}


@pytest.fixture
def rng():
    return np.random.default_rng(1234)
