import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def est_sd():
    from jyutvoice_b200 import synthetic as weights
    return weights.make_estimator_state_dict()


@pytest.fixture(scope="session")
def hift_sd():
    from jyutvoice_b200 import synthetic as weights
    return weights.make_hift_state_dict()


@pytest.fixture(scope="session")
def hift_sd_voiced():
    from jyutvoice_b200 import synthetic as weights
    return weights.make_hift_state_dict(f0_bias=200.0)


@pytest.fixture(scope="session")
def noise_bank():
    from jyutvoice_b200 import synthetic as weights
    return weights.noise_bank()


def snr_db(ref, x):
    ref = ref.double()
    x = x.double()
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - x) ** 2).sum().clamp_min(1e-300)))
