import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def constants():
    from smalify_b200 import model_io
    return model_io.load_asset()


@pytest.fixture(scope="session")
def oracle64(constants):
    import torch
    from oracle import smal_oracle as O
    return O.OracleModel.from_constants(constants, torch.float64)
