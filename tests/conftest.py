import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_model():
    from oracle import pyoracle as po
    return po.Model.synthetic(0)


@pytest.fixture(scope="session")
def model_blob(oracle_model):
    return oracle_model.to_bytes()
