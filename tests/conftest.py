import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import itensor_b200 as itb

    c = itb.Context(0)
    yield c


@pytest.fixture(scope="session", autouse=True)
def _built():
    # the product library must exist; tests never build a fallback
    import itensor_b200

    itensor_b200.lib()
    from oracle import orc

    orc.oracle()
