import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "h264-jm-commentary_b200", ROOT / "oracle", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as _o  # oracle/oracle.py
    _o.build()
    return _o.load()


@pytest.fixture(scope="session")
def cuda():
    """The product library; GPU tests fail loudly if it is missing or no device is present."""
    import torch
    import jmme
    assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
    return jmme.load()
