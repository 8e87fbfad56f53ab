import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture
def tuning():
    """Set QSAE_* tuning switches for one test: tuning(NAME, value) sets the variable and makes the library
    re-read its switches (they are cached, never read on a launch path); everything is restored afterwards."""
    from quantizedsae_b200 import _lib as L

    saved = {}

    def set_(name, value):
        if name not in saved:
            saved[name] = os.environ.get(name)
        os.environ[name] = str(value)
        L.check(L.load().qsae_reload_tuning())

    yield set_
    for name, old in saved.items():
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old
    if saved:
        L.check(L.load().qsae_reload_tuning())
