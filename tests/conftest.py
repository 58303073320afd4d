import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault("HOME", os.environ.get("HOME", "/tmp"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box")


@pytest.fixture(scope="session")
def lib():
    from fish_tts_b200 import _build, capi
    _build.build()
    return capi.load()
