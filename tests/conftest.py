import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pcq():
    from pcq_import import pcq as _pcq

    return _pcq


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as _o

    return _o


@pytest.fixture(scope="session")
def ctx(pcq):
    """one device context for the whole GPU session"""
    c = pcq.Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield c
    c.close()
