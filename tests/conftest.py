import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle(fen=1, hadme=1)


@pytest.fixture(scope="session")
def oracle_nofen():
    from oracle.pyoracle import Oracle
    return Oracle(fen=0, hadme=0)


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference; skipped where oracle/_ref was not built (needs /root/reference)."""
    from oracle.pyoracle import Reference
    try:
        return Reference(fen=1, hadme=1)
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"reference library unavailable: {e}")


@pytest.fixture(scope="session")
def hm():
    """The product library, initialised on cuda:0.  GPU tests only."""
    from video_codecs_b200 import HMB200
    h = HMB200()
    h.init(0)
    yield h
    h.shutdown()
