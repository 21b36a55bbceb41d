import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure both shared libraries exist (built in-tree; the .so files are git-ignored)."""
    import hare_b200
    from oracle import hare_oracle
    hare_b200.build()
    hare_oracle.build()


def has_gpu():
    try:
        import hare_b200
        return hare_b200.lib().hare_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    import hare_b200
    if not has_gpu():
        pytest.fail("no CUDA device visible: -m gpu tests need a B200 (there is no CPU fallback)")
    hare_b200.init([0])
    return hare_b200
