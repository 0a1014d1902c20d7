import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (and the CUDA library if it is not there yet) once per session."""
    import __graft_entry__ as g
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "liborc.so")) or not os.path.exists(
            os.path.join(root, "kma_b200", "libkmagpu.so")):
        g.build()
    yield
