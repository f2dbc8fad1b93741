import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native artefacts exist (the driver normally runs __graft_entry__.build() first)."""
    import euler2d_kokkos_b200 as e2d

    if not os.path.exists(e2d.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    import oracle

    oracle.lib()
    yield
