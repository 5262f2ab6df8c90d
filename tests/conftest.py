import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """The oracle and the synthetic-input generator are CPU code: build them on demand.  libasvgf.so (CUDA) is
    built by __graft_entry__.build(); tests never build or fall back for it."""
    import oracle
    import synth
    oracle.build()
    synth.build()
    yield
