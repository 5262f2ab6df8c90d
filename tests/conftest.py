import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if os.environ.get("ASVGF_TEST_HOSTSIM") == "1":
        # tests/test_hostsim.py re-runs a selection of the `-m gpu` tests in a child pytest against tests/hostsim/libasvgf_hostsim.so:
        # the library's own source compiled for the host (test infrastructure, tests/hostsim/README.md).  Only a test process
        # that sets this variable ever sees that file; raytracedgl1_b200 itself knows nothing about it.
        import hostsim
        from raytracedgl1_b200 import denoiser
        denoiser._LIB_PATH = hostsim.build()
        denoiser._TEST_ONLY_ALLOW_HOST_BUILD = True


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """The oracle and the synthetic-input generator are CPU code: build them on demand.  libasvgf.so (CUDA) is
    built by __graft_entry__.build(); tests never build or fall back for it."""
    import oracle
    import synth
    oracle.build()
    synth.build()
    yield
