import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

REFERENCE = "/root/reference"
REF_TLA = os.path.join(REFERENCE, "vsr-revisited/paper/VSR.tla")
REF_CFG = os.path.join(REFERENCE, "vsr-revisited/paper/VSR.cfg")
REF_TRACE = os.path.join(REFERENCE, "state_transfer_violation_trace.txt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds the native library if it is missing."""
    import _pkg
    so = os.path.join(ROOT, "vsr-tlaplus_b200", "libvsr_b200.so")
    if not os.path.exists(so) or not os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        import __graft_entry__
        __graft_entry__.build()
    return _pkg.load()


@pytest.fixture(scope="session")
def have_reference():
    return os.path.exists(REF_TLA)


needs_reference = pytest.mark.skipif(not os.path.exists(REF_TLA), reason="/root/reference is not mounted here")
