import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The in-tree .so is prebuilt by __graft_entry__.build(); build it when a checkout lacks it."""
    lib = os.path.join(ROOT, "matrel_b200", "libmatrel_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def session():
    import matrel_b200 as mb
    s = mb.MatfastSession(device=0)
    yield s
    s.stop()
