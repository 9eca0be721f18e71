import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from tests import _util
    return _util.load_oracle()


@pytest.fixture(scope="session")
def ref():
    from tests import _util
    if not _util.have_reference_build():
        pytest.skip("oracle/_ref/libtrt_ref.so not built (needs /root/reference; see oracle/Makefile)")
    return _util.load_reference()


@pytest.fixture(scope="session")
def trt():
    """The product library, loaded but not initialised (no GPU needed)."""
    from terminalraytracer_b200 import build, lib
    if not os.path.exists(lib.LIB_PATH):
        build.build_library()
    return lib.load()


@pytest.fixture(scope="session")
def renderer(trt):
    """Initialised renderer on cuda:0 — only GPU-marked tests may request it."""
    from terminalraytracer_b200 import renderer as R
    r = R.Renderer(0)
    yield r
    r.close()
