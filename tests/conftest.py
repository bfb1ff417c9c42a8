import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    from minirender_b200 import cabi
    if not os.path.exists(cabi.LIB_PATH):
        from minirender_b200 import build
        build.build_product()


@pytest.fixture(scope="session")
def be():
    """The product (libminirender_b200.so). Loading needs no GPU; rendering does."""
    _ensure_built()
    import minirender_b200 as m
    return m.Backend()


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled unchanged (oracle/_ref), if present."""
    import minirender_b200 as m
    import pyoracle
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/libminirender_ref.so not built (needs /root/reference)")
    return m.Backend(pyoracle.REF_PATH)


@pytest.fixture(scope="session")
def have_gpu():
    _ensure_built()
    from minirender_b200 import cabi
    return cabi.load().mr_device_count() > 0
