import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sqb():
    from _sqpkg import sqb as pkg
    return pkg


@pytest.fixture(scope="session")
def port():
    import oracle_py
    oracle_py.build(ref=True)
    return oracle_py.PortOracle()


@pytest.fixture(scope="session")
def ref_available():
    import oracle_py
    return oracle_py.have_ref()


@pytest.fixture(scope="session")
def gpu_lib(sqb):
    """the CUDA library, loaded; GPU tests must run native code, never a fallback"""
    lib = sqb.load_library()
    n = lib.sq_device_count()
    assert n > 0, "gpu test selected but no CUDA device is visible"
    return lib
