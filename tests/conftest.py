import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference in oracle/_ref")


def _have_gpu():
    try:
        import smolscale_b200 as sb
        return sb.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def sb():
    import __graft_entry__ as ge
    ge.build()
    import smolscale_b200
    return smolscale_b200


@pytest.fixture(scope="session")
def restatement():
    import oracle
    return oracle.restatement()


@pytest.fixture(scope="session")
def reference():
    import oracle
    r = oracle.reference()
    if r is None:
        pytest.skip("oracle/_ref/libsmolref.so not present (built only where /root/reference exists)")
    return r


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
