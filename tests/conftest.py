import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle(70)


@pytest.fixture(scope="session")
def reference():
    from oracle.oracle_py import Reference
    if not Reference.available():
        try:
            from oracle.oracle_py import build
            build()
        except Exception:
            pass
    if not Reference.available():
        pytest.skip("oracle/_ref/libyama_ref.so not built (needs /root/reference)")
    return Reference(70)


@pytest.fixture(scope="session")
def yama_ctx():
    """One CUDA context for the whole GPU session.  Fails (not skips) if the extension is missing."""
    from multiz_b200 import YamaB200
    ctx = YamaB200(devices=[0])
    yield ctx
    ctx.close()
