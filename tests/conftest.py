import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled headless (oracle/_ref/libbq2ref.so); skip when not built."""
    import oracle_lib
    r = oracle_lib.Reference.load()
    if r is None:
        pytest.skip("oracle/_ref/libbq2ref.so not built (needs /root/reference; run `make -C oracle ref`)")
    return r


@pytest.fixture(scope="session")
def gpu_ctx():
    import berserkerquake2_b200 as bq2
    ctx = bq2.Context(0)      # raises Bq2Error without a B200: there is no CPU path to fall back to
    yield ctx
    ctx.close()
