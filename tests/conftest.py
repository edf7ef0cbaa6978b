import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The parity oracle (test infrastructure).  Builds the C restatement on first use."""
    import oracle
    if not os.path.exists(oracle.PORT_LIB):
        oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu_ctx():
    import defuse_b200
    return defuse_b200.default_context(0)
